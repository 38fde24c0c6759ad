// Batch-1 decoder step as ONE persistent cooperative kernel (reference: magpie_build_decoder_layer_gpu_cached
// src/magpie.cpp:3484-3528, self-attention 3395-3480, cross-attention 1713-1767, conv-FFN 1769-1810,
// embedding 2746-2787, driver loop 4366-4405).
//
// The reference issues ~470 ggml nodes per step; the v1 path here issued 98 kernels.  At batch 1 the
// step is a chain of 7 dependent matrix-vector products per layer, each needing the *whole* previous
// vector, so the step is bound by (a) streaming 175 MB of weights and (b) 84 all-to-all exchanges.
// Design:
//   * grid = one CTA per SM (cooperative launch), every GEMV is row-sliced over all CTAs;
//   * each CTA's weight slices are known in advance and independent of the activations, so they are
//     PREFETCHED into a shared-memory ring with cp.async.bulk (TMA 1-D bulk copies completing on
//     mbarriers), up to ~1.8 layers ahead: HBM streaming runs in the background of the dependency
//     chain and the exposed time per phase is barrier + prologue latency only;
//   * phases are separated by a hand-rolled grid barrier (red.release / ld.acquire on one counter);
//     the prefetch for later phases is issued from inside the barrier, between arrive and spin;
//   * activations (<= 12 KB) are exchanged through L2 (st / ld.global.cg), never through L1.
// All reductions are in a fixed order => bitwise deterministic run to run.
#include <cooperative_groups.h>

#include "common.cuh"
#include "kernels.cuh"
#include "mega.h"

namespace mgb {

namespace {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kQD = 16;                   // prefetch queue depth (slices in flight)
constexpr int kRingBytes = 176 * 1024;    // weight ring
constexpr int kScratchFloats = 11 * 1024; // 44 KB: activation vectors / attention partials
constexpr int kMaxSplit = 16;

// ---- PTX helpers ----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void * p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t * bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t * bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t * bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
// Weights are read exactly once per step: stream them through L2 with an evict-first policy so that they
// do not displace the KV cache, the activations and the local-transformer weights, which ARE re-read.
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s(void * dst, const void * src, uint32_t bytes, uint64_t * bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void red_release_add(unsigned * p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned * p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ldcg(const float * p) { return __ldcg(p); }

template <typename T> struct Ld16;      // 16-byte load of weights/KV -> floats, from any address space
template <> struct Ld16<float> {
    static constexpr int VEC = 4;
    __device__ static __forceinline__ void smem(const float * p, float (&w)[4]) {
        const float4 v = *reinterpret_cast<const float4 *>(p); w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
    }
    __device__ static __forceinline__ void gcg(const float * p, float (&w)[4]) {
        const float4 v = __ldcg(reinterpret_cast<const float4 *>(p)); w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
    }
};
template <> struct Ld16<__nv_bfloat16> {
    static constexpr int VEC = 8;
    __device__ static __forceinline__ void cvt(const uint4 u, float (&w)[8]) {
        w[0] = bf16lo(u.x); w[1] = bf16hi(u.x); w[2] = bf16lo(u.y); w[3] = bf16hi(u.y);
        w[4] = bf16lo(u.z); w[5] = bf16hi(u.z); w[6] = bf16lo(u.w); w[7] = bf16hi(u.w);
    }
    __device__ static __forceinline__ void smem(const __nv_bfloat16 * p, float (&w)[8]) { cvt(*reinterpret_cast<const uint4 *>(p), w); }
    __device__ static __forceinline__ void gcg(const __nv_bfloat16 * p, float (&w)[8]) { cvt(__ldcg(reinterpret_cast<const uint4 *>(p)), w); }
};

// ---- shared-memory state --------------------------------------------------------------------------
struct alignas(128) MegaSmem {
    unsigned char ring[kRingBytes];
    float scratch[kScratchFloats];
    float red[32];
    float vec[1024 + 128];          // LN'd / staged input vector
    uint64_t mbar[kQD];
    int q_off[kQD];                 // ring offset of slice (weighted-phase index % kQD), -1 = empty slice
    int q_par[kQD];                 // mbarrier parity to wait for
    int q_use[kQD];                 // completed uses of each mbarrier (thread 0)
    // prefetch bookkeeping (thread 0 only)
    int issued, consumed, head, tail, count;
};

struct Slice { const unsigned char * src; uint32_t bytes; int r0, nr; };

template <typename T>
__device__ __forceinline__ Slice make_slice(const T * W, int N, int K, int G, int b) {
    const int rpc = (N + G - 1) / G;
    Slice s;
    s.r0 = b * rpc;
    s.nr = max(0, min(rpc, N - s.r0));
    s.bytes = (uint32_t)s.nr * K * sizeof(T);
    s.src = reinterpret_cast<const unsigned char *>(W + (size_t)s.r0 * K);
    return s;
}

template <typename T>
__device__ __forceinline__ Slice phase_slice(const MegaParams & p, int wp, int G, int b) {
    const int l = wp / 6, j = wp % 6;
    const MegaLayer & L = p.layer[l];
    const int d = p.d, f = p.f, dxa = p.dxa;
    switch (j) {
        case 0: return make_slice<T>((const T *)L.qkv, 3 * d, d, G, b);
        case 1: return make_slice<T>((const T *)L.o, d, d, G, b);
        case 2: return make_slice<T>((const T *)L.xq, dxa, d, G, b);
        case 3: return make_slice<T>((const T *)L.xo, d, dxa, G, b);
        case 4: return make_slice<T>((const T *)L.ff1, f, d, G, b);
        default: return make_slice<T>((const T *)L.ff2, d, f, G, b);
    }
}

// thread 0: issue upcoming weight slices into the ring.  Throttled (max_new per call) so that the stream is
// spread over the step instead of arriving in bursts that queue in front of the phases' demand loads.
template <typename T>
__device__ void prefetch_issue(MegaSmem & S, const MegaParams & p, int G, int b, int max_new) {
    const int total = p.L * 6;
    const uint64_t pol = l2_evict_first_policy();
    int n_new = 0;
    while (S.issued < total && S.issued - S.consumed < kQD && n_new < max_new) {
        const Slice s = phase_slice<T>(p, S.issued, G, b);
        const int slot = S.issued % kQD;
        if (s.bytes == 0) { S.q_off[slot] = -1; S.issued++; continue; }
        const int need = (int)((s.bytes + 127u) & ~127u);
        int off = -1;
        if (S.count == 0) { S.head = S.tail = 0; }
        if (S.count == 0 || S.tail > S.head) {
            if (S.tail + need <= kRingBytes) { off = S.tail; S.tail += need; }
            else if (S.count > 0 && need < S.head) { off = 0; S.tail = need; }
            else if (S.count == 0 && need <= kRingBytes) { off = 0; S.tail = need; }
        } else if (S.tail < S.head) {
            if (S.tail + need <= S.head) { off = S.tail; S.tail += need; }
        }
        if (off < 0) break;
        if (S.count == 0) S.head = off;
        S.count++;
        S.q_off[slot] = off;
        S.q_par[slot] = S.q_use[slot] & 1; S.q_use[slot]++;
        mbar_expect_tx(&S.mbar[slot], s.bytes);
        // chunks of <= 64 KB per bulk copy
        uint32_t done = 0;
        while (done < s.bytes) {
            const uint32_t n = min(s.bytes - done, 65536u);
            bulk_g2s(S.ring + off + done, s.src + done, n, &S.mbar[slot], pol);
            done += n;
        }
        S.issued++; n_new++;
    }
}

// thread 0: the oldest outstanding slice (weighted phase wp) has been consumed
__device__ __forceinline__ void prefetch_release(MegaSmem & S, int wp) {
    if (S.q_off[wp % kQD] >= 0) {
        S.count--;
        // next oldest non-empty slice defines the new head
        int nxt = wp + 1;
        while (nxt < S.issued && S.q_off[nxt % kQD] < 0) nxt++;
        S.head = (S.count > 0 && nxt < S.issued) ? S.q_off[nxt % kQD] : S.tail;
    }
    S.consumed = wp + 1;
}

// Grid barrier.  `consumed_wp` >= 0: the weighted phase that just finished (its ring space is recycled
// and new prefetches are issued while waiting).
template <typename T>
__device__ __forceinline__ void grid_barrier(MegaSmem & S, const MegaParams & p, unsigned & epoch, int G, int b, int consumed_wp) {
    __syncthreads();
    if (threadIdx.x == 0) {
        if (p.dbg) {       // profiling aid: global arrival time of every CTA at every barrier
            unsigned long long t;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
            p.dbg[1024 + (size_t)(epoch / (unsigned)G) * G + b] = t;
        }
        epoch += (unsigned)G;
        red_release_add(p.barrier, 1u);
        if (consumed_wp >= 0) prefetch_release(S, consumed_wp);
        fence_proxy_async();
        prefetch_issue<T>(S, p, G, b, (S.issued - S.consumed <= 3) ? 3 : 1);
        while (ld_acquire(p.barrier) < epoch) { }
    }
    __syncthreads();
}

// wait for the weight slice of weighted phase wp; returns its ring pointer (nullptr if this CTA has no rows)
__device__ __forceinline__ const unsigned char * slice_wait(MegaSmem & S, int wp) {
    const int slot = wp % kQD;
    const int off = S.q_off[slot];
    if (off < 0) return nullptr;
    mbar_wait(&S.mbar[slot], (uint32_t)S.q_par[slot]);
    return S.ring + off;
}

// y[r] = sum_k W[r][k] x[k] for the nr rows of a slice held in shared memory; epi(global_row, value)
template <typename T, typename Epi>
__device__ __forceinline__ void smem_gemv(const unsigned char * wbytes, int nr, int r0, int K, const float * x, Epi epi) {
    constexpr int VEC = Ld16<T>::VEC;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const T * W = reinterpret_cast<const T *>(wbytes);
    for (int r = warp; r < nr; r += kWarps) {
        const T * wr = W + (size_t)r * K;
        float acc0 = 0.0f, acc1 = 0.0f;
        for (int k = lane * VEC; k < K; k += 32 * VEC) {
            float w[VEC];
            Ld16<T>::smem(wr + k, w);
#pragma unroll
            for (int v4 = 0; v4 < VEC; v4 += 4) {
                const float4 xv = *reinterpret_cast<const float4 *>(x + k + v4);
                acc0 = fmaf(w[v4], xv.x, acc0); acc1 = fmaf(w[v4 + 1], xv.y, acc1);
                acc0 = fmaf(w[v4 + 2], xv.z, acc0); acc1 = fmaf(w[v4 + 3], xv.w, acc1);
            }
        }
        const float v = warp_sum(acc0 + acc1);
        if (lane == 0) epi(r0 + r, v);
    }
}

// vec = LN(src) * w   (src: global, read through L2; two-pass like ggml_norm)
__device__ __forceinline__ void load_layer_norm(MegaSmem & S, const float * src, const float * w, int d, float eps, bool from_smem) {
    const int tid = threadIdx.x;
    float v0 = 0.0f, v1 = 0.0f;
    const int i0 = tid, i1 = tid + kThreads;
    if (i0 < d) v0 = from_smem ? src[i0] : ldcg(src + i0);
    if (i1 < d) v1 = from_smem ? src[i1] : ldcg(src + i1);
    const float mean = block_sum(v0 + v1, S.red) / (float)d;
    const float c0 = i0 < d ? v0 - mean : 0.0f, c1 = i1 < d ? v1 - mean : 0.0f;
    const float var = block_sum(c0 * c0 + c1 * c1, S.red) / (float)d;
    const float scale = 1.0f / sqrtf(var + eps);
    if (i0 < d) S.vec[i0] = (c0 * scale) * w[i0];
    if (i1 < d) S.vec[i1] = (c1 * scale) * w[i1];
    __syncthreads();
}

template <typename T>
__global__ void __launch_bounds__(kThreads, 1) decoder_mega_kernel(const MegaParams p) {
    extern __shared__ __align__(128) unsigned char mega_smem[];
    MegaSmem & S = *reinterpret_cast<MegaSmem *>(mega_smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int G = gridDim.x, b = blockIdx.x;
    const int d = p.d, f = p.f, dxa = p.dxa, H = p.H, dh = d / H;
    const int pos = *p.pos;
    const int nk = pos + 1;
    unsigned epoch = 0;
    int dbg_i = 0;
#define MEGA_STAMP() do { if (p.dbg && b == 0 && tid == 0) p.dbg[dbg_i++] = (unsigned long long)clock64(); } while (0)
    MEGA_STAMP();

    if (tid == 0) {
        for (int i = 0; i < kQD; i++) { mbar_init(&S.mbar[i], 1); S.q_off[i] = -1; S.q_par[i] = 0; S.q_use[i] = 0; }
        S.issued = S.consumed = S.head = S.tail = S.count = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) { fence_proxy_async(); prefetch_issue<T>(S, p, G, b, 4); }

    // x0 = (sum_cb E_cb[code_cb]) / 8 + pos_emb[pos]: computed redundantly by every CTA (kept in scratch),
    // CTA 0 publishes it for the residual adds of later phases.
    float * x0 = S.scratch;
    for (int i = tid; i < d; i += kThreads) {
        float s = p.audio_emb[0][(size_t)p.codes[0] * d + i];
#pragma unroll
        for (int cb = 1; cb < 8; cb++) s = s + p.audio_emb[cb][(size_t)p.codes[cb] * d + i];
        const float v = s * 0.125f + p.dec_pos[(size_t)pos * d + i];
        x0[i] = v;
        if (b == 0) p.x[i] = v;
    }
    __syncthreads();

    const int S_split = p.n_split;
    for (int l = 0; l < p.L; l++) {
        const MegaLayer & L = p.layer[l];
        T * kc = (T *)p.kcache + (size_t)l * p.kv_layer_stride;
        T * vc = (T *)p.vcache + (size_t)l * p.kv_layer_stride;
        const int wp0 = l * 6;

        // ---- P1: LN -> QKV; q to global, K/V straight into the cache row `pos` ----------------------
        load_layer_norm(S, l == 0 ? x0 : p.x, L.n_self, d, p.eps, l == 0);
        {
            const Slice sl = phase_slice<T>(p, wp0 + 0, G, b);
            const unsigned char * w = slice_wait(S, wp0 + 0);
            if (w) smem_gemv<T>(w, sl.nr, sl.r0, d, S.vec, [&](int n, float v) {
                if (n < d) p.q[n] = v;
                else if (n < 2 * d) WT<T>::put(kc + (size_t)pos * d + (n - d), v);
                else WT<T>::put(vc + (size_t)pos * d + (n - 2 * d), v);
            });
        }
        MEGA_STAMP(); grid_barrier<T>(S, p, epoch, G, b, wp0 + 0); MEGA_STAMP();

        // ---- P2: attention partials: item (head, key split) -----------------------------------------
        if (b < H * S_split) {
            const int h = b / S_split, sp = b % S_split;
            const int per = (nk + S_split - 1) / S_split;
            const int k0 = sp * per, k1 = min(nk, k0 + per);
            float * sq = S.vec;                       // q_h
            if (tid < dh) sq[tid] = ldcg(p.q + h * dh + tid);
            __syncthreads();
            constexpr int VEC = Ld16<T>::VEC;
            const float scale = 1.0f / sqrtf((float)dh);
            float mx = -INFINITY, lsum = 0.0f, acc0 = 0.0f, acc1 = 0.0f;     // lane owns dims 2*lane, 2*lane+1 (dh == 64)
            for (int c0 = k0 + warp * 32; c0 < k1; c0 += kWarps * 32) {
                const int j = c0 + lane;
                float s = -INFINITY;
                if (j < k1) {
                    const T * kr = kc + (size_t)j * d + h * dh;
                    float dsum = 0.0f;
#pragma unroll
                    for (int c = 0; c < 64; c += VEC) {
                        float kv[VEC];
                        Ld16<T>::gcg(kr + c, kv);
#pragma unroll
                        for (int v = 0; v < VEC; v++) dsum = fmaf(kv[v], sq[c + v], dsum);
                    }
                    s = dsum * scale;
                }
                const float mnew = fmaxf(mx, warp_max(s));
                const float corr = expf(mx - mnew);
                const float pj = (j < k1) ? expf(s - mnew) : 0.0f;
                lsum = lsum * corr + warp_sum(pj);
                acc0 *= corr; acc1 *= corr;
                // P.V: all 32 value-row loads of the chunk are issued before the first FMA (latency-bound otherwise)
                const int cnt = min(32, k1 - c0);
                const T * vbase = vc + (size_t)c0 * d + h * dh + lane * 2;
                float v0[32], v1[32];
#pragma unroll
                for (int jj = 0; jj < 32; jj++) {
                    v0[jj] = 0.0f; v1[jj] = 0.0f;
                    if (jj < cnt) {
                        if constexpr (sizeof(T) == 2) {
                            const uint32_t u = __ldcg(reinterpret_cast<const uint32_t *>(vbase + (size_t)jj * d));
                            v0[jj] = bf16lo(u); v1[jj] = bf16hi(u);
                        } else {
                            const float2 u = __ldcg(reinterpret_cast<const float2 *>(vbase + (size_t)jj * d));
                            v0[jj] = u.x; v1[jj] = u.y;
                        }
                    }
                }
#pragma unroll
                for (int jj = 0; jj < 32; jj++) {
                    const float pb = __shfl_sync(0xffffffffu, pj, jj);
                    acc0 = fmaf(pb, v0[jj], acc0); acc1 = fmaf(pb, v1[jj], acc1);
                }
                mx = mnew;
            }
            // merge the warps of this CTA (fixed order), publish (m, l, acc[64])
            float * sm = S.scratch + 1024, * slv = sm + kWarps, * sacc = slv + kWarps;     // [kWarps][64]
            if (lane == 0) { sm[warp] = mx; slv[warp] = lsum; }
            sacc[warp * 64 + lane * 2] = acc0; sacc[warp * 64 + lane * 2 + 1] = acc1;
            __syncthreads();
            if (tid < 64) {
                float M = sm[0];
                for (int w = 1; w < kWarps; w++) M = fmaxf(M, sm[w]);
                float Lsum = 0.0f, o = 0.0f;
                for (int w = 0; w < kWarps; w++) {
                    const float fct = (sm[w] == -INFINITY) ? 0.0f : expf(sm[w] - M);
                    Lsum += fct * slv[w]; o += fct * sacc[w * 64 + tid];
                }
                float * part = p.attn_part + (size_t)b * 66;
                part[2 + tid] = o;
                if (tid == 0) { part[0] = M; part[1] = Lsum; }
            }
        }
        MEGA_STAMP(); grid_barrier<T>(S, p, epoch, G, b, -1); MEGA_STAMP();

        // ---- P3: combine partials -> attn[d]; O projection + residual (in place on x) -----------------
        {
            const int np = H * S_split * 66;
            float * part = S.scratch;
            for (int i = tid; i < np; i += kThreads) part[i] = ldcg(p.attn_part + i);
            __syncthreads();
            for (int i = tid; i < d; i += kThreads) {
                const int h = i / dh, dd = i % dh;
                const float * ph = part + (size_t)h * S_split * 66;
                float M = -INFINITY;
                for (int s2 = 0; s2 < S_split; s2++) M = fmaxf(M, ph[s2 * 66]);
                float Lsum = 0.0f, o = 0.0f;
                for (int s2 = 0; s2 < S_split; s2++) {
                    const float m2 = ph[s2 * 66];
                    const float fct = (m2 == -INFINITY) ? 0.0f : expf(m2 - M);
                    Lsum += fct * ph[s2 * 66 + 1]; o += fct * ph[s2 * 66 + 2 + dd];
                }
                S.vec[i] = o * (1.0f / Lsum);
            }
            __syncthreads();
            const Slice sl = phase_slice<T>(p, wp0 + 1, G, b);
            const unsigned char * w = slice_wait(S, wp0 + 1);
            if (w) smem_gemv<T>(w, sl.nr, sl.r0, d, S.vec, [&](int n, float v) { p.x[n] = v + ldcg(p.x + n); });
        }
        // (layer 0: the residual is the copy of x0 that CTA 0 published before the first barrier)
        MEGA_STAMP(); grid_barrier<T>(S, p, epoch, G, b, wp0 + 1); MEGA_STAMP();

        // ---- P4: LN -> cross-attention query -----------------------------------------------------------
        load_layer_norm(S, p.x, L.n_xq, d, p.eps, false);
        {
            const Slice sl = phase_slice<T>(p, wp0 + 2, G, b);
            const unsigned char * w = slice_wait(S, wp0 + 2);
            if (w) smem_gemv<T>(w, sl.nr, sl.r0, d, S.vec, [&](int n, float v) { p.xq[n] = v; });
        }
        MEGA_STAMP(); grid_barrier<T>(S, p, epoch, G, b, wp0 + 2); MEGA_STAMP();

        // ---- P5: cross-attention over the cached encoder K/V (redundant per CTA) + XO + residual ------
        {
            const int E = *p.n_ctx;
            const T * xk = (const T *)p.xk + (size_t)l * p.xkv_layer_stride;
            const T * xv = (const T *)p.xv + (size_t)l * p.xkv_layer_stride;
            float * sq = S.scratch;                  // [dxa]
            float * sc = S.scratch + 256;            // [E] scores (E <= 4096)
            if (tid < dxa) sq[tid] = ldcg(p.xq + tid);
            __syncthreads();
            const float scale = 1.0f / sqrtf((float)dxa);
            for (int j = warp; j < E; j += kWarps) {
                const T * kr = xk + (size_t)j * dxa;
                float dsum = 0.0f;
                for (int c = lane; c < dxa; c += 32) dsum = fmaf(WT<T>::get(kr + c), sq[c], dsum);
                dsum = warp_sum(dsum);
                if (lane == 0) sc[j] = dsum * scale;
            }
            __syncthreads();
            float mxs = -INFINITY;
            for (int j = 0; j < E; j++) mxs = fmaxf(mxs, sc[j]);
            if (tid < dxa) {
                float sum = 0.0f, o = 0.0f;
                for (int j = 0; j < E; j++) {
                    const float e = expf(sc[j] - mxs);
                    sum += e; o = fmaf(e, WT<T>::get(xv + (size_t)j * dxa + tid), o);
                }
                S.vec[tid] = o * (1.0f / sum);
            }
            __syncthreads();
            const Slice sl = phase_slice<T>(p, wp0 + 3, G, b);
            const unsigned char * w = slice_wait(S, wp0 + 3);
            if (w) smem_gemv<T>(w, sl.nr, sl.r0, dxa, S.vec, [&](int n, float v) { p.x[n] = v + ldcg(p.x + n); });
        }
        MEGA_STAMP(); grid_barrier<T>(S, p, epoch, G, b, wp0 + 3); MEGA_STAMP();

        // ---- P6: LN -> FFN1 -> GELU ---------------------------------------------------------------------
        load_layer_norm(S, p.x, L.n_ff, d, p.eps, false);
        {
            const Slice sl = phase_slice<T>(p, wp0 + 4, G, b);
            const unsigned char * w = slice_wait(S, wp0 + 4);
            if (w) smem_gemv<T>(w, sl.nr, sl.r0, d, S.vec, [&](int n, float v) { p.ffh[n] = gelu_ggml(v, p.gelu_f16); });
        }
        MEGA_STAMP(); grid_barrier<T>(S, p, epoch, G, b, wp0 + 4); MEGA_STAMP();

        // ---- P7: FFN2 + residual --------------------------------------------------------------------------
        {
            float * hbuf = S.scratch;                // [f] <= 11K floats
            for (int i = tid; i < f; i += kThreads) hbuf[i] = ldcg(p.ffh + i);
            __syncthreads();
            const Slice sl = phase_slice<T>(p, wp0 + 5, G, b);
            const unsigned char * w = slice_wait(S, wp0 + 5);
            if (w) smem_gemv<T>(w, sl.nr, sl.r0, f, hbuf, [&](int n, float v) { p.x[n] = v + ldcg(p.x + n); });
        }
        MEGA_STAMP(); grid_barrier<T>(S, p, epoch, G, b, wp0 + 5); MEGA_STAMP();
    }

    MEGA_STAMP();
    // final LayerNorm -> hidden (CTA 0), advance the position
    if (b == 0) {
        load_layer_norm(S, p.x, p.norm_out, d, p.eps, false);
        for (int i = tid; i < d; i += kThreads) p.hidden[i] = S.vec[i];
        if (tid == 0) { *p.pos_rw = pos + 1; *p.slot_rw = *p.slot_rw + 1; }
    }
    // leave the barrier counter at 0 for the next launch: the last CTA to pass the final barrier cannot be
    // identified cheaply, so the host-side launcher resets it with a memset node instead.
}

}  // namespace

size_t mega_smem_bytes() { return sizeof(MegaSmem) + 128; }

bool launch_decoder_mega(const MegaParams & p, int precision, int grid, cudaStream_t stream) {
    const size_t smem = mega_smem_bytes();
    void * kfn = precision == MGB_PREC_F32 ? (void *)decoder_mega_kernel<float> : (void *)decoder_mega_kernel<__nv_bfloat16>;
    static DeviceOnce attr_done[2];
    int dev = 0;
    MGB_CUDA_TRY(cudaGetDevice(&dev));
    const int pi = precision == MGB_PREC_F32 ? 0 : 1;
    if (!attr_done[pi].done(dev)) {
        MGB_CUDA_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done[pi].set(dev);
    }
    MGB_CUDA_TRY(cudaMemsetAsync(p.barrier, 0, sizeof(unsigned), stream));
    MegaParams pc = p;
    void * args[] = {(void *)&pc};
    MGB_CUDA_TRY(cudaLaunchCooperativeKernel(kfn, dim3(grid), dim3(kThreads), args, smem, stream));
    MGB_LAUNCH_CHECK();
    return true;
}

int mega_max_grid(int precision) {
    const size_t smem = mega_smem_bytes();
    void * kfn = precision == MGB_PREC_F32 ? (void *)decoder_mega_kernel<float> : (void *)decoder_mega_kernel<__nv_bfloat16>;
    int dev = 0, sms = 0, per_sm = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, kThreads, smem) != cudaSuccess) return 0;
    return per_sm >= 1 ? sms : 0;
}

}  // namespace mgb
