// Batch-1 synthesis loop as ONE persistent cooperative kernel: for every frame the 12-layer decoder step
// (reference: magpie_build_decoder_layer_gpu_cached src/magpie.cpp:3484-3528, self-attention 3395-3480,
// cross-attention 1713-1767, conv-FFN 1769-1810, embedding 2746-2787, loop 4321-4407) followed by the local
// transformer with its sampler (magpie.cpp:946-1048, 1072-1317) and the EOS rule (4341-4352).
//
// At batch 1 a frame is a chain of ~115 dependent matrix-vector products, each needing the whole previous
// vector: the frame is bound by (a) streaming 175 MB of bf16 weights and (b) the latency of ~115 all-to-all
// exchanges between the SMs.  Design:
//   * grid = one CTA per SM (cooperative launch only to guarantee co-residency), 15 compute warps + 1 weight
//     prefetch warp per CTA.  Every GEMV is row-sliced over all CTAs.
//   * weights: each CTA's slices are known in advance, so the prefetch warp streams them with cp.async.bulk
//     (evict-first in L2) into a shared-memory ring, full/empty mbarriers per slice, running ahead of the
//     dependency chain across layer AND frame boundaries; the local transformer's layer matrices stay
//     resident in shared memory for the whole launch.
//   * exchanges: no grid barrier.  A producer writes its output rows as 16-byte packets {v0,v1,v2,flag} with
//     the exchange's sequence number as flag (one vector store, 4 replicas to spread the polling over L2
//     slices); consumers spin on the packets themselves with volatile vector loads.  One L2 round trip per
//     exchange, no fences, no atomics; buffers are reused only after two later all-to-all exchanges, which
//     makes the reuse race-free.
//   * cross-attention (1 head over E <= 30 text tokens) is folded at prefill: M_l = scale K_l Wq_l and
//     N_l = V_l Wo_l^T, so that scores = M_l LN(x) and the output projection = softmax(scores) N_l; this
//     removes one exchange per layer (the reference's q_net / o_net GEMVs become part of the tables).
//   * the local transformer runs on all CTAs from shared memory; argmax is a 135-packet exchange, top-k
//     sampling gathers the logits and runs redundantly (deterministically) in every CTA.
// All reductions are in a fixed order => bitwise deterministic run to run.
#include <cooperative_groups.h>

#include "common.cuh"
#include "frame_loop.h"
#include "lt_common.cuh"

namespace mgb {

namespace {

using bf = __nv_bfloat16;

constexpr int kThreads = 512, kCW = 15, kCT = kCW * 32;      // warp 15 = weight prefetcher
constexpr int kR = 4;                                        // exchange replicas
constexpr int kQD = 8;                                       // ring slots
constexpr int kRingBytes = 112 * 1024;
constexpr int D = 768, F = 3072, H = 12, DH = 64, LD = 256, LF = 1024;
constexpr int kMaxSplit = 6;
// rows per CTA (multiples of 3 = one packet)
constexpr int RQ = 18, RO = 6, RF1 = 21, RF2 = 6, RIN = 3, RLQ = 6, RLO = 3, RLF1 = 9, RLF2 = 3, ROUT = 15;
constexpr int kSlicesPerFrameLayer = 4;
constexpr int kVecFloats = 4800;                             // >= max(F, H*kMaxSplit*66, 2025 + 2048)
constexpr int kPartStride = 12;

// ---- PTX helpers ----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void * p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t * bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t * bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t * bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t * bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t * bar, uint32_t parity) { while (!mbar_try(bar, parity)) { } }
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(void * dst, const void * src, uint32_t bytes, uint64_t * bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void * dst, const void * src, uint32_t bytes, uint64_t * bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void cbar() { asm volatile("bar.sync 1, %0;" ::"n"(kCT) : "memory"); }     // compute warps only
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// 16-byte flag packet {v0, v1, v2, flag}: one vector store / one vector load, both single L2 sector accesses
__device__ __forceinline__ uint4 ld_pkt(const uint4 * p) {
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_pkt(uint4 * p, float a, float b, float c, unsigned flag) {
    asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1,%2,%3,%4};"
                 ::"l"(p), "r"(__float_as_uint(a)), "r"(__float_as_uint(b)), "r"(__float_as_uint(c)), "r"(flag) : "memory");
}

enum { EPI_NONE = 0, EPI_RES = 1, EPI_GELU = 2, EPI_BIAS = 3 };
enum { PH_QKV = 0, PH_O, PH_FF1, PH_FF2, PH_LT_IN, PH_LT_QKV, PH_LT_O, PH_LT_FF1, PH_LT_FF2, NUM_PHASE_DESC };

// static description of a generic phase: poll inputs -> LayerNorm -> GEMV -> epilogue -> emit
struct PhaseDesc {
    int xin, npk; float * dst; int nfl;          // input exchange to poll -> dst[0, nfl)
    const float * ln_src; int nch;               // LayerNorm(ln_src) -> S.vec when the call passes ln weights
    const float * gin;                           // GEMV input vector (smem)
    const bf * w;                                // resident weights (smem) or nullptr = next ring slice
    int N, rpc, nseg;                            // rows of the matrix, rows per CTA, K / 256
    int epi; const float * res; const float * bias;
    int xout;
};

// ---- shared-memory state --------------------------------------------------------------------------
struct alignas(128) LoopSmem {
    unsigned char ring[kRingBytes];
    // local-transformer layer slices, resident for the whole launch
    bf w_in[RIN * D]; bf w_qkv[RLQ * LD]; bf w_o[RLO * LD]; bf w_ff1[RLF1 * LD]; bf w_ff2[RLF2 * LF];
    alignas(16) float xs[D];        // residual stream (full vector, refreshed by every x exchange)
    alignas(16) float vec[kVecFloats];   // staged GEMV input / attention partials / gathered logits
    alignas(16) float av[D];        // combined attention output; final hidden
    alignas(16) float part[24 * kPartStride];   // GEMV partial sums [row][k segment]
    alignas(16) float am[16], al[16], aacc[kCW * 64], aout[68];
    alignas(16) float qh[DH];
    alignas(16) float lx[LD], lx1[LD], latt[LD], lhout[LD];
    alignas(16) float lqkv[8][3 * LD];   // local transformer: [q | k | v] of every position of the frame
    alignas(16) float ltpos[8 * LD];
    alignas(16) float sc[32];
    alignas(16) float outv[24];
    alignas(16) float red[32]; int redi[32];
    alignas(16) float sel_v[2048]; uint16_t sel_i[2048], srt_i[2048], rank[2048];
    unsigned hist[256]; int misc[8];
    uint64_t full_bar[kQD], empty_bar[kQD], res_bar;
    int q_off[kQD];
    volatile int stop;
    int dbg_on, dbg_i;
    // launch constants the non-inlined helpers need (kernel parameters must not be passed by reference: that would
    // copy the 2 KB parameter block into every thread's local memory)
    uint4 * xbuf; size_t rstride; int xoff[X_COUNT + 1];
    float eps; int gelu_f16, E;
    unsigned long long * dbg;
    PhaseDesc desc[NUM_PHASE_DESC];
};

static_assert(sizeof(LoopSmem) + 128 <= 227 * 1024, "LoopSmem exceeds the 227 KB shared-memory limit");

struct Slice { const unsigned char * src; uint32_t bytes; int r0, nr; };

__device__ __forceinline__ Slice make_slice(const void * W, int N, int K, int rpc, int b) {
    Slice s;
    s.r0 = b * rpc;
    s.nr = max(0, min(rpc, N - s.r0));
    s.bytes = (uint32_t)s.nr * K * 2u;
    s.src = reinterpret_cast<const unsigned char *>(W) + (size_t)s.r0 * K * 2u;
    return s;
}

// slice j of a frame: 4 per decoder layer (qkv, o, ff1, ff2) then the 8 out-projections of the local transformer
__device__ __forceinline__ Slice frame_slice(const FrameLoopParams & p, int j, int b) {
    const int nl = p.L * kSlicesPerFrameLayer;
    if (j < nl) {
        const LoopLayer & L = p.layer[j >> 2];
        switch (j & 3) {
            case 0: return make_slice(L.qkv, 3 * D, D, RQ, b);
            case 1: return make_slice(L.o, D, D, RO, b);
            case 2: return make_slice(L.ff1, F, D, RF1, b);
            default: return make_slice(L.ff2, D, F, RF2, b);
        }
    }
    return make_slice(p.lt_out_w[j - nl], p.V, LD, ROUT, b);
}

// ---- weight prefetch warp (one lane) ----------------------------------------------------------------
__device__ void prefetch_lane(LoopSmem & S, const FrameLoopParams & p, int b) {
    // resident local-transformer slices
    {
        const Slice a = make_slice(p.lt_in_w, LD, D, RIN, b), q = make_slice(p.lt_qkv, 3 * LD, LD, RLQ, b),
                    o = make_slice(p.lt_o, LD, LD, RLO, b), f1 = make_slice(p.lt_ff1, LF, LD, RLF1, b),
                    f2 = make_slice(p.lt_ff2, LD, LF, RLF2, b);
        mbar_expect_tx(&S.res_bar, a.bytes + q.bytes + o.bytes + f1.bytes + f2.bytes);
        if (a.bytes) bulk_g2s(S.w_in, a.src, a.bytes, &S.res_bar);
        if (q.bytes) bulk_g2s(S.w_qkv, q.src, q.bytes, &S.res_bar);
        if (o.bytes) bulk_g2s(S.w_o, o.src, o.bytes, &S.res_bar);
        if (f1.bytes) bulk_g2s(S.w_ff1, f1.src, f1.bytes, &S.res_bar);
        if (f2.bytes) bulk_g2s(S.w_ff2, f2.src, f2.bytes, &S.res_bar);
    }
    const uint64_t pol = l2_evict_first_policy();
    const int per_frame = p.L * kSlicesPerFrameLayer + 8;
    const long total = (long)p.n_steps * per_frame;
    int aoff[kQD];
    int tail = 0, k_issued = 0, k_released = 0;
    bool stopped = false;
    int j = 0;
    for (long wp = 0; wp < total && !stopped; wp++, j = (j + 1 == per_frame ? 0 : j + 1)) {
        const int nl = p.L * kSlicesPerFrameLayer;
        const Slice s = frame_slice(p, (p.dbg_flags & 1) ? (j < nl ? (j & 3) : nl) : j, b);
        if (s.bytes == 0) continue;
        const int need = (int)((s.bytes + 127u) & ~127u);
        int off = -1;
        for (;;) {
            if (k_issued - k_released < kQD) {
                if (k_issued == k_released) { off = 0; tail = need; }
                else {
                    const int oldest = aoff[k_released % kQD];
                    if (tail > oldest) {
                        if (tail + need <= kRingBytes) { off = tail; tail += need; }
                        else if (need < oldest) { off = 0; tail = need; }
                    } else if (tail + need < oldest) { off = tail; tail += need; }
                }
            }
            if (off >= 0) break;
            // wait until the compute warps have released the oldest slice
            uint64_t * eb = &S.empty_bar[k_released % kQD];
            const uint32_t par = (uint32_t)(k_released / kQD) & 1u;
            while (!mbar_try(eb, par)) { if (S.stop) { stopped = true; break; } }
            if (stopped) break;
            k_released++;
        }
        if (stopped) break;
        const int slot = k_issued % kQD;
        aoff[slot] = off;
        S.q_off[slot] = off;
        fence_proxy_async();            // the ring bytes were last read through the generic proxy
        mbar_expect_tx(&S.full_bar[slot], s.bytes);
        uint32_t done = 0;
        while (done < s.bytes) {
            const uint32_t n = min(s.bytes - done, 32768u);
            bulk_g2s_hint(S.ring + off + done, s.src + done, n, &S.full_bar[slot], pol);
            done += n;
        }
        k_issued++;
    }
    // early stop (EOS): copies that were issued but never consumed must land before the CTA may exit
    for (int k = k_released; k < k_issued; k++) mbar_wait(&S.full_bar[k % kQD], (uint32_t)(k / kQD) & 1u);
}

// ---- compute-side helpers ---------------------------------------------------------------------------
// The frame is a chain of ~115 short dependent phases, so the kernel is latency bound, and instruction-cache
// misses are latency too: a first version with every phase inlined (190 KB of SASS) ran each phase 3-5x slower
// than the same code in a loop that fits the instruction cache.  Hence the structure below: ONE generic,
// non-inlined `phase` (poll inputs -> LayerNorm -> GEMV -> epilogue -> emit) driven by a small descriptor,
// plus three custom pieces (self-attention, partial combine, folded cross-attention).
#define LOOP_STAMP() do { if (S.dbg_on && threadIdx.x == 0 && S.dbg_i < kLoopDbgStamps) S.dbg[S.dbg_i++] = gtime(); } while (0)

__device__ __forceinline__ const uint4 * xin_of(const LoopSmem & S, int xb) { return S.xbuf + (size_t)(blockIdx.x % kR) * S.rstride + S.xoff[xb]; }
__device__ __forceinline__ uint4 * xout_of(const LoopSmem & S, int xb, int replica) { return S.xbuf + (size_t)replica * S.rstride + S.xoff[xb]; }

__device__ __forceinline__ const bf * ring_wait(LoopSmem & S, int kc) {
    const int slot = kc % kQD;
    mbar_wait(&S.full_bar[slot], (uint32_t)(kc / kQD) & 1u);
    return reinterpret_cast<const bf *>(S.ring + S.q_off[slot]);
}
__device__ __forceinline__ void ring_release(LoopSmem & S, int kc, int lane) {
    __syncwarp();
    if (lane == 0) mbar_arrive(&S.empty_bar[kc % kQD]);
}

// profiling aids (MGB_LOOP_FLAGS): bit 0 = the prefetcher re-reads layer 0's slices (L2 hits instead of HBM),
// bit 1 = polls do not wait (c_poll_mask = 0).  Timing experiments only: the results are garbage.
__constant__ unsigned c_poll_mask = 0xffffffffu;

// poll packets [0, n) of one exchange into dst[0, nfl) (packet i carries floats 3i .. 3i+2)
__device__ __noinline__ void poll_vec(const uint4 * src, int n, unsigned flag, float * dst, int nfl) {
    const int ctid = threadIdx.x;
    const unsigned mask = c_poll_mask;
    for (int i = ctid; i < n; i += 3 * kCT) {
        const int i1 = i + kCT, i2 = i + 2 * kCT;
        const uint4 * q0 = src + i, * q1 = src + i1, * q2 = src + i2;
        uint4 v0 = ld_pkt(q0), v1 = make_uint4(0, 0, 0, flag), v2 = make_uint4(0, 0, 0, flag);
        if (i1 < n) v1 = ld_pkt(q1);
        if (i2 < n) v2 = ld_pkt(q2);
        while ((v0.w ^ flag) & mask) v0 = ld_pkt(q0);
        while ((v1.w ^ flag) & mask) v1 = ld_pkt(q1);
        while ((v2.w ^ flag) & mask) v2 = ld_pkt(q2);
        int g = 3 * i;
        if (g < nfl) dst[g] = __uint_as_float(v0.x);
        if (g + 1 < nfl) dst[g + 1] = __uint_as_float(v0.y);
        if (g + 2 < nfl) dst[g + 2] = __uint_as_float(v0.z);
        if (i1 < n) {
            g = 3 * i1;
            if (g < nfl) dst[g] = __uint_as_float(v1.x);
            if (g + 1 < nfl) dst[g + 1] = __uint_as_float(v1.y);
            if (g + 2 < nfl) dst[g + 2] = __uint_as_float(v1.z);
        }
        if (i2 < n) {
            g = 3 * i2;
            if (g < nfl) dst[g] = __uint_as_float(v2.x);
            if (g + 1 < nfl) dst[g + 1] = __uint_as_float(v2.y);
            if (g + 2 < nfl) dst[g + 2] = __uint_as_float(v2.z);
        }
    }
}

// partial dot products: item = (row, 256-wide k segment); part[row*kPartStride + seg].  Three items per warp pass.
__device__ __noinline__ void gemv_part(const bf * w, int nr, int nseg, const float * x, float * part) {
    const int cw = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int K = nseg * 256;
    const int nitems = nr * nseg;
    for (int it = cw; it < nitems; it += kCW * 3) {
        float acc[3];
        int idx[3];
#pragma unroll
        for (int u = 0; u < 3; u++) {
            const int i = it + u * kCW;
            acc[u] = 0.0f; idx[u] = -1;
            if (i < nitems) {
                const int row = i / nseg, seg = i - row * nseg;
                idx[u] = row * kPartStride + seg;
                const uint4 wv = *reinterpret_cast<const uint4 *>(w + (size_t)row * K + seg * 256 + lane * 8);
                const float4 xa = *reinterpret_cast<const float4 *>(x + seg * 256 + lane * 8);
                const float4 xb = *reinterpret_cast<const float4 *>(x + seg * 256 + lane * 8 + 4);
                float a = bf16lo(wv.x) * xa.x;
                a = fmaf(bf16hi(wv.x), xa.y, a); a = fmaf(bf16lo(wv.y), xa.z, a); a = fmaf(bf16hi(wv.y), xa.w, a);
                a = fmaf(bf16lo(wv.z), xb.x, a); a = fmaf(bf16hi(wv.z), xb.y, a); a = fmaf(bf16lo(wv.w), xb.z, a);
                a = fmaf(bf16hi(wv.w), xb.w, a);
                acc[u] = a;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int u = 0; u < 3; u++) acc[u] += __shfl_xor_sync(0xffffffffu, acc[u], o);
        }
        if (lane == 0) {
#pragma unroll
            for (int u = 0; u < 3; u++) if (idx[u] >= 0) part[idx[u]] = acc[u];
        }
    }
}

// LayerNorm without bias (magpie.cpp:2237-2259): warps 0..nch-1 each compute the statistics of src[0, nch*128)
// and write chunk `cw` of dst = ((src - mean) * rsqrt(var + eps)) * w; wreg = this lane's 4 weights of chunk cw.
__device__ __noinline__ void layer_norm(const float * src, float4 wreg, float * dst, int nch, float eps) {
    const int cw = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (cw < nch) {
        float4 v[6];
        float s = 0.0f;
#pragma unroll
        for (int j = 0; j < 6; j++) {
            v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j < nch) { v[j] = *reinterpret_cast<const float4 *>(src + j * 128 + lane * 4); s += (v[j].x + v[j].y) + (v[j].z + v[j].w); }
        }
        const float inv_n = 1.0f / (float)(nch * 128);
        const float mean = warp_sum(s) * inv_n;
        float q = 0.0f;
#pragma unroll
        for (int j = 0; j < 6; j++)
            if (j < nch) {
                const float a = v[j].x - mean, b2 = v[j].y - mean, c2 = v[j].z - mean, d2 = v[j].w - mean;
                q += (a * a + b2 * b2) + (c2 * c2 + d2 * d2);
            }
        const float var = warp_sum(q) * inv_n;
        const float scale = 1.0f / sqrtf(var + eps);
        float4 m = v[0];
#pragma unroll
        for (int j = 1; j < 6; j++) if (j == cw) m = v[j];
        float4 o;
        o.x = ((m.x - mean) * scale) * wreg.x; o.y = ((m.y - mean) * scale) * wreg.y;
        o.z = ((m.z - mean) * scale) * wreg.z; o.w = ((m.w - mean) * scale) * wreg.w;
        *reinterpret_cast<float4 *>(dst + cw * 128 + lane * 4) = o;
    }
}
__device__ __forceinline__ float4 ln_weight(const float * w, int nch) {
    const int cw = threadIdx.x >> 5, lane = threadIdx.x & 31;
    return cw < nch ? __ldg(reinterpret_cast<const float4 *>(w + cw * 128 + lane * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
}

// one generic phase (descriptor pi); returns true if a ring slice was consumed
__device__ __noinline__ bool phase(LoopSmem & S, int pi, const float * ln_w, bool do_poll, int kc, unsigned seq) {
    const PhaseDesc & a = S.desc[pi];
    const int ctid = threadIdx.x, lane = ctid & 31;
    float4 wr = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ln_w) wr = ln_weight(ln_w, a.nch);
    if (do_poll) {
        poll_vec(xin_of(S, a.xin), a.npk, seq - 1, a.dst, a.nfl);
        cbar();
    }
    LOOP_STAMP();
    if (ln_w) {
        layer_norm(a.ln_src, wr, S.vec, a.nch, S.eps);
        cbar();
    }
    const int rpc = a.rpc, nseg = a.nseg;
    const int r0 = blockIdx.x * rpc, nr = max(0, min(rpc, a.N - r0));
    bool used_ring = false;
    if (nr > 0) {
        const bf * w = a.w;
        if (!w) { w = ring_wait(S, kc); used_ring = true; }
        gemv_part(w, nr, nseg, a.gin, S.part);
        if (used_ring) ring_release(S, kc, lane);
    }
    cbar();
    const int npk = (nr + 2) / 3;
    if (ctid < npk * kR) {
        const int pk = ctid % npk, r = ctid / npk;
        const int epi = a.epi;
        float v[3];
#pragma unroll
        for (int q = 0; q < 3; q++) {
            const int row = 3 * pk + q;
            float s = 0.0f;
            if (row < nr) {
                for (int g = 0; g < nseg; g++) s += S.part[row * kPartStride + g];
                if (epi == EPI_RES) s += a.res[r0 + row];
                else if (epi == EPI_GELU) s = gelu_ggml(s, S.gelu_f16);
                else if (epi == EPI_BIAS) s += __ldg(a.bias + r0 + row);
            }
            v[q] = s;
        }
        st_pkt(xout_of(S, a.xout, r) + r0 / 3 + pk, v[0], v[1], v[2], seq);
    }
    return used_ring;
}

__device__ __forceinline__ bool better(float v, int i, float bv, int bi) { return v > bv || (v == bv && i < bi); }

// argmax over n values of smem `v` by the compute warps; "first max wins" (magpie.cpp:1250-1259)
__device__ __noinline__ int cta_argmax(LoopSmem & S, const float * v, int n) {
    const int ctid = threadIdx.x, cw = ctid >> 5, lane = ctid & 31;
    float bv = -INFINITY; int bi = 0x7fffffff;
    for (int i = ctid; i < n; i += kCT) { const float f = v[i]; if (better(f, i, bv, bi)) { bv = f; bi = i; } }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    cbar();
    if (lane == 0) { S.red[cw] = bv; S.redi[cw] = bi; }
    cbar();
    bv = S.red[0]; bi = S.redi[0];
    for (int w = 1; w < kCW; w++) if (better(S.red[w], S.redi[w], bv, bi)) { bv = S.red[w]; bi = S.redi[w]; }
    return bi;
}

// sample_top_k (magpie.cpp:1072-1109) over the gathered logits (smem), run identically by every CTA.
// Same algorithm as lt_common.cuh::block_sample_top_k, on the 480 compute threads; srt_v aliases `logits`
// once the selection has been compacted.
__device__ __noinline__ int cta_sample_top_k(LoopSmem & S, float * logits, int V, float temperature, int top_k, float u) {
    using lt::order_key;
    const int tid = threadIdx.x, s_cw = tid >> 5, s_lane = tid & 31;
    int k = top_k < V ? top_k : V;
    if (k < 1) k = 1;
    unsigned prefix = 0, pmask = 0; int want = k;
    for (int pass = 0; pass < 4; pass++) {
        const int shift = 24 - 8 * pass;
        if (tid < 256) S.hist[tid] = 0;
        cbar();
        for (int i = tid; i < V; i += kCT) {
            const unsigned key = order_key(logits[i]);
            if ((key & pmask) == prefix) atomicAdd(&S.hist[(key >> shift) & 255u], 1u);
        }
        cbar();
        if (tid == 0) {
            int acc = 0, bb = 255;
            for (; bb > 0; bb--) { if (acc + (int)S.hist[bb] >= want) break; acc += (int)S.hist[bb]; }
            S.misc[0] = bb; S.misc[1] = want - acc;
        }
        cbar();
        prefix |= (unsigned)S.misc[0] << shift; pmask |= 255u << shift; want = S.misc[1];
        cbar();
    }
    const unsigned thr = prefix;
    if (s_cw == 0) {
        int cnt = 0, eq_taken = 0;
        for (int i0 = 0; i0 < V; i0 += 32) {
            const int i = i0 + s_lane;
            const unsigned key = i < V ? order_key(logits[i]) : 0u;
            const bool gt = i < V && key > thr, eq = i < V && key == thr;
            const unsigned eqm = __ballot_sync(0xffffffffu, eq);
            const int eq_rank = eq_taken + __popc(eqm & ((1u << s_lane) - 1u));
            const bool take = gt || (eq && eq_rank < want);
            const unsigned tm = __ballot_sync(0xffffffffu, take);
            if (take) { const int pp = cnt + __popc(tm & ((1u << s_lane) - 1u)); S.sel_v[pp] = logits[i]; S.sel_i[pp] = (uint16_t)i; }
            cnt += __popc(tm); eq_taken += __popc(eqm);
        }
    }
    cbar();
    float * srt_v = logits;
    // rank by counting -> sorted (value desc, index asc); ranks are kept in srt_i's slot until everybody is done reading
    for (int a = tid; a < k; a += kCT) {
        const float va = S.sel_v[a]; const int ia = S.sel_i[a];
        int r = 0;
        for (int q = 0; q < k; q++) { const float vb = S.sel_v[q]; r += (vb > va || (vb == va && (int)S.sel_i[q] < ia)) ? 1 : 0; }
        S.rank[a] = (uint16_t)r;
    }
    cbar();
    for (int a = tid; a < k; a += kCT) { const int r = S.rank[a]; srt_v[r] = S.sel_v[a]; S.srt_i[r] = S.sel_i[a]; }
    cbar();
    const float mx = srt_v[0];
    for (int a = tid; a < k; a += kCT) S.sel_v[a] = expf((srt_v[a] - mx) / temperature);
    cbar();
    if (tid == 0) {
        float sum = 0.0f;
        for (int a = 0; a < k; a++) sum += S.sel_v[a];
        float cum = 0.0f; int pick = S.srt_i[k - 1];
        for (int a = 0; a < k; a++) { cum += S.sel_v[a] / sum; if (u < cum) { pick = S.srt_i[a]; break; } }
        S.misc[2] = pick;
    }
    cbar();
    const int r = S.misc[2];
    cbar();
    return r;
}

// ---- self-attention partial of item (head h, key split sp): keys [k0, k1) --------------------------------------------
__device__ __noinline__ void attention_item(LoopSmem & S, bf * kcl, bf * vcl, int h, int k0, int k1, int pos, unsigned seq) {
    const int ctid = threadIdx.x, cw = ctid >> 5, lane = ctid & 31;
    const unsigned flag = seq - 1;
    const bool has_new = pos >= k0 && pos < k1;
    // q_h -> smem; for the split that holds the new key, k_h / v_h of this step go straight into the cache (bf16).
    // bar.sync orders these global writes before the loads below (same CTA).
    if (cw < 3 && (cw == 0 || has_new)) {
        const int f0 = cw * D + h * DH;
        const int p0 = f0 / 3, p1 = (f0 + DH - 1) / 3;
        if (lane <= p1 - p0) {
            const uint4 * q = xin_of(S, X_QKV) + p0 + lane;
            uint4 v = ld_pkt(q);
            while ((v.w ^ flag) & c_poll_mask) v = ld_pkt(q);
            const int g = 3 * (p0 + lane) - f0;
            const float vv[3] = {__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z)};
#pragma unroll
            for (int q3 = 0; q3 < 3; q3++) {
                const int gi = g + q3;
                if (gi >= 0 && gi < DH) {
                    if (cw == 0) S.qh[gi] = vv[q3];
                    else (cw == 1 ? kcl : vcl)[(size_t)pos * D + h * DH + gi] = __float2bfloat16_rn(vv[q3]);
                }
            }
        }
    }
    cbar();
    LOOP_STAMP();
    float mx = -INFINITY, lsum = 0.0f, acc0 = 0.0f, acc1 = 0.0f;       // lane owns dims 2*lane, 2*lane+1
    for (int c0 = k0 + cw * 32; c0 < k1; c0 += kCW * 32) {
        const int j = c0 + lane;
        const int cnt = min(32, k1 - c0);
        // issue the value loads of the whole chunk first: they do not depend on the scores
        const bf * vbase = vcl + (size_t)c0 * D + h * DH + lane * 2;
        uint32_t vraw[32];
#pragma unroll
        for (int jj = 0; jj < 32; jj++) vraw[jj] = jj < cnt ? __ldcg(reinterpret_cast<const uint32_t *>(vbase + (size_t)jj * D)) : 0u;
        float s = -INFINITY;
        if (j < k1) {
            const bf * kr = kcl + (size_t)j * D + h * DH;
            uint4 kv[8];
#pragma unroll
            for (int q = 0; q < 8; q++) kv[q] = __ldcg(reinterpret_cast<const uint4 *>(kr) + q);
            float dsum = 0.0f;
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const float4 qa = *reinterpret_cast<const float4 *>(S.qh + q * 8), qb = *reinterpret_cast<const float4 *>(S.qh + q * 8 + 4);
                dsum = fmaf(bf16lo(kv[q].x), qa.x, dsum); dsum = fmaf(bf16hi(kv[q].x), qa.y, dsum);
                dsum = fmaf(bf16lo(kv[q].y), qa.z, dsum); dsum = fmaf(bf16hi(kv[q].y), qa.w, dsum);
                dsum = fmaf(bf16lo(kv[q].z), qb.x, dsum); dsum = fmaf(bf16hi(kv[q].z), qb.y, dsum);
                dsum = fmaf(bf16lo(kv[q].w), qb.z, dsum); dsum = fmaf(bf16hi(kv[q].w), qb.w, dsum);
            }
            s = dsum * 0.125f;                                        // 1/sqrt(64)
        }
        const float mnew = fmaxf(mx, warp_max(s));
        const float corr = expf(mx - mnew);
        const float pj = (j < k1) ? expf(s - mnew) : 0.0f;
        lsum = lsum * corr + warp_sum(pj);
        acc0 *= corr; acc1 *= corr;
#pragma unroll
        for (int jj = 0; jj < 32; jj++) {
            const float pb = __shfl_sync(0xffffffffu, pj, jj);
            acc0 = fmaf(pb, bf16lo(vraw[jj]), acc0); acc1 = fmaf(pb, bf16hi(vraw[jj]), acc1);
        }
        mx = mnew;
    }
    if (lane == 0) { S.am[cw] = mx; S.al[cw] = lsum; }
    S.aacc[cw * 64 + lane * 2] = acc0; S.aacc[cw * 64 + lane * 2 + 1] = acc1;
    cbar();
    if (ctid < 64) {
        float M = S.am[0];
        for (int w = 1; w < kCW; w++) M = fmaxf(M, S.am[w]);
        float Ls = 0.0f, o = 0.0f;
        for (int w = 0; w < kCW; w++) {
            const float fct = (S.am[w] == -INFINITY) ? 0.0f : expf(S.am[w] - M);
            Ls += fct * S.al[w]; o += fct * S.aacc[w * 64 + ctid];
        }
        S.aout[2 + ctid] = o;
        if (ctid == 0) { S.aout[0] = M; S.aout[1] = Ls; }
    }
    cbar();
    if (ctid < 22 * kR) {
        const int pk = ctid % 22, r = ctid / 22;
        st_pkt(xout_of(S, X_ATT, r) + blockIdx.x * 22 + pk, S.aout[3 * pk], S.aout[3 * pk + 1], S.aout[3 * pk + 2], seq);
    }
}

// ---- combine the attention partials of all (head, split) items -> S.av ------------------------------------------------
__device__ __noinline__ void attention_combine(LoopSmem & S, int S_split, unsigned seq) {
    poll_vec(xin_of(S, X_ATT), H * S_split * 22, seq - 1, S.vec, H * S_split * 66);
    cbar();
    LOOP_STAMP();
    for (int i = threadIdx.x; i < D; i += kCT) {
        const int hh = i / DH, dd = i % DH;
        const float * ph = S.vec + hh * S_split * 66;
        float M = -INFINITY;
        for (int s2 = 0; s2 < S_split; s2++) M = fmaxf(M, ph[s2 * 66]);
        float Ls = 0.0f, o = 0.0f;
        for (int s2 = 0; s2 < S_split; s2++) {
            const float m2 = ph[s2 * 66];
            const float fct = (m2 == -INFINITY) ? 0.0f : expf(m2 - M);
            Ls += fct * ph[s2 * 66 + 1]; o += fct * ph[s2 * 66 + 2 + dd];
        }
        S.av[i] = o * (1.0f / Ls);
    }
    cbar();
}

// ---- folded cross-attention: x += softmax(M_l LN(x)) N_l  (rows of this CTA) -----------------------------------------------
__device__ __noinline__ void cross_attention(LoopSmem & S, const float * xm, const float * xn, const float * n_xq, unsigned seq) {
    const int ctid = threadIdx.x, cw = ctid >> 5, lane = ctid & 31, E = S.E;
    const float4 wr = ln_weight(n_xq, D / 128);
    const int r0 = blockIdx.x * RO, nr = max(0, min(RO, D - r0));
    // the tables are activation independent: fetch this warp's rows before waiting for x
    const int j0 = cw, j1 = cw + kCW;
    float4 m0[6], m1[6];
#pragma unroll
    for (int q = 0; q < 6; q++) {
        m0[q] = j0 < E ? __ldg(reinterpret_cast<const float4 *>(xm + (size_t)j0 * D + q * 128 + lane * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
        m1[q] = j1 < E ? __ldg(reinterpret_cast<const float4 *>(xm + (size_t)j1 * D + q * 128 + lane * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float nval = 0.0f;
    if (cw < nr && lane < E) nval = __ldg(xn + (size_t)lane * D + r0 + cw);
    poll_vec(xin_of(S, X_XA), D / 3, seq - 1, S.xs, D);
    cbar();
    LOOP_STAMP();
    layer_norm(S.xs, wr, S.vec, D / 128, S.eps);
    cbar();
    float d0 = 0.0f, d1 = 0.0f;
#pragma unroll
    for (int q = 0; q < 6; q++) {
        const float4 xv = *reinterpret_cast<const float4 *>(S.vec + q * 128 + lane * 4);
        d0 = fmaf(m0[q].x, xv.x, d0); d0 = fmaf(m0[q].y, xv.y, d0); d0 = fmaf(m0[q].z, xv.z, d0); d0 = fmaf(m0[q].w, xv.w, d0);
        d1 = fmaf(m1[q].x, xv.x, d1); d1 = fmaf(m1[q].y, xv.y, d1); d1 = fmaf(m1[q].z, xv.z, d1); d1 = fmaf(m1[q].w, xv.w, d1);
    }
    d0 = warp_sum(d0); d1 = warp_sum(d1);
    if (lane == 0) { if (j0 < E) S.sc[j0] = d0; if (j1 < E) S.sc[j1] = d1; }
    cbar();
    if (cw < nr) {
        const float sj = lane < E ? S.sc[lane] : -INFINITY;
        const float mxs = warp_max(sj);
        const float e = lane < E ? expf(sj - mxs) : 0.0f;
        const float sum = warp_sum(e);
        const float o = warp_sum(e * nval);
        if (lane == 0) S.outv[cw] = o * (1.0f / sum) + S.xs[r0 + cw];
    }
    cbar();
    const int npk = (nr + 2) / 3;
    if (ctid < npk * kR) {
        const int pk = ctid % npk, r = ctid / npk;
        st_pkt(xout_of(S, X_XB, r) + r0 / 3 + pk, S.outv[3 * pk], S.outv[3 * pk + 1], S.outv[3 * pk + 2], seq);
    }
}

__global__ void __launch_bounds__(kThreads, 1) frame_loop_kernel(const FrameLoopParams p) {
    extern __shared__ __align__(128) unsigned char loop_smem[];
    LoopSmem & S = *reinterpret_cast<LoopSmem *>(loop_smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.x;

    if (tid == 0) {
        for (int i = 0; i < kQD; i++) { mbar_init(&S.full_bar[i], 1); mbar_init(&S.empty_bar[i], kCW); S.q_off[i] = 0; }
        mbar_init(&S.res_bar, 1);
        S.stop = 0; S.dbg_on = 0; S.dbg_i = 0;
        S.xbuf = p.xbuf; S.rstride = (size_t)p.xoff[X_COUNT];
        for (int i = 0; i <= X_COUNT; i++) S.xoff[i] = p.xoff[i];
        S.eps = p.eps; S.gelu_f16 = p.gelu_f16; S.E = p.E; S.dbg = p.dbg;
        //                     xin     npk           dst     nfl  ln_src  nch       gin     w         N       rpc   nseg      epi       res    bias       xout
        S.desc[PH_QKV]    = {X_XC,   D / 3,        S.xs,   D,   S.xs,   D / 128,  S.vec,  nullptr,  3 * D,  RQ,   D / 256,  EPI_NONE, nullptr, nullptr,  X_QKV};
        S.desc[PH_O]      = {-1,     0,            nullptr, 0,  nullptr, 0,       S.av,   nullptr,  D,      RO,   D / 256,  EPI_RES,  S.xs,  nullptr,    X_XA};
        S.desc[PH_FF1]    = {X_XB,   D / 3,        S.xs,   D,   S.xs,   D / 128,  S.vec,  nullptr,  F,      RF1,  D / 256,  EPI_GELU, nullptr, nullptr,  X_H};
        S.desc[PH_FF2]    = {X_H,    F / 3,        S.vec,  F,   nullptr, 0,       S.vec,  nullptr,  D,      RF2,  F / 256,  EPI_RES,  S.xs,  nullptr,    X_XC};
        S.desc[PH_LT_IN]  = {-1,     0,            nullptr, 0,  nullptr, 0,       S.av,   S.w_in,   LD,     RIN,  D / 256,  EPI_BIAS, nullptr, p.lt_in_b, T_SEQ0};
        S.desc[PH_LT_QKV] = {-1,     0,            nullptr, 0,  S.lx,   LD / 128, S.vec,  S.w_qkv,  3 * LD, RLQ,  1,        EPI_NONE, nullptr, nullptr,  T_QKV};
        S.desc[PH_LT_O]   = {-1,     0,            nullptr, 0,  nullptr, 0,       S.latt, S.w_o,    LD,     RLO,  1,        EPI_RES,  S.lx,  nullptr,    T_X1};
        S.desc[PH_LT_FF1] = {T_X1,   (LD + 2) / 3, S.lx1,  LD,  S.lx1,  LD / 128, S.vec,  S.w_ff1,  LF,     RLF1, 1,        EPI_GELU, nullptr, nullptr,  T_H};
        S.desc[PH_LT_FF2] = {T_H,    (LF + 2) / 3, S.vec,  LF,  nullptr, 0,       S.vec,  S.w_ff2,  LD,     RLF2, LF / 256, EPI_RES,  S.lx1, nullptr,    T_HOUT};
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == kCW) {
        if (lane == 0) prefetch_lane(S, p, b);
        return;
    }

    const int ctid = tid, cw = warp;
    int kc = 0;                              // consumed non-empty ring slices
    unsigned seq = *p.seq;                   // sequence number (flag) of the NEXT exchange
    const int V = p.V, L = p.L;
    const float lt_scale = 1.0f / sqrtf((float)LD);

    for (int i = ctid; i < 8 * LD; i += kCT) S.ltpos[i] = p.lt_pos[i];

    // decoder-input embedding accumulators: x0 = (sum_cb E_cb[code_cb]) / 8 + pos_emb   (magpie.cpp:2746-2787)
    const int e0 = ctid, e1 = ctid + kCT;
    float emb0 = 0.0f, emb1 = 0.0f;
    for (int cb = 0; cb < 8; cb++) {
        const int code = p.codes_io[cb];
        const float a = p.audio_emb[cb][(size_t)code * D + e0];
        const float a1 = e1 < D ? p.audio_emb[cb][(size_t)code * D + e1] : 0.0f;
        emb0 = cb == 0 ? a : emb0 + a; emb1 = cb == 0 ? a1 : emb1 + a1;
    }
    float pe0 = p.dec_pos[(size_t)p.pos0 * D + e0], pe1 = e1 < D ? p.dec_pos[(size_t)p.pos0 * D + e1] : 0.0f;
    int eos_step = -1, frames = 0;

    for (int t = 0; t < p.n_steps; t++) {
        const int pos = p.pos0 + t, nk = pos + 1;
        const size_t row = (size_t)p.row0 + t;
        if (ctid == 0) S.dbg_on = (p.dbg != nullptr && b == 0 && t == p.n_steps - 1) ? 1 : 0;
        S.xs[e0] = emb0 * 0.125f + pe0;
        if (e1 < D) S.xs[e1] = emb1 * 0.125f + pe1;
        cbar();
        LOOP_STAMP();
        if (t + 1 < p.n_steps) {            // next frame's position row: fetched a whole frame ahead
            pe0 = __ldg(p.dec_pos + (size_t)(pos + 1) * D + e0);
            if (e1 < D) pe1 = __ldg(p.dec_pos + (size_t)(pos + 1) * D + e1);
        }

        const int S_split = min(kMaxSplit, max(1, (nk + 127) / 128));
#pragma unroll 1
        for (int l = 0; l < L; l++) {
            const LoopLayer & Ly = p.layer[l];
            // ---- P1: LN -> QKV ---------------------------------------------------------------------------
            if (phase(S, PH_QKV, Ly.n_self, l > 0, kc, seq)) kc++;
            seq++;
            LOOP_STAMP();
            // ---- P2: attention partials, item (head, key split) --------------------------------------------
            if (b < H * S_split) {
                const int h = b / S_split, sp = b % S_split;
                const int per = (nk + S_split - 1) / S_split;
                attention_item(S, (bf *)p.kcache + (size_t)l * p.kv_layer_stride, (bf *)p.vcache + (size_t)l * p.kv_layer_stride,
                               h, sp * per, min(nk, sp * per + per), pos, seq);
            }
            seq++;
            LOOP_STAMP();
            // ---- P3: combine partials -> attention output; O projection + residual -------------------------
            attention_combine(S, S_split, seq);
            if (phase(S, PH_O, nullptr, false, kc, seq)) kc++;
            seq++;
            LOOP_STAMP();
            // ---- P4: LN -> folded cross-attention + residual ---------------------------------------------------
            cross_attention(S, Ly.xm, Ly.xn, Ly.n_xq, seq);
            seq++;
            LOOP_STAMP();
            // ---- P5: LN -> FFN1 -> GELU ------------------------------------------------------------------------
            if (phase(S, PH_FF1, Ly.n_ff, true, kc, seq)) kc++;
            seq++;
            LOOP_STAMP();
            // ---- P6: FFN2 + residual ---------------------------------------------------------------------------
            if (phase(S, PH_FF2, nullptr, true, kc, seq)) kc++;
            seq++;
            LOOP_STAMP();
        }

        // ---- final LayerNorm -> hidden (every CTA holds it) ------------------------------------------------------
        {
            const float4 wr = ln_weight(p.norm_out, D / 128);
            poll_vec(xin_of(S, X_XC), D / 3, seq - 1, S.xs, D);
            cbar();
            LOOP_STAMP();
            layer_norm(S.xs, wr, S.av, D / 128, S.eps);
            cbar();
            if (b == 0) {
                for (int i = ctid; i < D; i += kCT) {
                    const float hv = S.av[i];
                    if (p.hidden_hist) p.hidden_hist[row * D + i] = hv;
                    if (p.hidden_last) p.hidden_last[i] = hv;
                }
            }
        }
        LOOP_STAMP();

        // ================= local transformer (magpie.cpp:1113-1317) =================================================
        if (t == 0) mbar_wait(&S.res_bar, 0);
        const bool forbid_eos = (p.step0 + t) < p.min_frames;
        const bool sampling = p.temperature >= 0.01f;
        bool hit_eos = false;
        phase(S, PH_LT_IN, nullptr, false, kc, seq);          // seq[0] = in_proj . hidden + b
        seq++;
        float seq_next = 0.0f;              // feedback row element `ctid` (ctid < LD) for the next codebook
#pragma unroll 1
        for (int cb = 0; cb < 8; cb++) {
            float * qkv = S.lqkv[cb];       // [q | k | v] of position cb
            // ---- A: x = seq + pos; LN -> QKV --------------------------------------------------------------------
            if (cb == 0) {
                poll_vec(xin_of(S, T_SEQ0), (LD + 2) / 3, seq - 1, S.lx, LD);
                cbar();
                if (ctid < LD) S.lx[ctid] += S.ltpos[ctid];
            } else if (ctid < LD) S.lx[ctid] = seq_next + S.ltpos[cb * LD + ctid];
            cbar();
            phase(S, PH_LT_QKV, p.lt_norm_self, false, kc, seq);
            seq++;
            LOOP_STAMP();
            // ---- B: attention over the <= 8 positions (redundant per CTA); O + residual ------------------------
            poll_vec(xin_of(S, T_QKV), LD, seq - 1, qkv, 3 * LD);
            cbar();
            if (cw <= cb) {
                float s = 0.0f;
#pragma unroll
                for (int q = 0; q < 8; q++) s = fmaf(S.lqkv[cw][LD + lane + 32 * q], qkv[lane + 32 * q], s);
                s = warp_sum(s);
                if (lane == 0) S.sc[cw] = s * lt_scale;
            }
            cbar();
            if (ctid < LD) {
                float mxs = S.sc[0];
                for (int j = 1; j <= cb; j++) mxs = fmaxf(mxs, S.sc[j]);
                float sum = 0.0f, o = 0.0f;
                for (int j = 0; j <= cb; j++) { const float e = expf(S.sc[j] - mxs); sum += e; o = fmaf(e, S.lqkv[j][2 * LD + ctid], o); }
                S.latt[ctid] = o * (1.0f / sum);
            }
            cbar();
            phase(S, PH_LT_O, nullptr, false, kc, seq);
            seq++;
            LOOP_STAMP();
            // ---- C: LN -> FFN1 -> GELU -----------------------------------------------------------------------------
            phase(S, PH_LT_FF1, p.lt_norm_ff, true, kc, seq);
            seq++;
            LOOP_STAMP();
            // ---- D: FFN2 + residual ---------------------------------------------------------------------------------
            phase(S, PH_LT_FF2, nullptr, true, kc, seq);
            seq++;
            LOOP_STAMP();
            // ---- E: output projection of codebook cb (+bias, forbidden-token mask); local argmax ---------------------
            {
                poll_vec(xin_of(S, T_HOUT), (LD + 2) / 3, seq - 1, S.lhout, LD);
                cbar();
                LOOP_STAMP();
                const int r0 = b * ROUT, nr = max(0, min(ROUT, V - r0));
                if (nr > 0) {
                    const bf * w = ring_wait(S, kc);
                    gemv_part(w, nr, 1, S.lhout, S.part);
                    ring_release(S, kc, lane);
                    kc++;
                }
                cbar();
                if (ctid < nr) {
                    const int n = r0 + ctid;
                    float v = S.part[ctid * kPartStride] + __ldg(p.lt_out_b[cb] + n);
                    const bool masked = n == p.bos_id || (n >= p.bos_id + 2 && n <= p.bos_id + 7) || (forbid_eos && n == p.eos_id);
                    if (masked) v = -INFINITY;                                    // magpie.cpp:1131-1145, 1243-1248
                    S.outv[ctid] = v;
                    if (p.logits) p.logits[(row * 8 + cb) * V + n] = v;
                }
                cbar();
                if (sampling) {
                    const int npk = (nr + 2) / 3;
                    if (ctid < npk * kR) {
                        const int pk = ctid % npk, r = ctid / npk;
                        const float v0 = S.outv[3 * pk], v1 = 3 * pk + 1 < nr ? S.outv[3 * pk + 1] : 0.0f, v2 = 3 * pk + 2 < nr ? S.outv[3 * pk + 2] : 0.0f;
                        st_pkt(xout_of(S, T_LOGITS, r) + r0 / 3 + pk, v0, v1, v2, seq);
                    }
                } else if (nr > 0 && ctid < kR) {
                    float bv = S.outv[0]; int bi = 0;
                    for (int r = 1; r < nr; r++) if (S.outv[r] > bv) { bv = S.outv[r]; bi = r; }
                    st_pkt(xout_of(S, T_AMAX, ctid) + b, bv, __int_as_float(r0 + bi), 0.0f, seq);
                }
                seq++;
            }
            LOOP_STAMP();
            // ---- F: global argmax / top-k sample; feedback ---------------------------------------------------------------
            int am, pick;
            if (!sampling) {
                const int nprod = (V + ROUT - 1) / ROUT;
                float bv = -INFINITY; int bi = 0x7fffffff;
                if (ctid < nprod) {
                    const uint4 * q = xin_of(S, T_AMAX) + ctid;
                    uint4 v = ld_pkt(q);
                    while ((v.w ^ (seq - 1)) & c_poll_mask) v = ld_pkt(q);
                    bv = __uint_as_float(v.x); bi = (int)v.y;
                    if (c_poll_mask == 0u) bi = min(max(bi, 0), V - 1);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float ov = __shfl_xor_sync(0xffffffffu, bv, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
                }
                if (lane == 0) { S.red[cw] = bv; S.redi[cw] = bi; }
                cbar();
                LOOP_STAMP();
                bv = S.red[0]; bi = S.redi[0];
                for (int w = 1; w < kCW; w++) if (better(S.red[w], S.redi[w], bv, bi)) { bv = S.red[w]; bi = S.redi[w]; }
                cbar();
                am = bi; pick = bi;
            } else {
                poll_vec(xin_of(S, T_LOGITS), (V + 2) / 3, seq - 1, S.vec, V);
                cbar();
                LOOP_STAMP();
                am = cta_argmax(S, S.vec, V);
                float u;
                if (p.uniforms) u = p.uniforms[row * 8 + cb];
                else {
                    uint32_t r4[4];
                    lt::philox4x32_10((uint32_t)(p.step0 + t), 0u, (uint32_t)cb, 0u, (uint32_t)p.seed, (uint32_t)(p.seed >> 32), r4);
                    u = (float)(r4[0] >> 8) * (1.0f / 16777216.0f);
                }
                pick = cta_sample_top_k(S, S.vec, V, p.temperature, p.top_k, u);
            }
            hit_eos = hit_eos || pick == p.eos_id || am == p.eos_id;
            const int fed = p.forced ? p.forced[row * 8 + cb] : pick;
            if (b == 0 && ctid == 0) { p.argmax[row * 8 + cb] = am; p.sampled[row * 8 + cb] = pick; p.result[2 + cb] = fed; }
            // feedback: seq[cb+1] = row `fed` of P_cb (no 1/8 scale, magpie.cpp:1285-1291); next frame's embedding
            if (cb < 7 && ctid < LD) seq_next = __ldg(p.lt_in_table[cb] + (size_t)fed * LD + ctid);
            {
                const float a = __ldg(p.audio_emb[cb] + (size_t)fed * D + e0);
                const float a1 = e1 < D ? __ldg(p.audio_emb[cb] + (size_t)fed * D + e1) : 0.0f;
                emb0 = cb == 0 ? a : emb0 + a; emb1 = cb == 0 ? a1 : emb1 + a1;
            }
            LOOP_STAMP();
        }
        frames = t + 1;
        if (hit_eos && eos_step < 0) eos_step = t;
        if (hit_eos && !p.teacher && !p.ignore_eos) break;         // magpie.cpp:4341-4352
    }

    cbar();
    if (ctid == 0) S.stop = 1;
    if (b == 0 && ctid == 0) {
        p.result[0] = frames; p.result[1] = eos_step;
        *p.seq = seq;
    }
}

// ---- folded cross-attention tables -----------------------------------------------------------------------------------
__global__ void xattn_fold_kernel(const bf * xk, const bf * xv, const bf * wq, const bf * wo, int d, int dxa, float scale,
                                  float * xm, float * xn) {
    const int j = blockIdx.x;                 // text token
    extern __shared__ float kv[];             // [2][dxa]
    for (int i = threadIdx.x; i < dxa; i += blockDim.x) {
        kv[i] = __bfloat162float(xk[(size_t)j * dxa + i]);
        kv[dxa + i] = __bfloat162float(xv[(size_t)j * dxa + i]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < d; i += blockDim.x) {
        float m = 0.0f, n = 0.0f;
        for (int cc = 0; cc < dxa; cc++) {
            m = fmaf(kv[cc], __bfloat162float(wq[(size_t)cc * d + i]), m);
            n = fmaf(kv[dxa + cc], __bfloat162float(wo[(size_t)i * dxa + cc]), n);
        }
        xm[(size_t)j * d + i] = m * scale;
        xn[(size_t)j * d + i] = n;
    }
}

}  // namespace

bool frame_loop_shape_ok(int d, int f, int h, int ld, int lf, int V, int L) {
    return d == D && f == F && h == H && ld == LD && lf == LF && V > 0 && V <= 2048 && L <= kLoopMaxLayers;
}

void frame_loop_xchg_layout(int V, int * xoff) {
    const int n[X_COUNT] = {3 * D / 3, H * kMaxSplit * 22, D / 3, D / 3, F / 3, D / 3, (LD + 2) / 3, LD, (LD + 2) / 3, (LF + 2) / 3,
                            (LD + 2) / 3, 160, (V + 2) / 3 + 8};
    int o = 0;
    for (int i = 0; i < X_COUNT; i++) { xoff[i] = o; o += (n[i] + 7) & ~7; }     // 128-byte aligned starts
    xoff[X_COUNT] = o;
}
size_t frame_loop_xchg_bytes(const int * xoff) { return (size_t)xoff[X_COUNT] * kR * sizeof(uint4); }

static size_t loop_smem_bytes() { return sizeof(LoopSmem) + 128; }

int frame_loop_max_grid() {
    int dev = 0, sms = 0, per_sm = 0, coop = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev) != cudaSuccess || !coop) return 0;
    if (cudaFuncSetAttribute(frame_loop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)loop_smem_bytes()) != cudaSuccess) { cudaGetLastError(); return 0; }
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, frame_loop_kernel, kThreads, loop_smem_bytes()) != cudaSuccess) { cudaGetLastError(); return 0; }
    return (per_sm >= 1 && sms >= 147) ? sms : 0;
}

bool launch_frame_loop(const FrameLoopParams & p, int grid, cudaStream_t stream) {
    FrameLoopParams pc = p;
    const unsigned mask = (p.dbg_flags & 2) ? 0u : 0xffffffffu;
    MGB_CUDA_TRY(cudaMemcpyToSymbolAsync(c_poll_mask, &mask, sizeof(mask), 0, cudaMemcpyHostToDevice, stream));
    void * args[] = {(void *)&pc};
    MGB_CUDA_TRY(cudaFuncSetAttribute(frame_loop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)loop_smem_bytes()));
    MGB_CUDA_TRY(cudaLaunchCooperativeKernel((void *)frame_loop_kernel, dim3(grid), dim3(kThreads), args, loop_smem_bytes(), stream));
    MGB_LAUNCH_CHECK();
    return true;
}

bool launch_xattn_fold(const void * xk, const void * xv, const void * wq, const void * wo, int E, int d, int dxa, float scale,
                       float * xm, float * xn, cudaStream_t stream) {
    xattn_fold_kernel<<<E, 256, 2 * dxa * sizeof(float), stream>>>((const bf *)xk, (const bf *)xv, (const bf *)wq, (const bf *)wo, d, dxa,
                                                                   scale, xm, xn);
    MGB_LAUNCH_CHECK();
    return true;
}

}  // namespace mgb
