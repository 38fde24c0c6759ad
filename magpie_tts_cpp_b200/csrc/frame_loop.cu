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
//   * cross-attention (1 head over E <= kLoopMaxCtx text tokens) is folded at prefill: M_l = scale K_l Wq_l and
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
constexpr int kAmaxSlots = 160;          // argmax packets of one codebook (one per producing CTA)
constexpr int D = 768, F = 3072, H = 12, DH = 64, LD = 256, LF = 1024;
constexpr int kMaxSplit = 6;
// rows per CTA (multiples of 3 = one packet)
constexpr int RQ = 18, RO = 6, RF1 = 21, RF2 = 6, RIN = 3, RLQ = 9, RLF1 = 9, RLF2 = 3, ROUT = 15;
constexpr int LQ = 4 * LD;           // local transformer: q | k | hi(Wo Wv n) | lo(Wo Wv n) rows
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

// ---- shared-memory state --------------------------------------------------------------------------
struct alignas(128) LoopSmem {
    unsigned char ring[kRingBytes];
    // local-transformer layer slices, resident for the whole launch
    bf w_in[RIN * D]; bf w_qkv[RLQ * LD]; bf w_ff1[RLF1 * LD]; bf w_ff2[RLF2 * LF];
    alignas(16) float xs[D];        // residual stream (full vector, refreshed by every x exchange)
    alignas(16) float vec[kVecFloats];   // staged GEMV input / attention partials / gathered logits
    alignas(16) float av[D];        // combined attention output; final hidden
    alignas(16) float part[24 * kPartStride];   // GEMV partial sums [row][k segment]
    alignas(16) float am[16], al[16], aacc[kCW * 64], aout[68];
    alignas(16) float qh[DH], knew[DH], vnew[DH];
    alignas(16) float lx[LD], lx1[LD], lhout[LD];
    alignas(16) float lqkv[8][3 * LD];   // local transformer: [q | k | vo] of every position of the frame (vo = Wo Wv n = hi + lo rows)
    alignas(16) float ltpos[8 * LD];
    alignas(16) float sc[32];
    alignas(16) float outv[24];
    alignas(16) float red[32], red2[32]; int redi[32];
    alignas(16) float sel_v[2048]; uint16_t sel_i[2048], srt_i[2048], rank[2048];
    unsigned hist[256]; int misc[8];
    uint64_t full_bar[kQD], empty_bar[kQD], res_bar;
    int q_off[kQD];
    volatile int stop;
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
        const Slice a = make_slice(p.lt_in_w, LD, D, RIN, b), q = make_slice(p.lt_qkvo, LQ, LD, RLQ, b),
                    f1 = make_slice(p.lt_ff1, LF, LD, RLF1, b), f2 = make_slice(p.lt_ff2, LD, LF, RLF2, b);
        mbar_expect_tx(&S.res_bar, a.bytes + q.bytes + f1.bytes + f2.bytes);
        if (a.bytes) bulk_g2s(S.w_in, a.src, a.bytes, &S.res_bar);
        if (q.bytes) bulk_g2s(S.w_qkv, q.src, q.bytes, &S.res_bar);
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
// A frame is a chain of ~115 short dependent phases, so everything below is written for LATENCY, not throughput:
// what costs time is the number of dependent instructions, shared-memory wavefronts and shuffles on the critical
// path of a phase (measured: 100 dependent instructions ~ 0.25 us).  Hence: one packet per thread with the
// LayerNorm statistics taken from registers, GEMV warps that load their x segment once and reuse it for several
// rows with a transposed shuffle reduction, and an epilogue spread over the lanes of one warp.
struct Ctx {
    int b, ctid, cw, lane;
    int kc;                     // consumed non-empty ring slices
    unsigned seq;               // sequence number (flag) of the NEXT exchange
    size_t rstride;
    const uint4 * xin;          // the replica this CTA polls
    bool dbg_on; int dbg_i;
};

#define LOOP_STAMP() do { if (c.dbg_on && c.ctid == 0 && c.dbg_i < kLoopDbgPerCta) p.dbg[(size_t)c.b * kLoopDbgPerCta + c.dbg_i++] = gtime(); } while (0)

// profiling aids (MGB_LOOP_FLAGS): bit 0 = the prefetcher re-reads layer 0's slices (L2 hits instead of HBM),
// bit 1 = polls do not wait (c_poll_mask = 0).  Timing experiments only: the results are garbage.
__constant__ unsigned c_poll_mask = 0xffffffffu;

__device__ __forceinline__ const bf * ring_wait(LoopSmem & S, const Ctx & c) {
    const int slot = c.kc % kQD;
    mbar_wait(&S.full_bar[slot], (uint32_t)(c.kc / kQD) & 1u);
    return reinterpret_cast<const bf *>(S.ring + S.q_off[slot]);
}
__device__ __forceinline__ void ring_release(LoopSmem & S, Ctx & c) {
    __syncwarp();
    if (c.lane == 0) mbar_arrive(&S.empty_bar[c.kc % kQD]);
    c.kc++;
}

__device__ __forceinline__ uint4 poll_pkt(const uint4 * q, unsigned flag) {
    uint4 r = ld_pkt(q);
    while ((r.w ^ flag) & c_poll_mask) r = ld_pkt(q);
    return r;
}

// poll packets [0, n) of one exchange into dst[0, nfl) (packet i carries floats 3i .. 3i+2); up to 3 packets per thread
__device__ __forceinline__ void poll_vec(const uint4 * src, int n, unsigned flag, float * dst, int nfl, int ctid) {
    const unsigned mask = c_poll_mask;
    for (int i = ctid; i < n; i += 3 * kCT) {
        const int i1 = i + kCT, i2 = i + 2 * kCT;
        const uint4 * q0 = src + i, * q1 = src + i1, * q2 = src + i2;
        uint4 v0 = make_uint4(0, 0, 0, ~flag), v1 = make_uint4(0, 0, 0, i1 < n ? ~flag : flag), v2 = make_uint4(0, 0, 0, i2 < n ? ~flag : flag);
        bool p0 = true, p1 = i1 < n, p2 = i2 < n;
        do {                                    // all pending packets are re-polled together: one round trip per round
            if (p0) v0 = ld_pkt(q0);
            if (p1) v1 = ld_pkt(q1);
            if (p2) v2 = ld_pkt(q2);
            p0 = ((v0.w ^ flag) & mask) != 0; p1 = ((v1.w ^ flag) & mask) != 0; p2 = ((v2.w ^ flag) & mask) != 0;
        } while (p0 || p1 || p2);
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

// ---- full-vector input of a phase: thread i < ceil(N/3) owns floats 3i .. 3i+2 -------------------------------------------
template <int N> __device__ __forceinline__ void ln_weights3(const float * w, int i, float (&w3)[3]) {
#pragma unroll
    for (int q = 0; q < 3; q++) w3[q] = (i < (N + 2) / 3 && 3 * i + q < N) ? __ldg(w + 3 * i + q) : 0.0f;
}
// poll this thread's packet, keep the raw values in `xs` (smem, for residual adds) and in registers
template <int N> __device__ __forceinline__ void load3_poll(const uint4 * src, unsigned flag, float * xs, int i, float (&v)[3]) {
    v[0] = v[1] = v[2] = 0.0f;
    if (i < (N + 2) / 3) {
        const uint4 r = poll_pkt(src + i, flag);
        const float t[3] = {__uint_as_float(r.x), __uint_as_float(r.y), __uint_as_float(r.z)};
#pragma unroll
        for (int q = 0; q < 3; q++) if (3 * i + q < N) { v[q] = t[q]; xs[3 * i + q] = t[q]; }
    }
}
template <int N> __device__ __forceinline__ void load3_smem(const float * xs, int i, float (&v)[3]) {
#pragma unroll
    for (int q = 0; q < 3; q++) v[q] = (i < (N + 2) / 3 && 3 * i + q < N) ? xs[3 * i + q] : 0.0f;
}
// LayerNorm without bias (magpie.cpp:2237-2259) of the vector held 3-per-thread in registers -> out (smem).
// Statistics in one pass (sum, sum of squares): two warp reductions + one barrier.  Ends with a barrier.
template <int N>
__device__ __forceinline__ void ln3(LoopSmem & S, const float (&v)[3], const float (&w3)[3], float * out, float eps, const Ctx & c) {
    float s = (v[0] + v[1]) + v[2], ss = fmaf(v[0], v[0], fmaf(v[1], v[1], v[2] * v[2]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); ss += __shfl_xor_sync(0xffffffffu, ss, o); }
    if (c.lane == 0) { S.red[c.cw] = s; S.red2[c.cw] = ss; }
    cbar();
    const float4 a0 = *reinterpret_cast<const float4 *>(S.red), a1 = *reinterpret_cast<const float4 *>(S.red + 4),
                 a2 = *reinterpret_cast<const float4 *>(S.red + 8), a3 = *reinterpret_cast<const float4 *>(S.red + 12);
    const float4 b0 = *reinterpret_cast<const float4 *>(S.red2), b1 = *reinterpret_cast<const float4 *>(S.red2 + 4),
                 b2 = *reinterpret_cast<const float4 *>(S.red2 + 8), b3 = *reinterpret_cast<const float4 *>(S.red2 + 12);
    const float ts = (((a0.x + a0.y) + (a0.z + a0.w)) + ((a1.x + a1.y) + (a1.z + a1.w))) + (((a2.x + a2.y) + (a2.z + a2.w)) + ((a3.x + a3.y) + a3.z));
    const float tss = (((b0.x + b0.y) + (b0.z + b0.w)) + ((b1.x + b1.y) + (b1.z + b1.w))) + (((b2.x + b2.y) + (b2.z + b2.w)) + ((b3.x + b3.y) + b3.z));
    const float mean = ts * (1.0f / N);
    const float var = fmaxf(tss * (1.0f / N) - mean * mean, 0.0f);
    const float scale = 1.0f / sqrtf(var + eps);
    const int i = c.ctid;
    if (i < (N + 2) / 3) {
#pragma unroll
        for (int q = 0; q < 3; q++) if (3 * i + q < N) out[3 * i + q] = ((v[q] - mean) * scale) * w3[q];
    }
    cbar();
}

// ---- GEMV over this CTA's nr rows: warp -> (256-wide k segment, row group); x segment loaded once per warp ---------------
// Row sums of up to P rows are reduced together ("transposed" butterfly: 9 shuffles for 8 rows instead of 40).
template <int P> __device__ __forceinline__ float reduce_rows(float (&a)[P], int lane, int & r) {
    r = 0;
    if constexpr (P >= 8) {
        const bool up = lane & 16;
#pragma unroll
        for (int j = 0; j < 4; j++) { const float send = up ? a[j] : a[j + 4], keep = up ? a[j + 4] : a[j]; a[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16); }
        r += up ? 4 : 0;
    }
    if constexpr (P >= 4) {
        constexpr int lv = P >= 8 ? 8 : 16;
        const bool up = lane & lv;
#pragma unroll
        for (int j = 0; j < 2; j++) { const float send = up ? a[j] : a[j + 2], keep = up ? a[j + 2] : a[j]; a[j] = keep + __shfl_xor_sync(0xffffffffu, send, lv); }
        r += up ? 2 : 0;
    }
    if constexpr (P >= 2) {
        constexpr int lv = P >= 8 ? 4 : (P >= 4 ? 8 : 16);
        const bool up = lane & lv;
        const float send = up ? a[0] : a[1], keep = up ? a[1] : a[0];
        a[0] = keep + __shfl_xor_sync(0xffffffffu, send, lv);
        r += up ? 1 : 0;
    }
    constexpr int rest = P >= 8 ? 2 : (P >= 4 ? 4 : (P >= 2 ? 8 : 16));
#pragma unroll
    for (int o = rest; o > 0; o >>= 1) a[0] += __shfl_xor_sync(0xffffffffu, a[0], o);
    return a[0];
}

template <int NSEG, int P>      // P = rows per warp, padded to a power of two (1, 2, 4, 8)
__device__ __forceinline__ void gemv_rows(const bf * w, int nr, const float * x, float * part, const Ctx & c) {
    constexpr int K = NSEG * 256, WPS = kCW / NSEG;
    const int seg = c.cw / WPS, g = c.cw - seg * WPS;
    if (seg >= NSEG || g >= nr) return;
    const float4 xa = *reinterpret_cast<const float4 *>(x + seg * 256 + c.lane * 8);
    const float4 xb = *reinterpret_cast<const float4 *>(x + seg * 256 + c.lane * 8 + 4);
    float acc[P];
#pragma unroll
    for (int r = 0; r < P; r++) {
        const int row = g + r * WPS;
        acc[r] = 0.0f;
        if (row < nr) {
            const uint4 wv = *reinterpret_cast<const uint4 *>(w + (size_t)row * K + seg * 256 + c.lane * 8);
            float a = bf16lo(wv.x) * xa.x;
            a = fmaf(bf16hi(wv.x), xa.y, a); a = fmaf(bf16lo(wv.y), xa.z, a); a = fmaf(bf16hi(wv.y), xa.w, a);
            a = fmaf(bf16lo(wv.z), xb.x, a); a = fmaf(bf16hi(wv.z), xb.y, a); a = fmaf(bf16lo(wv.w), xb.z, a);
            a = fmaf(bf16hi(wv.w), xb.w, a);
            acc[r] = a;
        }
    }
    int r;
    const float sum = reduce_rows<P>(acc, c.lane, r);
    const int row = g + r * WPS;
    if ((c.lane & (32 / P - 1)) == 0 && row < nr) part[row * kPartStride + seg] = sum;
}

enum { EPI_NONE = 0, EPI_RES = 1, EPI_GELU = 2, EPI_BIAS = 3 };

// epilogue + emit by warp 0: lane r finishes row r (nr <= 21), packets are assembled with shuffles
template <int NSEG, int EPI>
__device__ __forceinline__ void emit_rows(const LoopSmem & S, const FrameLoopParams & p, const Ctx & c, int xb, int r0, int nr,
                                          const float * res, const float * bias, const float * add2) {
    if (c.cw != 0 || nr <= 0) return;
    float s = 0.0f;
    if (c.lane < nr) {
        s = S.part[c.lane * kPartStride];
#pragma unroll
        for (int g = 1; g < NSEG; g++) s += S.part[c.lane * kPartStride + g];
        if (EPI == EPI_RES) s += res[r0 + c.lane];
        else if (EPI == EPI_GELU) s = gelu_ggml(s, p.gelu_f16);
        else if (EPI == EPI_BIAS) s += __ldg(bias + r0 + c.lane) + add2[r0 + c.lane];
    }
    const int npk = (nr + 2) / 3;
    const int pk = c.lane % npk, rep = c.lane / npk;
    const float v0 = __shfl_sync(0xffffffffu, s, 3 * pk), v1 = __shfl_sync(0xffffffffu, s, 3 * pk + 1), v2 = __shfl_sync(0xffffffffu, s, 3 * pk + 2);
    if (c.lane < npk * kR) st_pkt(p.xbuf + (size_t)rep * c.rstride + p.xoff[xb] + r0 / 3 + pk, v0, v1, v2, c.seq);
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
    cbar();
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
    // (same three changes as lt_common.cuh block_sample_top_k: warp-aggregated histogram atomics, the bin of a radix pass found by a
    //  parallel suffix sum instead of one thread walking 256 bins, compaction by all warps around a prefix over the index chunks)
    unsigned prefix = 0, pmask = 0; int want = k;
    if (tid < 256) S.hist[tid] = 0;
    cbar();
    for (int pass = 0; pass < 4; pass++) {
        const int shift = 24 - 8 * pass;
        for (int i0 = 0; i0 < V; i0 += kCT) {
            const int i = i0 + tid;
            const unsigned key = i < V ? order_key(logits[i]) : 0u;
            const bool valid = i < V && (key & pmask) == prefix;
            const unsigned bin = (key >> shift) & 255u;
            const unsigned vm = __ballot_sync(0xffffffffu, valid);
            if (valid) {
                if (pass == 0) {                     // all V keys, two or three hot bins (sign + exponent byte): one atomic per bin and warp
                    const unsigned peers = __match_any_sync(vm, bin);
                    if (s_lane == __ffs(peers) - 1) atomicAdd(&S.hist[bin], (unsigned)__popc(peers));
                } else atomicAdd(&S.hist[bin], 1u);  // later passes: only the keys of one bin of the previous pass
            }
        }
        cbar();
        unsigned c = 0, sfx = 0;
        if (tid < 256) {
            c = S.hist[tid]; sfx = c;
            S.hist[tid] = 0;                       // (this thread's bin, ready for the next pass: no separate clearing step)
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_down_sync(0xffffffffu, sfx, o); if (s_lane + o < 32) sfx += t; }
        }
        if (tid < 256 && s_lane == 0) reinterpret_cast<unsigned *>(S.sel_v)[s_cw] = sfx;      // (sel_v is free until the compaction)
        cbar();
        if (tid < 256) {
            unsigned off = 0;
            for (int w = s_cw + 1; w < 8; w++) off += reinterpret_cast<const unsigned *>(S.sel_v)[w];
            const unsigned suf = sfx + off;
            if (suf - c < (unsigned)want && (tid == 0 || suf >= (unsigned)want)) { S.misc[0] = tid; S.misc[1] = want - (int)(suf - c); }
        }
        cbar();
        prefix |= (unsigned)S.misc[0] << shift; pmask |= 255u << shift; want = S.misc[1];
        // (no barrier here: the next pass writes misc only after three more barriers, and hist was last read before the previous one)
    }
    cbar();
    const unsigned thr = prefix;
    const int nchunk = (V + 31) >> 5;              // <= 64: hist = [greater | equal | base | equal before], 64 entries each
    for (int ch = s_cw; ch < nchunk; ch += kCW) {
        const int i = ch * 32 + s_lane;
        const unsigned key = i < V ? order_key(logits[i]) : 0u;
        const unsigned gm = __ballot_sync(0xffffffffu, i < V && key > thr), em = __ballot_sync(0xffffffffu, i < V && key == thr);
        if (s_lane == 0) { S.hist[ch] = (unsigned)__popc(gm); S.hist[64 + ch] = (unsigned)__popc(em); }
    }
    cbar();
    if (s_cw == 0) {
        int carry_e = 0, carry_t = 0;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int ch = s_lane + 32 * h;
            const int e = ch < nchunk ? (int)S.hist[64 + ch] : 0, g = ch < nchunk ? (int)S.hist[ch] : 0;
            int pe = e;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, pe, o); if (s_lane >= o) pe += t; }
            const int eqb = carry_e + pe - e;
            const int te = min(max(want - eqb, 0), e), tk = g + te;
            int pt = tk;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, pt, o); if (s_lane >= o) pt += t; }
            const int base = carry_t + pt - tk;
            carry_e += __shfl_sync(0xffffffffu, pe, 31); carry_t += __shfl_sync(0xffffffffu, pt, 31);
            if (ch < nchunk) { S.hist[128 + ch] = (unsigned)base; S.hist[192 + ch] = (unsigned)eqb; }
        }
    }
    cbar();
    for (int ch = s_cw; ch < nchunk; ch += kCW) {
        const int i = ch * 32 + s_lane;
        const unsigned key = i < V ? order_key(logits[i]) : 0u;
        const bool gt = i < V && key > thr, eq = i < V && key == thr;
        const unsigned eqm = __ballot_sync(0xffffffffu, eq);
        const int eq_rank = (int)S.hist[192 + ch] + __popc(eqm & ((1u << s_lane) - 1u));
        const bool take = gt || (eq && eq_rank < want);
        const unsigned tm = __ballot_sync(0xffffffffu, take);
        if (take) { const int pp = (int)S.hist[128 + ch] + __popc(tm & ((1u << s_lane) - 1u)); S.sel_v[pp] = logits[i]; S.sel_i[pp] = (uint16_t)i; }
    }
    cbar();
    float * srt_v = logits;
    // rank by counting -> sorted (value desc, index asc)
    // (k <= 128: element a = tid mod 128, a third of the k comparisons per thread, partial ranks added in shared memory)
    if (k <= 128) {
        if (tid < 128) S.hist[tid] = 0;
        cbar();
        const int a = tid & 127, part = tid >> 7, per = (k + 2) / 3;          // 480 threads: parts 0..2 of 128 elements (part 3 = 96 threads: unused)
        if (a < k && part < 3) {
            const float va = S.sel_v[a]; const int ia = S.sel_i[a];
            int r = 0;
            const int q1 = min(k, (part + 1) * per);
#pragma unroll 4
            for (int q = part * per; q < q1; q++) { const float vb = S.sel_v[q]; r += (vb > va || (vb == va && (int)S.sel_i[q] < ia)) ? 1 : 0; }
            if (r) atomicAdd(&S.hist[a], (unsigned)r);
        }
        cbar();
        if (tid < k) S.rank[tid] = (uint16_t)S.hist[tid];
    } else {
        for (int a = tid; a < k; a += kCT) {
            const float va = S.sel_v[a]; const int ia = S.sel_i[a];
            int r = 0;
            for (int q = 0; q < k; q++) { const float vb = S.sel_v[q]; r += (vb > va || (vb == va && (int)S.sel_i[q] < ia)) ? 1 : 0; }
            S.rank[a] = (uint16_t)r;
        }
    }
    cbar();
    for (int a = tid; a < k; a += kCT) { const int r = S.rank[a]; srt_v[r] = S.sel_v[a]; S.srt_i[r] = S.sel_i[a]; }
    cbar();
    const float mx = srt_v[0];
    float ex[5];
    { int n = 0; for (int a = tid; a < k; a += kCT, n++) ex[n] = expf((srt_v[a] - mx) / temperature); }
    cbar();
    { int n = 0; for (int a = tid; a < k; a += kCT, n++) S.sel_v[a] = ex[n]; }
    cbar();
    if (tid == 0) {                                // sequential float accumulation, as the reference
        float sum = 0.0f;
#pragma unroll 8
        for (int a = 0; a < k; a++) sum += S.sel_v[a];
        S.misc[3] = __float_as_int(sum);
    }
    cbar();
    {
        const float sum = __int_as_float(S.misc[3]);
        for (int a = tid; a < k; a += kCT) srt_v[a] = S.sel_v[a] / sum;        // (the sorted values are no longer needed)
    }
    cbar();
    if (tid == 0) {
        float cum = 0.0f; int pa = k - 1; bool found = false;
#pragma unroll 8
        for (int a = 0; a < k; a++) { cum += srt_v[a]; if (!found && u < cum) { pa = a; found = true; } }
        S.misc[2] = S.srt_i[pa];
    }
    cbar();
    const int r = S.misc[2];
    cbar();
    return r;
}

// ---- self-attention partial of item (head h, key split sp): keys [k0, k1) --------------------------------------------
// Latency structure: the K / V rows of the OLD keys do not depend on this step, so their loads are issued before the
// wait for q; the new key (position pos, last split only) is taken from the exchange, not from the cache, and its
// cache rows are written off the critical path.
__device__ __forceinline__ void attention_item_small(LoopSmem & S, const FrameLoopParams & p, Ctx & c, bf * kcl, bf * vcl,
                                               int h, int k0, int k1, int pos) {
    const int cw = c.cw, lane = c.lane, ctid = c.ctid;
    const unsigned flag = c.seq - 1;
    const bool has_new = pos >= k0 && pos < k1;
    const int ke = has_new ? k1 - 1 : k1;                       // old keys [k0, ke); pos == k1 - 1 when has_new
    const int nw = min(kCW, (ke - k0 + 31) / 32);                // warps holding a chunk (one chunk per warp: per <= 480)
    // 1. old keys: this warp's chunk of 32 keys; loads in flight while q is awaited
    const int c0 = k0 + cw * 32, j = c0 + lane;
    const int cnt = max(0, min(32, ke - c0));
    uint4 kv[8];
    uint32_t vraw[32];
    if (j < ke) {
        const bf * kr = kcl + (size_t)j * D + h * DH;
#pragma unroll
        for (int q = 0; q < 8; q++) kv[q] = __ldcg(reinterpret_cast<const uint4 *>(kr) + q);
    } else {
#pragma unroll
        for (int q = 0; q < 8; q++) kv[q] = make_uint4(0u, 0u, 0u, 0u);
    }
    {
        const bf * vbase = vcl + (size_t)c0 * D + h * DH + lane * 2;
#pragma unroll
        for (int jj = 0; jj < 32; jj++) vraw[jj] = jj < cnt ? __ldcg(reinterpret_cast<const uint32_t *>(vbase + (size_t)jj * D)) : 0u;
    }
    // 2. q_h (all items) and k_h / v_h of this step (last split): warps 0..2 poll <= 23 packets each
    if (cw < 3 && (cw == 0 || has_new)) {
        const int f0 = cw * D + h * DH;
        const int p0 = f0 / 3, p1 = (f0 + DH - 1) / 3;
        if (lane <= p1 - p0) {
            const uint4 v = poll_pkt(c.xin + p.xoff[X_QKV] + p0 + lane, flag);
            const int g = 3 * (p0 + lane) - f0;
            const float vv[3] = {__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z)};
#pragma unroll
            for (int q3 = 0; q3 < 3; q3++) {
                const int gi = g + q3;
                if (gi >= 0 && gi < DH) {
                    if (cw == 0) S.qh[gi] = vv[q3];
                    else {          // the cache holds bf16: use the rounded value now, as later steps will
                        const bf r = __float2bfloat16_rn(vv[q3]);
                        (cw == 1 ? S.knew : S.vnew)[gi] = __bfloat162float(r);
                        (cw == 1 ? kcl : vcl)[(size_t)pos * D + h * DH + gi] = r;
                    }
                }
            }
        }
    }
    cbar();
    LOOP_STAMP();
    float mx = -INFINITY, lsum = 0.0f, acc0 = 0.0f, acc1 = 0.0f;       // lane owns dims 2*lane, 2*lane+1
    if (cw < nw) {
        float s = -INFINITY;
        if (j < ke) {
            float d0 = 0.0f, d1 = 0.0f, d2 = 0.0f, d3 = 0.0f;
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const float4 qa = *reinterpret_cast<const float4 *>(S.qh + q * 8), qb = *reinterpret_cast<const float4 *>(S.qh + q * 8 + 4);
                d0 = fmaf(bf16lo(kv[q].x), qa.x, d0); d1 = fmaf(bf16hi(kv[q].x), qa.y, d1);
                d2 = fmaf(bf16lo(kv[q].y), qa.z, d2); d3 = fmaf(bf16hi(kv[q].y), qa.w, d3);
                d0 = fmaf(bf16lo(kv[q].z), qb.x, d0); d1 = fmaf(bf16hi(kv[q].z), qb.y, d1);
                d2 = fmaf(bf16lo(kv[q].w), qb.z, d2); d3 = fmaf(bf16hi(kv[q].w), qb.w, d3);
            }
            s = ((d0 + d1) + (d2 + d3)) * 0.125f;                     // 1/sqrt(64)
        }
        mx = warp_max(s);
        const float pj = (j < ke) ? expf(s - mx) : 0.0f;
        lsum = warp_sum(pj);
        float a0 = 0.0f, a1 = 0.0f, b0 = 0.0f, b1 = 0.0f;
#pragma unroll
        for (int jj = 0; jj < 32; jj += 2) {
            const float pa = __shfl_sync(0xffffffffu, pj, jj), pb = __shfl_sync(0xffffffffu, pj, jj + 1);
            a0 = fmaf(pa, bf16lo(vraw[jj]), a0); a1 = fmaf(pa, bf16hi(vraw[jj]), a1);
            b0 = fmaf(pb, bf16lo(vraw[jj + 1]), b0); b1 = fmaf(pb, bf16hi(vraw[jj + 1]), b1);
        }
        acc0 = a0 + b0; acc1 = a1 + b1;
    }
    if (has_new && cw == nw % kCW) {          // the new key: one more online-softmax update on a warp of its own if there is one
        const float s = warp_sum(fmaf(S.knew[lane], S.qh[lane], S.knew[lane + 32] * S.qh[lane + 32])) * 0.125f;
        const float mnew = fmaxf(mx, s);
        const float corr = (mx == -INFINITY) ? 0.0f : expf(mx - mnew);
        const float pn = expf(s - mnew);
        lsum = lsum * corr + pn;
        acc0 = fmaf(pn, S.vnew[lane * 2], acc0 * corr); acc1 = fmaf(pn, S.vnew[lane * 2 + 1], acc1 * corr);
        mx = mnew;
    }
    const int nm = has_new ? max(nw, nw % kCW + 1) : nw;        // warps that hold a partial
    if (cw < nm) {
        if (lane == 0) { S.am[cw] = mx; S.al[cw] = lsum; }
        S.aacc[cw * 64 + lane * 2] = acc0; S.aacc[cw * 64 + lane * 2 + 1] = acc1;
    }
    cbar();
    if (ctid < 64) {
        float M = S.am[0];
        for (int w = 1; w < nm; w++) M = fmaxf(M, S.am[w]);
        float Ls = 0.0f, o = 0.0f;
        for (int w = 0; w < nm; w++) {
            const float fct = (S.am[w] == -INFINITY) ? 0.0f : expf(S.am[w] - M);
            Ls += fct * S.al[w]; o += fct * S.aacc[w * 64 + ctid];
        }
        S.aout[2 + ctid] = o;
        if (ctid == 0) { S.aout[0] = M; S.aout[1] = Ls; }
    }
    cbar();
    if (ctid < 22 * kR) {
        const int pk = ctid % 22, r = ctid / 22;
        st_pkt(p.xbuf + (size_t)r * c.rstride + p.xoff[X_ATT] + c.b * 22 + pk, S.aout[3 * pk], S.aout[3 * pk + 1], S.aout[3 * pk + 2], c.seq);
    }
}

// items longer than 480 keys (KV beyond 2 880 cached keys, or fewer key splits): the same partial with the warps scanning further
// 32-key chunks after the wait for q (online-softmax merge per warp); a separate branch, the common case keeps its instruction sequence
__device__ __forceinline__ void attention_item_long(LoopSmem & S, const FrameLoopParams & p, Ctx & c, bf * kcl, bf * vcl,
                                               int h, int k0, int k1, int pos) {
    const int cw = c.cw, lane = c.lane, ctid = c.ctid;
    const unsigned flag = c.seq - 1;
    const bool has_new = pos >= k0 && pos < k1;
    const int ke = has_new ? k1 - 1 : k1;                       // old keys [k0, ke); pos == k1 - 1 when has_new
    const int nch = (ke - k0 + 31) / 32;                         // 32-key chunks of this item; warp cw scans chunks cw, cw + 15, ...
    const int nw = min(kCW, nch);                                // warps holding a partial
    // 1. old keys: this warp's FIRST chunk of 32 keys; loads in flight while q is awaited
    int c0 = k0 + cw * 32, j = c0 + lane;
    int cnt = max(0, min(32, ke - c0));
    uint4 kv[8];
    uint32_t vraw[32];
    if (j < ke) {
        const bf * kr = kcl + (size_t)j * D + h * DH;
#pragma unroll
        for (int q = 0; q < 8; q++) kv[q] = __ldcg(reinterpret_cast<const uint4 *>(kr) + q);
    } else {
#pragma unroll
        for (int q = 0; q < 8; q++) kv[q] = make_uint4(0u, 0u, 0u, 0u);
    }
    {
        const bf * vbase = vcl + (size_t)c0 * D + h * DH + lane * 2;
#pragma unroll
        for (int jj = 0; jj < 32; jj++) vraw[jj] = jj < cnt ? __ldcg(reinterpret_cast<const uint32_t *>(vbase + (size_t)jj * D)) : 0u;
    }
    // 2. q_h (all items) and k_h / v_h of this step (last split): warps 0..2 poll <= 23 packets each
    if (cw < 3 && (cw == 0 || has_new)) {
        const int f0 = cw * D + h * DH;
        const int p0 = f0 / 3, p1 = (f0 + DH - 1) / 3;
        if (lane <= p1 - p0) {
            const uint4 v = poll_pkt(c.xin + p.xoff[X_QKV] + p0 + lane, flag);
            const int g = 3 * (p0 + lane) - f0;
            const float vv[3] = {__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z)};
#pragma unroll
            for (int q3 = 0; q3 < 3; q3++) {
                const int gi = g + q3;
                if (gi >= 0 && gi < DH) {
                    if (cw == 0) S.qh[gi] = vv[q3];
                    else {          // the cache holds bf16: use the rounded value now, as later steps will
                        const bf r = __float2bfloat16_rn(vv[q3]);
                        (cw == 1 ? S.knew : S.vnew)[gi] = __bfloat162float(r);
                        (cw == 1 ? kcl : vcl)[(size_t)pos * D + h * DH + gi] = r;
                    }
                }
            }
        }
    }
    cbar();
    LOOP_STAMP();
    float mx = -INFINITY, lsum = 0.0f, acc0 = 0.0f, acc1 = 0.0f;       // lane owns dims 2*lane, 2*lane+1
    for (int cc = cw; cc < nch; cc += kCW) {
        if (cc != cw) {                 // items longer than 480 keys (KV beyond 2 880): further chunks, loaded after the wait
            c0 = k0 + cc * 32; j = c0 + lane; cnt = max(0, min(32, ke - c0));
            if (j < ke) {
                const bf * kr = kcl + (size_t)j * D + h * DH;
#pragma unroll
                for (int q = 0; q < 8; q++) kv[q] = __ldcg(reinterpret_cast<const uint4 *>(kr) + q);
            }
            const bf * vbase = vcl + (size_t)c0 * D + h * DH + lane * 2;
#pragma unroll
            for (int jj = 0; jj < 32; jj++) vraw[jj] = jj < cnt ? __ldcg(reinterpret_cast<const uint32_t *>(vbase + (size_t)jj * D)) : 0u;
        }
        float s = -INFINITY;
        if (j < ke) {
            float d0 = 0.0f, d1 = 0.0f, d2 = 0.0f, d3 = 0.0f;
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const float4 qa = *reinterpret_cast<const float4 *>(S.qh + q * 8), qb = *reinterpret_cast<const float4 *>(S.qh + q * 8 + 4);
                d0 = fmaf(bf16lo(kv[q].x), qa.x, d0); d1 = fmaf(bf16hi(kv[q].x), qa.y, d1);
                d2 = fmaf(bf16lo(kv[q].y), qa.z, d2); d3 = fmaf(bf16hi(kv[q].y), qa.w, d3);
                d0 = fmaf(bf16lo(kv[q].z), qb.x, d0); d1 = fmaf(bf16hi(kv[q].z), qb.y, d1);
                d2 = fmaf(bf16lo(kv[q].w), qb.z, d2); d3 = fmaf(bf16hi(kv[q].w), qb.w, d3);
            }
            s = ((d0 + d1) + (d2 + d3)) * 0.125f;                     // 1/sqrt(64)
        }
        const float cmx = warp_max(s);
        const float pj = (j < ke) ? expf(s - cmx) : 0.0f;
        const float cl = warp_sum(pj);
        float a0 = 0.0f, a1 = 0.0f, b0 = 0.0f, b1 = 0.0f;
#pragma unroll
        for (int jj = 0; jj < 32; jj += 2) {
            const float pa = __shfl_sync(0xffffffffu, pj, jj), pb = __shfl_sync(0xffffffffu, pj, jj + 1);
            a0 = fmaf(pa, bf16lo(vraw[jj]), a0); a1 = fmaf(pa, bf16hi(vraw[jj]), a1);
            b0 = fmaf(pb, bf16lo(vraw[jj + 1]), b0); b1 = fmaf(pb, bf16hi(vraw[jj + 1]), b1);
        }
        if (cc == cw) { mx = cmx; lsum = cl; acc0 = a0 + b0; acc1 = a1 + b1; }
        else {                          // online-softmax merge of this chunk into the warp's running partial
            const float mnew = fmaxf(mx, cmx);
            const float fo = expf(mx - mnew), fn = expf(cmx - mnew);
            lsum = lsum * fo + cl * fn;
            acc0 = acc0 * fo + (a0 + b0) * fn; acc1 = acc1 * fo + (a1 + b1) * fn;
            mx = mnew;
        }
    }
    if (has_new && cw == nw % kCW) {          // the new key: one more online-softmax update on a warp of its own if there is one
        const float s = warp_sum(fmaf(S.knew[lane], S.qh[lane], S.knew[lane + 32] * S.qh[lane + 32])) * 0.125f;
        const float mnew = fmaxf(mx, s);
        const float corr = (mx == -INFINITY) ? 0.0f : expf(mx - mnew);
        const float pn = expf(s - mnew);
        lsum = lsum * corr + pn;
        acc0 = fmaf(pn, S.vnew[lane * 2], acc0 * corr); acc1 = fmaf(pn, S.vnew[lane * 2 + 1], acc1 * corr);
        mx = mnew;
    }
    const int nm = has_new ? max(nw, nw % kCW + 1) : nw;        // warps that hold a partial
    if (cw < nm) {
        if (lane == 0) { S.am[cw] = mx; S.al[cw] = lsum; }
        S.aacc[cw * 64 + lane * 2] = acc0; S.aacc[cw * 64 + lane * 2 + 1] = acc1;
    }
    cbar();
    if (ctid < 64) {
        float M = S.am[0];
        for (int w = 1; w < nm; w++) M = fmaxf(M, S.am[w]);
        float Ls = 0.0f, o = 0.0f;
        for (int w = 0; w < nm; w++) {
            const float fct = (S.am[w] == -INFINITY) ? 0.0f : expf(S.am[w] - M);
            Ls += fct * S.al[w]; o += fct * S.aacc[w * 64 + ctid];
        }
        S.aout[2 + ctid] = o;
        if (ctid == 0) { S.aout[0] = M; S.aout[1] = Ls; }
    }
    cbar();
    if (ctid < 22 * kR) {
        const int pk = ctid % 22, r = ctid / 22;
        st_pkt(p.xbuf + (size_t)r * c.rstride + p.xoff[X_ATT] + c.b * 22 + pk, S.aout[3 * pk], S.aout[3 * pk + 1], S.aout[3 * pk + 2], c.seq);
    }
}

__device__ __forceinline__ void attention_item(LoopSmem & S, const FrameLoopParams & p, Ctx & c, bf * kcl, bf * vcl,
                                               int h, int k0, int k1, int pos) {
    const int ke = (pos >= k0 && pos < k1) ? k1 - 1 : k1;
    if (ke - k0 <= 32 * kCW) attention_item_small(S, p, c, kcl, vcl, h, k0, k1, pos);
    else attention_item_long(S, p, c, kcl, vcl, h, k0, k1, pos);
}

// ---- combine the attention partials of all (head, split) items -> S.av ------------------------------------------------
__device__ __forceinline__ void attention_combine(LoopSmem & S, const FrameLoopParams & p, Ctx & c, int S_split) {
    poll_vec(c.xin + p.xoff[X_ATT], H * S_split * 22, c.seq - 1, S.vec, H * S_split * 66, c.ctid);
    cbar();
    LOOP_STAMP();
    for (int i = c.ctid; i < D; i += kCT) {
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
// Text tokens are handled 30 at a time (two table rows per warp).  The first 30 rows are activation independent and are
// fetched BEFORE the wait for x (all of a "Hello, world!"-sized text); longer texts (up to kLoopMaxCtx tokens) stream the
// remaining rows from L2 after it, 2 x 3 KB per warp and round.
// E <= 30 (two table rows per warp, everything in registers before the wait for x): the latency-critical common case
__device__ __forceinline__ void cross_attention_small(LoopSmem & S, const FrameLoopParams & p, Ctx & c, const LoopLayer & Ly) {
    const int cw = c.cw, lane = c.lane, E = p.E;
    float w3[3], v[3];
    ln_weights3<D>(Ly.n_xq, c.ctid, w3);
    const int r0 = c.b * RO, nr = max(0, min(RO, D - r0));
    // the tables are activation independent: fetch this warp's rows before waiting for x
    const int j0 = cw, j1 = cw + kCW;
    float4 m0[6], m1[6];
#pragma unroll
    for (int q = 0; q < 6; q++) {
        m0[q] = j0 < E ? __ldg(reinterpret_cast<const float4 *>(Ly.xm + (size_t)j0 * D + q * 128 + lane * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
        m1[q] = j1 < E ? __ldg(reinterpret_cast<const float4 *>(Ly.xm + (size_t)j1 * D + q * 128 + lane * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float nval = 0.0f;
    if (cw < nr && lane < E) nval = __ldg(Ly.xn + (size_t)lane * D + r0 + cw);
    load3_poll<D>(c.xin + p.xoff[X_XA], c.seq - 1, S.xs, c.ctid, v);
    LOOP_STAMP();
    ln3<D>(S, v, w3, S.vec, p.eps, c);
    float d0 = 0.0f, d1 = 0.0f;
#pragma unroll
    for (int q = 0; q < 6; q++) {
        const float4 xv = *reinterpret_cast<const float4 *>(S.vec + q * 128 + lane * 4);
        d0 = fmaf(m0[q].x, xv.x, d0); d0 = fmaf(m0[q].y, xv.y, d0); d0 = fmaf(m0[q].z, xv.z, d0); d0 = fmaf(m0[q].w, xv.w, d0);
        d1 = fmaf(m1[q].x, xv.x, d1); d1 = fmaf(m1[q].y, xv.y, d1); d1 = fmaf(m1[q].z, xv.z, d1); d1 = fmaf(m1[q].w, xv.w, d1);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { d0 += __shfl_xor_sync(0xffffffffu, d0, o); d1 += __shfl_xor_sync(0xffffffffu, d1, o); }
    if (lane == 0) { if (j0 < E) S.sc[j0] = d0; if (j1 < E) S.sc[j1] = d1; }
    cbar();
    // softmax over the E tokens and this CTA's output rows: warp r < nr computes row r, warp 0 emits
    if (cw < nr) {
        const float sj = lane < E ? S.sc[lane] : -INFINITY;
        const float mxs = warp_max(sj);
        const float e = lane < E ? expf(sj - mxs) : 0.0f;
        float sum = e, o = e * nval;
#pragma unroll
        for (int of = 16; of > 0; of >>= 1) { sum += __shfl_xor_sync(0xffffffffu, sum, of); o += __shfl_xor_sync(0xffffffffu, o, of); }
        if (lane == 0) S.outv[cw] = o * (1.0f / sum) + S.xs[r0 + cw];
    }
    cbar();
    const int npk = (nr + 2) / 3;
    if (c.ctid < npk * kR) {
        const int pk = c.ctid % npk, r = c.ctid / npk;
        st_pkt(p.xbuf + (size_t)r * c.rstride + p.xoff[X_XB] + r0 / 3 + pk, S.outv[3 * pk], S.outv[3 * pk + 1], S.outv[3 * pk + 2], c.seq);
    }
}

// longer texts: the same phase with the table rows beyond the first 30 streamed from L2 after the wait (a separate branch: the
// common case keeps the round-1 instruction sequence)
constexpr int kScOff = 1024;             // scores live in S.vec[kScOff .. kScOff + kLoopMaxCtx) (LN(x) occupies [0, D))
__device__ __forceinline__ void cross_attention_long(LoopSmem & S, const FrameLoopParams & p, Ctx & c, const LoopLayer & Ly) {
    const int cw = c.cw, lane = c.lane, E = p.E;
    float * sc = S.vec + kScOff;
    float w3[3], v[3];
    ln_weights3<D>(Ly.n_xq, c.ctid, w3);
    const int r0 = c.b * RO, nr = max(0, min(RO, D - r0));
    // the tables are activation independent: fetch this warp's first rows before waiting for x
    float4 m0[6], m1[6];
    {
        const int j0 = cw, j1 = cw + kCW;
#pragma unroll
        for (int q = 0; q < 6; q++) {
            m0[q] = j0 < E ? __ldg(reinterpret_cast<const float4 *>(Ly.xm + (size_t)j0 * D + q * 128 + lane * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
            m1[q] = j1 < E ? __ldg(reinterpret_cast<const float4 *>(Ly.xm + (size_t)j1 * D + q * 128 + lane * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    float nval = 0.0f;
    if (cw < nr && lane < E) nval = __ldg(Ly.xn + (size_t)lane * D + r0 + cw);
    load3_poll<D>(c.xin + p.xoff[X_XA], c.seq - 1, S.xs, c.ctid, v);
    LOOP_STAMP();
    ln3<D>(S, v, w3, S.vec, p.eps, c);
    for (int jb = 0; jb < E; jb += 2 * kCW) {
        const int j0 = jb + cw, j1 = jb + cw + kCW;
        if (jb > 0) {
#pragma unroll
            for (int q = 0; q < 6; q++) {
                m0[q] = j0 < E ? __ldg(reinterpret_cast<const float4 *>(Ly.xm + (size_t)j0 * D + q * 128 + lane * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
                m1[q] = j1 < E ? __ldg(reinterpret_cast<const float4 *>(Ly.xm + (size_t)j1 * D + q * 128 + lane * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        float d0 = 0.0f, d1 = 0.0f;
#pragma unroll
        for (int q = 0; q < 6; q++) {
            const float4 xv = *reinterpret_cast<const float4 *>(S.vec + q * 128 + lane * 4);
            d0 = fmaf(m0[q].x, xv.x, d0); d0 = fmaf(m0[q].y, xv.y, d0); d0 = fmaf(m0[q].z, xv.z, d0); d0 = fmaf(m0[q].w, xv.w, d0);
            d1 = fmaf(m1[q].x, xv.x, d1); d1 = fmaf(m1[q].y, xv.y, d1); d1 = fmaf(m1[q].z, xv.z, d1); d1 = fmaf(m1[q].w, xv.w, d1);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { d0 += __shfl_xor_sync(0xffffffffu, d0, o); d1 += __shfl_xor_sync(0xffffffffu, d1, o); }
        if (lane == 0) { if (j0 < E) sc[j0] = d0; if (j1 < E) sc[j1] = d1; }
    }
    cbar();
    // softmax over the E tokens and this CTA's output rows: warp r < nr computes row r, warp 0 emits
    if (cw < nr) {
        float mxs = -INFINITY;
        for (int j = lane; j < E; j += 32) mxs = fmaxf(mxs, sc[j]);
        mxs = warp_max(mxs);
        float sum = 0.0f, o = 0.0f;
        for (int j = lane; j < E; j += 32) {
            const float e = expf(sc[j] - mxs);
            const float nv = j < 32 ? nval : __ldg(Ly.xn + (size_t)j * D + r0 + cw);
            sum += e; o = fmaf(e, nv, o);
        }
#pragma unroll
        for (int of = 16; of > 0; of >>= 1) { sum += __shfl_xor_sync(0xffffffffu, sum, of); o += __shfl_xor_sync(0xffffffffu, o, of); }
        if (lane == 0) S.outv[cw] = o * (1.0f / sum) + S.xs[r0 + cw];
    }
    cbar();
    const int npk = (nr + 2) / 3;
    if (c.ctid < npk * kR) {
        const int pk = c.ctid % npk, r = c.ctid / npk;
        st_pkt(p.xbuf + (size_t)r * c.rstride + p.xoff[X_XB] + r0 / 3 + pk, S.outv[3 * pk], S.outv[3 * pk + 1], S.outv[3 * pk + 2], c.seq);
    }
}

__device__ __forceinline__ void cross_attention(LoopSmem & S, const FrameLoopParams & p, Ctx & c, const LoopLayer & Ly) {
    if (p.E <= 2 * kCW) cross_attention_small(S, p, c, Ly);
    else cross_attention_long(S, p, c, Ly);
}

__global__ void __launch_bounds__(kThreads, 1) frame_loop_kernel(const FrameLoopParams p) {
    extern __shared__ __align__(128) unsigned char loop_smem[];
    LoopSmem & S = *reinterpret_cast<LoopSmem *>(loop_smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.x;

    if (tid == 0) {
        for (int i = 0; i < kQD; i++) { mbar_init(&S.full_bar[i], 1); mbar_init(&S.empty_bar[i], kCW); S.q_off[i] = 0; }
        mbar_init(&S.res_bar, 1);
        S.stop = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) { S.red[tid] = 0.0f; S.red2[tid] = 0.0f; }
    __syncthreads();
    if (warp == kCW) {
        if (lane == 0) prefetch_lane(S, p, b);
        return;
    }

    Ctx c;
    c.b = b; c.ctid = tid; c.cw = warp; c.lane = lane; c.kc = 0;
    c.seq = *p.seq; c.rstride = (size_t)p.xoff[X_COUNT];
    c.xin = p.xbuf + (size_t)(b % kR) * c.rstride;
    c.dbg_on = false; c.dbg_i = 0;
    const int V = p.V, L = p.L;
    const int ctid = c.ctid, cw = c.cw;
    const float lt_scale = 1.0f / sqrtf((float)LD);

    for (int i = ctid; i < 8 * LD; i += kCT) S.ltpos[i] = p.lt_pos[i];

    // decoder-input embedding accumulators: x0 = (sum_cb E_cb[code_cb]) / 8 + pos_emb   (magpie.cpp:2746-2787);
    // thread i < 256 owns elements 3i .. 3i+2 (the same mapping as the exchange packets)
    const bool own3 = ctid < D / 3;
    float emb[3] = {0.0f, 0.0f, 0.0f}, pe[3] = {0.0f, 0.0f, 0.0f};
    if (own3) {
        for (int cb = 0; cb < 8; cb++) {
            const int code = p.codes_io[cb];
#pragma unroll
            for (int q = 0; q < 3; q++) { const float a = p.audio_emb[cb][(size_t)code * D + 3 * ctid + q]; emb[q] = cb == 0 ? a : emb[q] + a; }
        }
#pragma unroll
        for (int q = 0; q < 3; q++) pe[q] = p.dec_pos[(size_t)p.pos0 * D + 3 * ctid + q];
    }
    const bool own3l = ctid < (LD + 2) / 3;
    int eos_step = -1, frames = 0;
    cbar();

    for (int t = 0; t < p.n_steps; t++) {
        const int pos = p.pos0 + t, nk = pos + 1;
        const size_t row = (size_t)p.row0 + t;
        c.dbg_on = p.dbg != nullptr && t == p.n_steps - 1;
        LOOP_STAMP();
        float xv[3];
#pragma unroll
        for (int q = 0; q < 3; q++) xv[q] = emb[q] * 0.125f + pe[q];
        if (own3) { S.xs[3 * ctid] = xv[0]; S.xs[3 * ctid + 1] = xv[1]; S.xs[3 * ctid + 2] = xv[2]; }
        if (own3 && t + 1 < p.n_steps) {      // next frame's position row: fetched a whole frame ahead
#pragma unroll
            for (int q = 0; q < 3; q++) pe[q] = __ldg(p.dec_pos + (size_t)(pos + 1) * D + 3 * ctid + q);
        }

        const int S_split = min(p.max_split, max(1, (nk + 127) / 128));
#pragma unroll 1
        for (int l = 0; l < L; l++) {
            const LoopLayer & Ly = p.layer[l];
            // ---- P1: LN -> QKV ---------------------------------------------------------------------------
            {
                float w3[3];
                ln_weights3<D>(Ly.n_self, ctid, w3);
                if (l > 0) load3_poll<D>(c.xin + p.xoff[X_XC], c.seq - 1, S.xs, ctid, xv);
                LOOP_STAMP();
                ln3<D>(S, xv, w3, S.vec, p.eps, c);
                const int r0 = b * RQ, nr = max(0, min(RQ, 3 * D - r0));
                if (nr > 0) {
                    const bf * w = ring_wait(S, c);
                    gemv_rows<D / 256, 4>(w, nr, S.vec, S.part, c);
                    ring_release(S, c);
                }
                cbar();
                emit_rows<D / 256, EPI_NONE>(S, p, c, X_QKV, r0, nr, nullptr, nullptr, nullptr);
                c.seq++;
            }
            LOOP_STAMP();
            // ---- P2: attention partials, item (head, key split) --------------------------------------------
            if (b < H * S_split) {
                const int h = b / S_split, sp = b % S_split;
                const int per = (nk + S_split - 1) / S_split;
                attention_item(S, p, c, (bf *)p.kcache + (size_t)l * p.kv_layer_stride, (bf *)p.vcache + (size_t)l * p.kv_layer_stride,
                               h, sp * per, min(nk, sp * per + per), pos);
            }
            c.seq++;
            LOOP_STAMP();
            // ---- P3: combine partials -> attention output; O projection + residual -------------------------
            {
                attention_combine(S, p, c, S_split);
                const int r0 = b * RO, nr = max(0, min(RO, D - r0));
                if (nr > 0) {
                    const bf * w = ring_wait(S, c);
                    gemv_rows<D / 256, 2>(w, nr, S.av, S.part, c);
                    ring_release(S, c);
                }
                cbar();
                emit_rows<D / 256, EPI_RES>(S, p, c, X_XA, r0, nr, S.xs, nullptr, nullptr);
                c.seq++;
            }
            LOOP_STAMP();
            // ---- P4: LN -> folded cross-attention + residual ---------------------------------------------------
            cross_attention(S, p, c, Ly);
            c.seq++;
            LOOP_STAMP();
            // ---- P5: LN -> FFN1 -> GELU ------------------------------------------------------------------------
            {
                float w3[3];
                ln_weights3<D>(Ly.n_ff, ctid, w3);
                load3_poll<D>(c.xin + p.xoff[X_XB], c.seq - 1, S.xs, ctid, xv);
                LOOP_STAMP();
                ln3<D>(S, xv, w3, S.vec, p.eps, c);
                const int r0 = b * RF1, nr = max(0, min(RF1, F - r0));
                if (nr > 0) {
                    const bf * w = ring_wait(S, c);
                    gemv_rows<D / 256, 8>(w, nr, S.vec, S.part, c);
                    ring_release(S, c);
                }
                cbar();
                emit_rows<D / 256, EPI_GELU>(S, p, c, X_H, r0, nr, nullptr, nullptr, nullptr);
                c.seq++;
            }
            LOOP_STAMP();
            // ---- P6: FFN2 + residual ---------------------------------------------------------------------------
            {
                poll_vec(c.xin + p.xoff[X_H], F / 3, c.seq - 1, S.vec, F, ctid);
                cbar();
                LOOP_STAMP();
                const int r0 = b * RF2, nr = max(0, min(RF2, D - r0));
                if (nr > 0) {
                    const bf * w = ring_wait(S, c);
                    gemv_rows<F / 256, 8>(w, nr, S.vec, S.part, c);
                    ring_release(S, c);
                }
                cbar();
                emit_rows<F / 256, EPI_RES>(S, p, c, X_XC, r0, nr, S.xs, nullptr, nullptr);
                c.seq++;
            }
            LOOP_STAMP();
        }

        // ---- final LayerNorm -> hidden (every CTA holds it) ------------------------------------------------------
        {
            float w3[3];
            ln_weights3<D>(p.norm_out, ctid, w3);
            load3_poll<D>(c.xin + p.xoff[X_XC], c.seq - 1, S.xs, ctid, xv);
            LOOP_STAMP();
            ln3<D>(S, xv, w3, S.av, p.eps, c);
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
        // x_0 = in_proj . hidden + b + pos_lt[0]   (the position row is added by the producer)
        {
            const int r0 = b * RIN, nr = max(0, min(RIN, LD - r0));
            gemv_rows<D / 256, 1>(S.w_in, nr, S.av, S.part, c);
            cbar();
            emit_rows<D / 256, EPI_BIAS>(S, p, c, T_SEQ0, r0, nr, nullptr, p.lt_in_b, S.ltpos);
            c.seq++;
        }
        float fb[3] = {0.0f, 0.0f, 0.0f};     // feedback row elements 3i .. 3i+2 for the next codebook
        unsigned amax_seq0 = 0u, amax_seq1 = 0u;   // exchange numbers of the two argmax regions
        float4 row_pf = make_float4(0.0f, 0.0f, 0.0f, 0.0f); bool have_row_pf = false;      // (teacher forcing: next position's table row, see E)
        const bool defer_amax = p.forced != nullptr && !sampling && !p.no_defer_amax;
        int fed_prev = 0;                      // code fed back by the previous codebook
#pragma unroll 1
        for (int cb = 0; cb < 8; cb++) {
            float * qkv = S.lqkv[cb];         // [q | k | v] of position cb
            // ---- A (position 0 only): x = seq + pos; LN -> QKV.  Positions 1..7 see x = P_cb[fed] + pos[cb], a function of
            //      (codebook, fed code) alone: their [q | k | vo] row is GATHERED from a table built at load (model.cu),
            //      which removes the LayerNorm, the GEMV and the exchange of its result ------------------------------
            if (cb == 0) {
                float w3[3], lv[3];
                ln_weights3<LD>(p.lt_norm_self, ctid, w3);
                load3_poll<LD>(c.xin + p.xoff[T_SEQ0], c.seq - 1, S.lx, ctid, lv);
                LOOP_STAMP();
                ln3<LD>(S, lv, w3, S.vec, p.eps, c);
                const int r0 = b * RLQ, nr = max(0, min(RLQ, LQ - r0));
                gemv_rows<1, 1>(S.w_qkv, nr, S.vec, S.part, c);
                cbar();
                emit_rows<1, EPI_NONE>(S, p, c, T_QKV, r0, nr, nullptr, nullptr, nullptr);
                c.seq++;
            } else {
                if (own3l) {
#pragma unroll
                    for (int q = 0; q < 3; q++) if (3 * ctid + q < LD) S.lx[3 * ctid + q] = fb[q] + S.ltpos[cb * LD + 3 * ctid + q];
                }
                LOOP_STAMP();
            }
            LOOP_STAMP();
            // ---- B: attention over the <= 8 positions, redundantly in every CTA.  The O-projection is folded into the value
            //      rows (vo_j = Wo Wv n_j, hi + lo), so x1 = x + sum_j p_j vo_j needs no GEMV and NO exchange of its own ----
            {
                if (cb == 0) {
                    poll_vec(c.xin + p.xoff[T_QKV], (LQ + 2) / 3, c.seq - 1, S.vec, LQ, ctid);
                    cbar();
                } else {
                    const float4 * row = reinterpret_cast<const float4 *>(p.lt_qkv_tab + ((size_t)(cb - 1) * V + fed_prev) * (3 * LD));
                    if (ctid < 3 * LD / 4) {
                        const float4 f = have_row_pf ? row_pf : __ldg(row + ctid);
                        S.vec[4 * ctid] = f.x; S.vec[4 * ctid + 1] = f.y; S.vec[4 * ctid + 2] = f.z; S.vec[4 * ctid + 3] = f.w;
                    }
                    if (ctid < LD) S.vec[3 * LD + ctid] = 0.0f;       // no separate lo part: the table holds hi + lo
                    cbar();
                }
                if (ctid < LD) {
                    qkv[ctid] = S.vec[ctid]; qkv[LD + ctid] = S.vec[LD + ctid];
                    qkv[2 * LD + ctid] = S.vec[2 * LD + ctid] + S.vec[3 * LD + ctid];
                }
                cbar();
                LOOP_STAMP();
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
                    for (int j = 0; j <= cb; j++) {
                        const float e = expf(S.sc[j] - mxs);
                        sum += e;
                        o = fmaf(e, S.lqkv[j][2 * LD + ctid], o);
                    }
                    S.lx1[ctid] = S.lx[ctid] + o * (1.0f / sum);
                }
                cbar();
            }
            LOOP_STAMP();
            // ---- C: LN -> FFN1 -> GELU -----------------------------------------------------------------------------
            {
                float w3[3], lv[3];
                ln_weights3<LD>(p.lt_norm_ff, ctid, w3);
#pragma unroll
                for (int q = 0; q < 3; q++) lv[q] = (own3l && 3 * ctid + q < LD) ? S.lx1[3 * ctid + q] : 0.0f;
                LOOP_STAMP();
                ln3<LD>(S, lv, w3, S.vec, p.eps, c);
                const int r0 = b * RLF1, nr = max(0, min(RLF1, LF - r0));
                gemv_rows<1, 1>(S.w_ff1, nr, S.vec, S.part, c);
                cbar();
                emit_rows<1, EPI_GELU>(S, p, c, T_H, r0, nr, nullptr, nullptr, nullptr);
                c.seq++;
            }
            LOOP_STAMP();
            // ---- D: FFN2 + residual ---------------------------------------------------------------------------------
            {
                poll_vec(c.xin + p.xoff[T_H], (LF + 2) / 3, c.seq - 1, S.vec, LF, ctid);
                cbar();
                LOOP_STAMP();
                const int r0 = b * RLF2, nr = max(0, min(RLF2, LD - r0));
                gemv_rows<LF / 256, 1>(S.w_ff2, nr, S.vec, S.part, c);
                cbar();
                emit_rows<LF / 256, EPI_RES>(S, p, c, T_HOUT, r0, nr, S.lx1, nullptr, nullptr);
                c.seq++;
            }
            LOOP_STAMP();
            // ---- E: output projection of codebook cb (+bias, forbidden-token mask); local argmax ---------------------
            // (teacher forcing: the code fed to position cb + 1 is known, so its feedback row, its [q | k | vo] table row and the next
            //  frame's embedding rows are requested HERE, ahead of the wait for FFN2, instead of on the chain after the argmax)
            float4 pf_row = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            float pf_fb[3] = {0.0f, 0.0f, 0.0f}, pf_emb[3] = {0.0f, 0.0f, 0.0f};
            if (defer_amax) {
                const int fedn = p.forced[row * 8 + cb];
                if (cb < 7) {
                    if (ctid < 3 * LD / 4) pf_row = __ldg(reinterpret_cast<const float4 *>(p.lt_qkv_tab + ((size_t)cb * V + fedn) * (3 * LD)) + ctid);
                    if (own3l) {
#pragma unroll
                        for (int q = 0; q < 3; q++) if (3 * ctid + q < LD) pf_fb[q] = __ldg(p.lt_in_table[cb] + (size_t)fedn * LD + 3 * ctid + q);
                    }
                }
                if (own3) {
#pragma unroll
                    for (int q = 0; q < 3; q++) pf_emb[q] = __ldg(p.audio_emb[cb] + (size_t)fedn * D + 3 * ctid + q);
                }
            }
            {
                poll_vec(c.xin + p.xoff[T_HOUT], (LD + 2) / 3, c.seq - 1, S.lhout, LD, ctid);
                cbar();
                LOOP_STAMP();
                const int r0 = b * ROUT, nr = max(0, min(ROUT, V - r0));
                const float obias = (cw == 0 && lane < nr) ? __ldg(p.lt_out_b[cb] + r0 + lane) : 0.0f;      // (requested ahead of the GEMV)
                if (nr > 0) {
                    const bf * w = ring_wait(S, c);
                    gemv_rows<1, 1>(w, nr, S.lhout, S.part, c);
                    ring_release(S, c);
                }
                cbar();
                if (cw == 0 && nr > 0) {
                    float v = -INFINITY;
                    if (lane < nr) {
                        const int n = r0 + lane;
                        v = S.part[lane * kPartStride] + obias;
                        const bool masked = n == p.bos_id || (n >= p.bos_id + 2 && n <= p.bos_id + 7) || (forbid_eos && n == p.eos_id);
                        if (masked) v = -INFINITY;                                // magpie.cpp:1131-1145, 1243-1248
                        if (p.logits) p.logits[(row * 8 + cb) * V + n] = v;
                    }
                    if (sampling) {
                        const int npk = (nr + 2) / 3;
                        const int pk = lane % npk, rep = lane / npk;
                        const float v0 = __shfl_sync(0xffffffffu, v, 3 * pk), v1 = __shfl_sync(0xffffffffu, v, 3 * pk + 1), v2 = __shfl_sync(0xffffffffu, v, 3 * pk + 2);
                        if (lane < npk * kR) st_pkt(p.xbuf + (size_t)rep * c.rstride + p.xoff[T_LOGITS] + r0 / 3 + pk, v0, v1, v2, c.seq);
                    } else {
                        float bv = v; int bi = lane < nr ? r0 + lane : 0x7fffffff;
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) {
                            const float ov = __shfl_xor_sync(0xffffffffu, bv, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                            if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
                        }
                        if (lane < kR) st_pkt(p.xbuf + (size_t)lane * c.rstride + p.xoff[T_AMAX] + (defer_amax ? (cb & 1) * kAmaxSlots : 0) + b, bv, __int_as_float(bi), 0.0f, c.seq);
                    }
                }
                if (cb & 1) amax_seq1 = c.seq; else amax_seq0 = c.seq;
                c.seq++;
            }
            LOOP_STAMP();
            // ---- F: global argmax / top-k sample; feedback ---------------------------------------------------------------
            // Teacher forcing + greedy (BASELINE config 2): the code fed back is known, so nothing downstream waits for the argmax of
            // this codebook.  Its packets (own region per codebook parity) are collected one codebook LATER, when they have long
            // arrived, which takes this exchange's wait off the chain for codebooks 0..6; codebook 7 collects 6 and 7.  (A region is
            // rewritten two codebooks later, after every CTA has emitted FF1 of the codebook in between, i.e. after its collection.)
            auto collect_amax = [&](int cbx) -> int {
                const int nprod = (V + ROUT - 1) / ROUT;
                float bv = -INFINITY; int bi = 0x7fffffff;
                if (ctid < nprod) {
                    const uint4 v = poll_pkt(c.xin + p.xoff[T_AMAX] + (cbx & 1) * kAmaxSlots + ctid, (cbx & 1) ? amax_seq1 : amax_seq0);
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
                bv = S.red[0]; bi = S.redi[0];
                for (int w = 1; w < kCW; w++) if (better(S.red[w], S.redi[w], bv, bi)) { bv = S.red[w]; bi = S.redi[w]; }
                cbar();
                return bi;
            };
            if (defer_amax) {
                const int fedd = p.forced[row * 8 + cb];
                for (int cbx = (cb == 0 ? 1 : cb - 1); cbx <= (cb == 7 ? 7 : cb - 1); cbx++) {      // cb 0: none; cb 1..6: cb - 1; cb 7: 6 and 7
                    const int amx = collect_amax(cbx);
                    hit_eos = hit_eos || amx == p.eos_id;
                    if (b == 0 && ctid == 0) { p.argmax[row * 8 + cbx] = amx; p.sampled[row * 8 + cbx] = amx; }
                }
                if (b == 0 && ctid == 0) p.result[2 + cb] = fedd;
                fed_prev = fedd;
                if (cb < 7 && own3l) {
#pragma unroll
                    for (int q = 0; q < 3; q++) fb[q] = pf_fb[q];
                }
                if (own3) {
#pragma unroll
                    for (int q = 0; q < 3; q++) emb[q] = cb == 0 ? pf_emb[q] : emb[q] + pf_emb[q];
                }
                row_pf = pf_row; have_row_pf = true;
                LOOP_STAMP(); LOOP_STAMP();
                continue;
            }
            int am, pick;
            if (!sampling) {
                const int nprod = (V + ROUT - 1) / ROUT;
                float bv = -INFINITY; int bi = 0x7fffffff;
                if (ctid < nprod) {
                    const uint4 v = poll_pkt(c.xin + p.xoff[T_AMAX] + ctid, c.seq - 1);
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
                poll_vec(c.xin + p.xoff[T_LOGITS], (V + 2) / 3, c.seq - 1, S.vec, V, ctid);
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
            fed_prev = fed;
            // feedback: seq[cb+1] = row `fed` of P_cb (no 1/8 scale, magpie.cpp:1285-1291); next frame's embedding
            if (cb < 7 && own3l) {
#pragma unroll
                for (int q = 0; q < 3; q++) if (3 * ctid + q < LD) fb[q] = __ldg(p.lt_in_table[cb] + (size_t)fed * LD + 3 * ctid + q);
            }
            if (own3) {
#pragma unroll
                for (int q = 0; q < 3; q++) { const float a = __ldg(p.audio_emb[cb] + (size_t)fed * D + 3 * ctid + q); emb[q] = cb == 0 ? a : emb[q] + a; }
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
        *p.seq = c.seq;
    }
}

// ---- folded cross-attention tables -----------------------------------------------------------------------------------
__global__ void xattn_fold_kernel(const bf * xk, const bf * xv, const bf * wq, const bf * wo, int d, int dxa, float scale,
                                  float * xm, float * xn, const int32_t * n_ctx, int rows_per_utt) {
    const int j = blockIdx.x;                 // row of the cross K/V storage: (utterance, text token)
    if (n_ctx && (j % rows_per_utt) >= n_ctx[j / rows_per_utt]) return;       // rows past the text are never read
    extern __shared__ float kv[];             // [2][dxa]
    for (int i = threadIdx.x; i < dxa; i += blockDim.x) {
        kv[i] = __bfloat162float(xk[(size_t)j * dxa + i]);
        kv[dxa + i] = __bfloat162float(xv[(size_t)j * dxa + i]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < d; i += blockDim.x) {
        float m = 0.0f, n = 0.0f;
        // M row: column i of Wq (coalesced across the threads); N row: row i of Wo, contiguous per thread (16-byte loads)
        for (int cc = 0; cc < dxa; cc++) m = fmaf(kv[cc], __bfloat162float(wq[(size_t)cc * d + i]), m);
        const uint4 * wr = reinterpret_cast<const uint4 *>(wo + (size_t)i * dxa);
        for (int c8 = 0; c8 < dxa / 8; c8++) {
            const uint4 u = wr[c8];
            const float * v = kv + dxa + c8 * 8;
            n = fmaf(bf16lo(u.x), v[0], n); n = fmaf(bf16hi(u.x), v[1], n); n = fmaf(bf16lo(u.y), v[2], n); n = fmaf(bf16hi(u.y), v[3], n);
            n = fmaf(bf16lo(u.z), v[4], n); n = fmaf(bf16hi(u.z), v[5], n); n = fmaf(bf16lo(u.w), v[6], n); n = fmaf(bf16hi(u.w), v[7], n);
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
    const int n[X_COUNT] = {3 * D / 3, H * kMaxSplit * 22, D / 3, D / 3, F / 3, D / 3, (LD + 2) / 3, (LQ + 2) / 3, (LD + 2) / 3, (LF + 2) / 3,
                            (LD + 2) / 3, 2 * kAmaxSlots, (V + 2) / 3 + 8};      // (T_AMAX: two regions, see the deferred argmax in the LT)
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
                       float * xm, float * xn, cudaStream_t stream, const int32_t * n_ctx, int rows_per_utt) {
    if (dxa % 8 != 0) { set_error("xattn_fold: cross-attention width must be a multiple of 8"); return false; }
    xattn_fold_kernel<<<E, 256, 2 * dxa * sizeof(float), stream>>>((const bf *)xk, (const bf *)xv, (const bf *)wq, (const bf *)wo, d, dxa,
                                                                   scale, xm, xn, n_ctx, rows_per_utt);
    MGB_LAUNCH_CHECK();
    return true;
}

}  // namespace mgb
