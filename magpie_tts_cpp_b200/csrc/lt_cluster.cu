// Local transformer + sampler for 4..80 utterances: ONE 16-CTA CLUSTER PER GROUP OF U <= 10 UTTERANCES, weights resident in the
// cluster's shared memory, phases separated by cluster barriers and fed through distributed shared memory.
//
// Same arithmetic as lt_batch.cu / lt_kernel.cu (reference src/magpie.cpp:946-1048 LT layer, 1072-1109 sample_top_k, 1113-1317
// magpie_local_transformer_sample_all).  lt_batch.cu slices every LT matrix over all 148 CTAs and synchronises its 35 phases with
// GRID barriers; every phase there re-stages the activations of all utterances into every CTA (64 x 256 floats = 65 KB through L2 per
// CTA and pass): 4.4-9.8 us of work + 1.5 us of barrier per phase, 304 us per step at 64 utterances (profiles/r1_lt_batch_phase_timeline.txt).
// Here a cluster owns its utterances end to end:
//   * every CTA keeps its slices of the FFN matrices and of the current codebook's output projection in shared memory as tcgen05
//     operand images (bf16 SWIZZLE_128B K-major tiles, the 128-row images model.cu packs for gemm_tc.cu: 128 KB, bulk copies);
//   * the three matrix products of a codebook run on the tensor cores: weights = MMA "A" (M = 64 / 128 rows), the cluster's utterances
//     = MMA "B" (N = 16 columns, bf16 hi + lo halves of the f32 activations, two MMAs per k slice), accumulators in TMEM (64 columns);
//     one elected thread issues 16-32 MMAs per product, all 16 warps read the accumulators back (tcgen05.ld) for the epilogues.  (The
//     CUDA-core version of these products was issue-bound: 13.6 us per codebook for 650 k FMAs per CTA; profiles/r2_lt_cluster_phases.txt);
//   * FF1 and FF2 are FUSED without an exchange: CTA r computes FFN-hidden rows [64 r, 64 r + 64) and multiplies them straight into
//     the matching 64 COLUMNS of W2 (= k tile r of the packed W2); the 16 partial sums per output are reduce-scattered through DSMEM
//     and added in rank order (deterministic);
//   * results go only where they are needed: activations (as operand images) to all 16 CTAs, logits / q / k / vo rows to the
//     utterance's owner CTA;
//   * the phase boundaries are cluster barriers (~0.6-2 us) instead of grid barriers.
// A GPU schedules 7 or 8 such clusters side by side (one per GPC with >= 16 SMs), so 64 utterances are 8 x 8 or 7 x 10.
//   per codebook:  FF1 (LN prologue, GELU) + FF2 partials | reduce + residual, gather | out-projection + bias | owner CTA: mask /
//                  argmax / top-k sample, gather of the next position's [q | k | vo] row (model.cu table), its attention, x1 broadcast
#include <cstdlib>

#include "gemm_tc.cuh"
#include "lt_common.cuh"

namespace mgb {

namespace {

using namespace lt;
using bf = __nv_bfloat16;

constexpr int kCS = 16;                                  // CTAs per cluster
constexpr int kF1Rows = kF / kCS;                        // 64 FF1 rows (= W2 columns) per CTA
constexpr int kF2Rows = kL / kCS;                        // 16 layer-output rows reduced per CTA
constexpr int kOutRows = (kV + kCS - 1) / kCS;           // 128 out-projection rows per CTA (127 used at V = 2024)
constexpr int kXL = kL + 4;                              // padded activation row stride (f32)
constexpr int kDm = 768, kHS = kDm + 4;                  // decoder width this kernel is built for; padded hidden-state row stride
constexpr int kScratch = 10 * kHS;                       // floats: hidden states of <= 10 utterances (prologue) / sampler scratch (owner phase)

constexpr int kBTile = 16 * 128;                         // one k tile of the utterance operand: 16 rows (utterances) x 64 k, bf16
constexpr int kTmemCols = 128;                           // accumulators (hi | lo column halves): FF1 [0, 32), FF2 partials [32, 96), out-projection [96, 128)

template <int U> struct alignas(1024) ClSmem {
    // tcgen05 operand images (every one starts on a 1024-byte boundary)
    bf w_ff1[kF1Rows * kL];            // FF1 rows [64 r, 64 r + 64): 4 k tiles of [64 rows x 128 B]
    bf w_ff2[kL * kF1Rows];            // W2[:, 64 r .. 64 r + 64): 2 tiles of [128 rows x 128 B]; in the prologue: the vo "lo" rows of position 0 (row-major)
    bf w_out[kOutRows * kL];           // current codebook's output-projection rows [128 r, 128 r + 128): 4 k tiles of [128 rows x 128 B];
                                       // in the prologue: in-projection + [q | k | vo-hi] rows (row-major)
    unsigned char b_act[kL / 64][2][kBTile];   // [k tile][hi | lo][16 utterances x 128 B] = 32-row operand images: LN(x1) for FF1, then the layer output for the out-projection
    unsigned char b_ffh[2][kBTile];            // [hi | lo] this CTA's 64 GELU'd FFN-hidden values per utterance
    float scratch[kScratch];           // prologue: decoder hidden states [U][772], then LN(seq) [U][260]; owner phase: the sampler's scratch
    float x1[U][kXL];                  // residual stream entering the FFN (replicated in every CTA)
    float recv[kCS][U][kF2Rows];       // FF2 partial sums of this CTA's 16 output rows from every rank
    // owner CTA of an utterance (rank u owns utterance u of the cluster)
    float logits[kV];
    float q[kL], seq[kL], vlo[kL];     // vlo: the "lo" half of position 0's folded value row (added to vc[0] by the owner)
    float kc[8][kL], vc[8][kL];
    float red[32]; int redi[32];
    float scores[8];
    uint64_t mbar[4];                  // 0: FFN images, 1: out-projection image, 2: prologue rows, 3: MMA completion
    uint32_t tmem_slot;
};

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
__device__ __forceinline__ void bulk_g2s(void * dst, const void * src, uint32_t bytes, uint64_t * bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t dsmem_addr(const void * local, int rank) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(local)), "r"(rank));
    return ra;
}
__device__ __forceinline__ void dsmem_st(float * local, int rank, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(dsmem_addr(local, rank)), "f"(v) : "memory");
}
__device__ __forceinline__ void dsmem_st4(float * local, int rank, float4 v) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(dsmem_addr(local, rank)), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void dsmem_st4_all(float * local, float4 v) {
#pragma unroll
    for (int r = 0; r < kCS; r++) dsmem_st4(local, r, v);
}

__device__ __forceinline__ void dsmem_st2u(void * local, int rank, uint32_t a, uint32_t b) {
    asm volatile("st.shared::cluster.v2.u32 [%0], {%1,%2};" ::"r"(dsmem_addr(local, rank)), "r"(a), "r"(b) : "memory");
}
// x = hi + lo (bf16 each, |x - hi - lo| <= 2^-17 |x|): the activation operand of the tensor-core products (gemm_tc.cuh)
__device__ __forceinline__ void split_bf(float x, uint32_t & h, uint32_t & l) {
    const bf hb = __float2bfloat16_rn(x);
    h = __bfloat16_as_ushort(hb);
    l = __bfloat16_as_ushort(__float2bfloat16_rn(x - __bfloat162float(hb)));
}
__device__ __forceinline__ void fence_async_proxy() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void tc_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ float tmem_ld1(uint32_t taddr) {
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
    return __uint_as_float(v);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// lanes of a dot product for `items` (row, utterance) pairs on 512 threads: the largest power of two <= 512 / items, at most 32
__host__ __device__ constexpr int ks_for(int items) {
    return 512 / items >= 32 ? 32 : 512 / items >= 16 ? 16 : 512 / items >= 8 ? 8 : 512 / items >= 4 ? 4 : 512 / items >= 2 ? 2 : 1;
}

// Register-tiled GEMV block: for rows [0, nrows) of w (bf16, shared, unpadded, row stride K) and utterances [0, U) of x (f32, shared,
// row stride ldx), out(r0, u, v[4]) receives the 4 dots of rows r0 .. r0 + 3 (r0 a multiple of 4) with x[u].  A thread owns a tile of
// 4 rows x 2 utterances (each 16-byte weight chunk it loads feeds 2 utterances, each activation chunk 4 rows: the phases are bound by
// shared-memory bandwidth otherwise) and, when there are fewer tiles than threads, 1/KS of the k range (shuffle reduction).  ROT: the
// 16-byte chunk order is rotated by the tile's row index (K / 8 a power of two), so that the row tiles a warp reads at the same time
// fall into different banks without padding.  Every thread of the CTA must call it.
template <int NROWS, int U, int K, bool ROT, typename Out>
__device__ __forceinline__ void tile_dots(const bf * w, int nrows, const float * x, int ldx, Out out) {
    constexpr int NRT = (NROWS + 3) / 4, NUT = (U + 1) / 2, TILES = NRT * NUT, KS = ks_for(TILES), NC = K / 8;
    for (int base = 0; base < TILES * KS; base += kLtThreads) {
        const int it = base + (int)threadIdx.x, itc = min(it, TILES * KS - 1);
        const int kq = itc % KS, tu = (itc / KS) % NUT, tr = itc / (KS * NUT);
        const int u0 = 2 * tu, u1 = min(2 * tu + 1, U - 1);
        const bf * wr[4];
#pragma unroll
        for (int i = 0; i < 4; i++) wr[i] = w + (size_t)min(4 * tr + i, nrows - 1) * K;
        const float * x0 = x + (size_t)u0 * ldx, * x1 = x + (size_t)u1 * ldx;
        float acc[4][2];
#pragma unroll
        for (int i = 0; i < 4; i++) { acc[i][0] = 0.0f; acc[i][1] = 0.0f; }
        for (int c = kq; c < NC; c += KS) {
            const int cc = ROT ? ((c + tr) & (NC - 1)) : c;
            const float4 a0 = *reinterpret_cast<const float4 *>(x0 + cc * 8), a1 = *reinterpret_cast<const float4 *>(x0 + cc * 8 + 4);
            const float4 b0 = *reinterpret_cast<const float4 *>(x1 + cc * 8), b1 = *reinterpret_cast<const float4 *>(x1 + cc * 8 + 4);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint4 q = *reinterpret_cast<const uint4 *>(wr[i] + cc * 8);
                const float w0 = bf16lo(q.x), w1 = bf16hi(q.x), w2 = bf16lo(q.y), w3 = bf16hi(q.y), w4 = bf16lo(q.z), w5 = bf16hi(q.z), w6 = bf16lo(q.w), w7 = bf16hi(q.w);
                float s0 = acc[i][0], s1 = acc[i][1];
                s0 = fmaf(w0, a0.x, s0); s0 = fmaf(w1, a0.y, s0); s0 = fmaf(w2, a0.z, s0); s0 = fmaf(w3, a0.w, s0);
                s0 = fmaf(w4, a1.x, s0); s0 = fmaf(w5, a1.y, s0); s0 = fmaf(w6, a1.z, s0); s0 = fmaf(w7, a1.w, s0);
                s1 = fmaf(w0, b0.x, s1); s1 = fmaf(w1, b0.y, s1); s1 = fmaf(w2, b0.z, s1); s1 = fmaf(w3, b0.w, s1);
                s1 = fmaf(w4, b1.x, s1); s1 = fmaf(w5, b1.y, s1); s1 = fmaf(w6, b1.z, s1); s1 = fmaf(w7, b1.w, s1);
                acc[i][0] = s0; acc[i][1] = s1;
            }
        }
#pragma unroll
        for (int o = KS / 2; o > 0; o >>= 1)
#pragma unroll
            for (int i = 0; i < 4; i++) { acc[i][0] += __shfl_xor_sync(0xffffffffu, acc[i][0], o); acc[i][1] += __shfl_xor_sync(0xffffffffu, acc[i][1], o); }
        if (kq == 0 && it < TILES * KS) {
            const float v0[4] = {acc[0][0], acc[1][0], acc[2][0], acc[3][0]}, v1[4] = {acc[0][1], acc[1][1], acc[2][1], acc[3][1]};
            out(4 * tr, u0, v0);
            if (2 * tu + 1 < U) out(4 * tr, u0 + 1, v1);
        }
    }
}

// LayerNorm (no bias, magpie.cpp:2237-2259: mean, centred variance) of U rows of L values, one warp per row: y = LN(x + add) * w
template <int U> __device__ __forceinline__ void warp_ln_rows(const float (*x)[kXL], const float * add, const float * w, float (*y)[kXL], float eps, int L) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp < U) {
        float v[kL / 32], s = 0.0f;
#pragma unroll
        for (int k = 0; k < kL / 32; k++) { const int i = lane + 32 * k; v[k] = i < L ? x[warp][i] + (add ? add[i] : 0.0f) : 0.0f; s += v[k]; }
        const float mean = warp_sum(s) / (float)L;
        float s2 = 0.0f;
#pragma unroll
        for (int k = 0; k < kL / 32; k++) { const int i = lane + 32 * k; v[k] = i < L ? v[k] - mean : 0.0f; s2 += v[k] * v[k]; }
        const float scale = 1.0f / sqrtf(warp_sum(s2) / (float)L + eps);
#pragma unroll
        for (int k = 0; k < kL / 32; k++) { const int i = lane + 32 * k; if (i < L) y[warp][i] = (v[k] * scale) * w[i]; }
    }
    __syncthreads();
}

// Same LayerNorm, written as the tensor-core "B" operand: row u of the [16 x K] K-major SWIZZLE_128B images (hi | lo halves).
// Lane l owns the 16-byte chunk of elements 8 l .. 8 l + 7.  Ends with the generic -> async proxy fence + __syncthreads.
template <int U> __device__ __forceinline__ void warp_ln_image(const float (*x)[kXL], const float * w, unsigned char (*img)[2][kBTile], float eps) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp < U) {
        const float4 a = *reinterpret_cast<const float4 *>(&x[warp][8 * lane]), b = *reinterpret_cast<const float4 *>(&x[warp][8 * lane + 4]);
        float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        float s = 0.0f;
#pragma unroll
        for (int k = 0; k < 8; k++) s += v[k];
        const float mean = warp_sum(s) / (float)kL;
        float s2 = 0.0f;
#pragma unroll
        for (int k = 0; k < 8; k++) { v[k] -= mean; s2 += v[k] * v[k]; }
        const float scale = 1.0f / sqrtf(warp_sum(s2) / (float)kL + eps);
        const float4 wa = *reinterpret_cast<const float4 *>(w + 8 * lane), wb = *reinterpret_cast<const float4 *>(w + 8 * lane + 4);
        const float ww[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
        uint32_t h[4], l[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint32_t h0, l0, h1, l1;
            split_bf((v[2 * k] * scale) * ww[2 * k], h0, l0);
            split_bf((v[2 * k + 1] * scale) * ww[2 * k + 1], h1, l1);
            h[k] = h0 | (h1 << 16); l[k] = l0 | (l1 << 16);
        }
        const int off = tc::swz_offset(warp, (lane & 7) * 8);
        *reinterpret_cast<uint4 *>(&img[lane >> 3][0][off]) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4 *>(&img[lane >> 3][1][off]) = make_uint4(l[0], l[1], l[2], l[3]);
    }
    tc_before_sync();
    __syncthreads();
}

// the sampler of lt_common.cuh works on any object with these members; here they point into shared memory
struct SampView {
    float * logits, * sel_v, * srt_v; uint16_t * sel_i, * srt_i; unsigned * hist; int * misc;
};
static_assert(2 * kV * 4 + 2 * kV * 2 + 256 * 4 + 8 * 4 <= kScratch * 4, "sampler scratch");

struct ClParams {
    unsigned long long * dbg;        // MGB_LT_DBG: globaltimer stamps of CTA 0 at every phase boundary
    int dbg_fine;                    // MGB_LT_DBG=2: additional stamps inside the phases (LN | FF1 | barrier | FF2 partials; barrier wait | out-projection)
    LtParams p;
    const void * qkvo;               // [4L][L] bf16: [Wq; Wk; hi(Wo Wv); lo(Wo Wv)]
    const float * qkv_tab;           // [7][V][3L] f32: [q | k | vo] of position cb+1 per fed code of codebook cb
    const void * ff1_t, * ff2_t;     // packed 128-row tile images of W1 [F][L] and W2 [L][F] (gemm_tc.cu pack_w_kernel)
    const void * out_t[8];           // same for the 8 output projections [V -> 128-row tiles][L]
};

template <int U>
__global__ void __launch_bounds__(kLtThreads, 1) lt_cluster_kernel(const ClParams cp) {
    extern __shared__ unsigned char smem_raw[];
    ClSmem<U> & S = *reinterpret_cast<ClSmem<U> *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const LtParams & p = cp.p;
    unsigned rank_u;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank_u));
    const int rank = (int)rank_u, cl = blockIdx.x / kCS;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int B = p.B, d = kDm, L = kL, V = p.V;             // (lt_cluster_plan admits only these widths)
    const int u0 = cl * U;                                   // first utterance of this cluster
    const float att_scale = 1.0f / sqrtf((float)L);
    const bool owner = rank < U && u0 + rank < B;            // this CTA samples / attends for utterance u0 + rank
    const int my_utt = u0 + rank;
    if (warp == 0) {                                         // TMEM: 64 fp32 columns for the three accumulators
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_slot)), "n"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }

    // output-projection rows of this CTA: vocabulary ids [128 r, 128 r + 128) = row tile r of the packed matrix (zero rows beyond V)
    const int out_r0 = rank * kOutRows;
    const bool has_out = out_r0 < V;
    // prologue rows: in-projection rows [16 r, 16 r + 16) and the 48 logical [q | k | vo] outputs [48 r, 48 r + 48) of position 0
    const int in_r0 = rank * (kL / kCS), qkv_r0 = rank * 48;
    const int lo_first = max(qkv_r0, 2 * L), lo_n = max(0, qkv_r0 + 48 - lo_first);          // outputs that also need their "lo" row
    bf * w_in = S.w_out;                                     // [16][768]
    bf * w_qkv = S.w_out + (kL / kCS) * kDm;                 // [48][256]
    bf * w_lo = S.w_ff2;                                     // [<= 48][256] (the W2 tiles are loaded after the prologue)
    float (*hn)[kXL] = reinterpret_cast<float (*)[kXL]>(S.scratch);      // prologue only: LN(seq[0] + pos[0]) (after the hidden states are consumed)
    if (tid == 0) {
        mbar_init(&S.mbar[0], 1); mbar_init(&S.mbar[1], 1); mbar_init(&S.mbar[2], 1); mbar_init(&S.mbar[3], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // operand rows of utterances >= U are never written again: zero them once (their accumulator columns are ignored)
    for (int i = tid; i < (int)(sizeof(S.b_act) + sizeof(S.b_ffh)) / 16; i += kLtThreads) reinterpret_cast<uint4 *>(&S.b_act[0][0][0])[i] = make_uint4(0u, 0u, 0u, 0u);
    tc_before_sync();
    __syncthreads();
    tc_after_sync();
    const uint32_t tmem = S.tmem_slot;
    constexpr uint32_t kTile128 = tc::BM * 128;              // bytes of one packed [128 rows x 64 k] weight tile
    if (tid == 0) {
        mbar_expect_tx(&S.mbar[2], (uint32_t)((kL / kCS) * d + 48 * L + lo_n * L) * 2u);
        bulk_g2s(w_in, (const bf *)p.in_w + (size_t)in_r0 * d, (uint32_t)(kL / kCS) * d * 2u, &S.mbar[2]);
        bulk_g2s(w_qkv, (const bf *)cp.qkvo + (size_t)qkv_r0 * L, 48u * L * 2u, &S.mbar[2]);
        if (lo_n) bulk_g2s(w_lo, (const bf *)cp.qkvo + (size_t)(lo_first + L) * L, (uint32_t)lo_n * L * 2u, &S.mbar[2]);
        mbar_expect_tx(&S.mbar[0], (uint32_t)(kF1Rows * L + L * kF1Rows) * 2u);        // (the W2 half is issued after the prologue)
        // W1 rows [64 r, 64 r + 64) of k tile kt: the upper or lower half of packed tile (r / 2, kt)
        for (int kt = 0; kt < kL / 64; kt++)
            bulk_g2s(reinterpret_cast<unsigned char *>(S.w_ff1) + kt * (kF1Rows * 128),
                     reinterpret_cast<const unsigned char *>(cp.ff1_t) + ((size_t)(rank / 2) * (kL / 64) + kt) * kTile128 + (size_t)(rank % 2) * (kF1Rows * 128),
                     (uint32_t)kF1Rows * 128u, &S.mbar[0]);
    }
    int n_stamp = 0;
    auto stamp = [&]() {
        if (cp.dbg && blockIdx.x == 0 && tid == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); cp.dbg[n_stamp] = t; }
        n_stamp++;
    };
    stamp();
    const bool loop = p.d_step != nullptr || p.utt_step != nullptr;
    auto step_of = [&](int ug) { return p.utt_step ? p.utt_step[ug] : (loop && p.d_step ? *p.d_step : (int)p.step); };

    // ---- prologue: decoder hidden states of the cluster's utterances ----
    float (*hid)[kHS] = reinterpret_cast<float (*)[kHS]>(S.scratch);
    for (int i = tid; i < U * d; i += kLtThreads) {
        const int u = i / d, k = i % d, ug = min(u0 + u, B - 1);
        const float h = p.hidden[(size_t)ug * d + k];
        hid[u][k] = h;
        if (p.hidden_hist && rank == 0 && u0 + u < B) p.hidden_hist[((loop ? (size_t)ug * p.T_total + step_of(ug) : (size_t)ug)) * d + k] = h;
    }
    __syncthreads();
    mbar_wait(&S.mbar[2], 0);
    stamp(); cluster_sync_all(); stamp();       // every CTA of the cluster is running before anyone writes into a peer's shared memory
    // seq[0] = in_proj . hidden + b   (magpie.cpp:1153-1185): 16 rows of this CTA for all U utterances -> every CTA's x1
    tile_dots<kL / kCS, U, kDm, false>(w_in, kL / kCS, &hid[0][0], kHS, [&](int r0, int u, const float (&v)[4]) {
        const float4 bb = *reinterpret_cast<const float4 *>(p.in_b + in_r0 + r0);
        dsmem_st4_all(&S.x1[u][in_r0 + r0], make_float4(v[0] + bb.x, v[1] + bb.y, v[2] + bb.z, v[3] + bb.w));
    });
    stamp(); cluster_sync_all(); stamp();
    // position 0: [q | k | vo] = [Wq; Wk; Wo Wv] . LN(seq + pos[0])  (magpie.cpp:1026-1030, 1501-1503) -> the utterance's owner CTA.
    // vo = hi + lo rows of the folded matrix (model.cu); logical output n < 3L: q (n < L), k (n < 2L), vo (else)
    warp_ln_rows<U>(S.x1, p.pos, p.norm_self, hn, p.eps, L);
    // (48 consecutive logical outputs never straddle q / k / vo inside a group of 4: 48 r and the region bounds are multiples of 4)
    tile_dots<48, U, kL, true>(w_qkv, 48, &hn[0][0], kXL, [&](int r0, int u, const float (&v)[4]) {
        const int n = qkv_r0 + r0;
        float * dst = n < L ? &S.q[n] : (n < 2 * L ? &S.kc[0][n - L] : &S.vc[0][n - 2 * L]);
        dsmem_st4(dst, u, make_float4(v[0], v[1], v[2], v[3]));                  // owner of utterance u = CTA u
    });
    if (lo_n > 0)                                            // CTA-uniform
        tile_dots<48, U, kL, true>(w_lo, lo_n, &hn[0][0], kXL, [&](int r0, int u, const float (&v)[4]) {
            if (r0 < lo_n) dsmem_st4(&S.vlo[lo_first - 2 * L + r0], u, make_float4(v[0], v[1], v[2], v[3]));
        });
    __syncthreads();
    if (tid == 0) {                                          // the prologue rows are consumed: the W2 tiles and codebook 0's output-projection tiles may land
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (int t = 0; t < 2; t++)                          // W2 rows [128 t, 128 t + 128), columns [64 r, 64 r + 64) = packed tile (t, r)
            bulk_g2s(reinterpret_cast<unsigned char *>(S.w_ff2) + t * kTile128,
                     reinterpret_cast<const unsigned char *>(cp.ff2_t) + ((size_t)t * (kF / 64) + rank) * kTile128, kTile128, &S.mbar[0]);
        mbar_expect_tx(&S.mbar[1], has_out ? 4u * kTile128 : 0u);
        if (has_out) bulk_g2s(S.w_out, reinterpret_cast<const unsigned char *>(cp.out_t[0]) + (size_t)rank * 4 * kTile128, 4u * kTile128, &S.mbar[1]);
    }
    stamp(); cluster_sync_all(); stamp();
    // owner: attention over position 0 alone is the identity on vo_0: x1 = (seq + pos[0]) + vo_0; broadcast
    if (owner && tid < L / 4) {
        const float4 a = *reinterpret_cast<const float4 *>(&S.x1[rank][tid * 4]), b = *reinterpret_cast<const float4 *>(p.pos + tid * 4),
                     h = *reinterpret_cast<const float4 *>(&S.vc[0][tid * 4]), l = *reinterpret_cast<const float4 *>(&S.vlo[tid * 4]);
        const float4 c = make_float4(h.x + l.x, h.y + l.y, h.z + l.z, h.w + l.w);       // vo = hi + lo
        *reinterpret_cast<float4 *>(&S.vc[0][tid * 4]) = c;
        dsmem_st4_all(&S.x1[rank][tid * 4], make_float4((a.x + b.x) + c.x, (a.y + b.y) + c.y, (a.z + b.z) + c.z, (a.w + b.w) + c.w));
    }
    stamp(); cluster_sync_all(); stamp();
    mbar_wait(&S.mbar[0], 0);

    bool hit_eos = false;
    uint32_t mma_par = 0;                                    // phase parity of the MMA-completion barrier
    const bool leader = tc::elect_one();                     // (every warp elects; only warp 0's leader issues MMAs)
    const int wq = warp & 3, wg = warp >> 2;                 // epilogues: TMEM lane quarter of this warp; utterances wg, wg + 4, wg + 8
    static_assert(U <= 12, "epilogue utterance mapping");
    const int my_step = owner ? step_of(my_utt) : 0;
    const size_t my_row = loop ? (size_t)my_utt * p.T_total + my_step : (size_t)my_utt;
    const int32_t * forced = (owner && p.forced) ? p.forced + my_row * 8 : nullptr;
    const bool forbid_eos = owner && (p.forbid_eos_all || (p.forbid_eos && p.forbid_eos[my_utt]) || (loop && my_step < p.min_frames));

    for (int cb = 0; cb < 8; cb++) {
        // ---- FF1 (this CTA's 64 hidden rows, all utterances) fused with FF2's partial sums over those 64 columns ----
        warp_ln_image<U>(S.x1, p.norm_ff, S.b_act, p.eps);
        if (cp.dbg_fine) stamp();
        if (warp == 0) {                                     // D[64 x (16 hi | 16 lo)] = W1[64 r .., :] . LN(x1)^T
            // (issued warp-uniformly by one elected lane: descriptors stay in uniform registers, gemm_tc.cuh)
            fence_async_proxy();                             // the operand rows were written through the generic proxy (this CTA's warps)
            tc_after_sync();
            constexpr uint32_t idesc = tc::umma_idesc_bf16(kF1Rows, 32);
            const uint32_t a0 = tc::desc_lo(smem_u32(S.w_ff1)), b0 = tc::desc_lo(smem_u32(&S.b_act[0][0][0]));
            if (leader) {
#pragma unroll
                for (int kt = 0; kt < kL / 64; kt++)
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        tc::umma_lo(tmem, a0 + (uint32_t)(kt * (kF1Rows * 128) + j * 32) / 16, b0 + (uint32_t)(kt * 2 * kBTile + j * 32) / 16, idesc, (kt | j) != 0);
                tc::umma_commit(&S.mbar[3]);
            }
            __syncwarp();
        }
        mbar_wait(&S.mbar[3], mma_par); mma_par ^= 1u;
        tc_after_sync();
        if (cp.dbg_fine) stamp();
        {   // GELU -> this CTA's [16 x 64] operand of the second product.  M = 64: row m on TMEM lane 32 (m / 16) + m % 16
            float v[3], vl[3];
#pragma unroll
            for (int j = 0; j < 3; j++) {
                const int u = wg + 4 * j;
                v[j] = u < U ? tmem_ld1(tmem + ((uint32_t)(32 * wq) << 16) + (uint32_t)u) : 0.0f;
                vl[j] = u < U ? tmem_ld1(tmem + ((uint32_t)(32 * wq) << 16) + (uint32_t)(16 + u)) : 0.0f;
            }
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 3; j++) v[j] += vl[j];
            if (lane < 16) {
                const int m = 16 * wq + lane;
#pragma unroll
                for (int j = 0; j < 3; j++) {
                    const int u = wg + 4 * j;
                    if (u < U) {
                        uint32_t h, l;
                        split_bf(gelu_ggml(v[j], p.gelu_f16), h, l);
                        const int off = tc::swz_offset(u, m);
                        *reinterpret_cast<uint16_t *>(&S.b_ffh[0][off]) = (uint16_t)h;
                        *reinterpret_cast<uint16_t *>(&S.b_ffh[1][off]) = (uint16_t)l;
                    }
                }
            }
        }
        tc_before_sync();
        __syncthreads();
        if (cp.dbg_fine) stamp();
        if (warp == 0) {                                     // partial[n][u] = sum_{k < 64} W2[n][64 r + k] ffh[u][k]: two 128-row tiles
            fence_async_proxy();
            tc_after_sync();
            constexpr uint32_t idesc = tc::umma_idesc_bf16(128, 32);
            const uint32_t a0 = tc::desc_lo(smem_u32(S.w_ff2)), b0 = tc::desc_lo(smem_u32(&S.b_ffh[0][0]));
            if (leader) {
#pragma unroll
                for (int t = 0; t < 2; t++)
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        tc::umma_lo(tmem + 32 + 32 * t, a0 + (uint32_t)(t * kTile128 + j * 32) / 16, b0 + (uint32_t)(j * 32) / 16, idesc, j != 0);
                tc::umma_commit(&S.mbar[3]);
            }
            __syncwarp();
        }
        mbar_wait(&S.mbar[3], mma_par); mma_par ^= 1u;
        tc_after_sync();
        {   // partial sums of output n = 128 t + 32 wq + lane go to the rank that reduces it (n / 16)
            float v[2][3], vl[2][3];
#pragma unroll
            for (int t = 0; t < 2; t++)
#pragma unroll
                for (int j = 0; j < 3; j++) {
                    const int u = wg + 4 * j;
                    v[t][j] = u < U ? tmem_ld1(tmem + ((uint32_t)(32 * wq) << 16) + (uint32_t)(32 + 32 * t + u)) : 0.0f;
                    vl[t][j] = u < U ? tmem_ld1(tmem + ((uint32_t)(32 * wq) << 16) + (uint32_t)(48 + 32 * t + u)) : 0.0f;
                }
            tmem_ld_wait();
#pragma unroll
            for (int t = 0; t < 2; t++) {
                const int n = 128 * t + 32 * wq + lane;
#pragma unroll
                for (int j = 0; j < 3; j++) { const int u = wg + 4 * j; if (u < U) dsmem_st(&S.recv[rank][u][n % kF2Rows], n / kF2Rows, v[t][j] + vl[t][j]); }
            }
        }
        tc_before_sync();
        stamp(); cluster_sync_all(); stamp();
        // ---- reduce the 16 partial sums of this CTA's 16 output rows (rank order), add the residual, gather the layer output as the
        //      next product's operand (hi | lo images) in every CTA ----
        //      Four lanes per (utterance, 4 outputs): each adds 4 source ranks, a fixed shuffle tree joins them (deterministic), each
        //      stores to 4 destination ranks.
        {
            const int item = tid >> 2, part = tid & 3;
            const bool act = item < (kF2Rows / 4) * U;       // (U <= 10: 160 threads = whole warps, the shuffles below see full warps)
            const int itc = act ? item : 0, u = itc % U, g = itc / U, n = rank * kF2Rows + 4 * g;
            float4 a = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (tid < ((kF2Rows / 4) * U * 4 + 31) / 32 * 32) {
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const float4 t = *reinterpret_cast<const float4 *>(&S.recv[4 * part + i][u][4 * g]);
                    a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
                }
#pragma unroll
                for (int o = 1; o <= 2; o <<= 1) {
                    a.x += __shfl_xor_sync(0xffffffffu, a.x, o); a.y += __shfl_xor_sync(0xffffffffu, a.y, o);
                    a.z += __shfl_xor_sync(0xffffffffu, a.z, o); a.w += __shfl_xor_sync(0xffffffffu, a.w, o);
                }
                if (act) {
                    const float4 r = *reinterpret_cast<const float4 *>(&S.x1[u][n]);
                    a.x += r.x; a.y += r.y; a.z += r.z; a.w += r.w;
                    uint32_t h[4], l[4];
                    split_bf(a.x, h[0], l[0]); split_bf(a.y, h[1], l[1]); split_bf(a.z, h[2], l[2]); split_bf(a.w, h[3], l[3]);
                    const int off = tc::swz_offset(u, n & 63);
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        dsmem_st2u(&S.b_act[n >> 6][0][off], 4 * part + i, h[0] | (h[1] << 16), h[2] | (h[3] << 16));
                        dsmem_st2u(&S.b_act[n >> 6][1][off], 4 * part + i, l[0] | (l[1] << 16), l[2] | (l[3] << 16));
                    }
                }
            }
        }
        stamp(); cluster_sync_all(); stamp();
        // ---- out-projection of codebook cb (+bias): 128 rows x U utterances -> the utterance's owner CTA (magpie.cpp:1037-1048) ----
        if (has_out) {                                       // (CTA-uniform)
            if (warp == 0) {
                mbar_wait(&S.mbar[1], (uint32_t)(cb & 1));
                fence_async_proxy();
                tc_after_sync();
                constexpr uint32_t idesc = tc::umma_idesc_bf16(128, 32);
                const uint32_t a0 = tc::desc_lo(smem_u32(S.w_out)), b0 = tc::desc_lo(smem_u32(&S.b_act[0][0][0]));
                if (leader) {
#pragma unroll
                    for (int kt = 0; kt < kL / 64; kt++)
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            tc::umma_lo(tmem + 96, a0 + (uint32_t)(kt * kTile128 + j * 32) / 16, b0 + (uint32_t)(kt * 2 * kBTile + j * 32) / 16, idesc, (kt | j) != 0);
                    tc::umma_commit(&S.mbar[3]);
                }
                __syncwarp();
            }
            const int id = out_r0 + 32 * wq + lane;
            const float bias = id < V ? p.out_b[cb][id] : 0.0f;      // (requested before the wait: off the dependent chain)
            mbar_wait(&S.mbar[3], mma_par); mma_par ^= 1u;
            tc_after_sync();
            if (cp.dbg_fine) stamp();
            float v[3], vl[3];
#pragma unroll
            for (int j = 0; j < 3; j++) {
                const int u = wg + 4 * j;
                v[j] = u < U ? tmem_ld1(tmem + ((uint32_t)(32 * wq) << 16) + (uint32_t)(96 + u)) : 0.0f;
                vl[j] = u < U ? tmem_ld1(tmem + ((uint32_t)(32 * wq) << 16) + (uint32_t)(112 + u)) : 0.0f;
            }
            tmem_ld_wait();
            if (id < V) {
#pragma unroll
                for (int j = 0; j < 3; j++) { const int u = wg + 4 * j; if (u < U) dsmem_st(&S.logits[id], u, (v[j] + vl[j]) + bias); }
            }
            tc_before_sync();
        } else if (warp == 0) mbar_wait(&S.mbar[1], (uint32_t)(cb & 1));
        __syncthreads();                                     // the MMAs have consumed w_out: prefetch the next codebook's tiles
        if (tid == 0 && cb < 7) {
            mbar_expect_tx(&S.mbar[1], has_out ? 4u * kTile128 : 0u);
            if (has_out) bulk_g2s(S.w_out, reinterpret_cast<const unsigned char *>(cp.out_t[cb + 1]) + (size_t)rank * 4 * kTile128, 4u * kTile128, &S.mbar[1]);
        }
        stamp(); cluster_sync_all(); stamp();
        // ---- owner: mask, argmax, top-k sample, feedback gather, attention of position cb + 1, x1 broadcast ----
        if (owner) {
            // forbidden ids: BOS, BOS+2..BOS+7, and EOS while forbid_eos (magpie.cpp:1131-1145, 1243-1248)
            if (tid < 8) {
                const int id = tid == 0 ? p.bos_id : (tid < 7 ? p.bos_id + 1 + tid : (forbid_eos ? p.eos_id : -1));
                if (id >= 0 && id < V) S.logits[id] = -INFINITY;
            }
            __syncthreads();
            // teacher forcing: the code fed to the next position is known, so its table rows are requested ahead of the argmax / sampler
            float g_seq = 0.0f, g_q = 0.0f, g_k = 0.0f, g_v = 0.0f;
            if (forced && cb < 7 && tid < L) {
                const int fed = forced[cb];
                const float * row = cp.qkv_tab + ((size_t)cb * V + fed) * (3 * L);
                g_seq = p.in_table[cb][(size_t)fed * L + tid]; g_q = row[tid]; g_k = row[L + tid]; g_v = row[2 * L + tid];
            }
            if (p.logits)
                for (int i = tid; i < V; i += kLtThreads) p.logits[(my_row * 8 + cb) * V + i] = S.logits[i];
            const int am = block_argmax(S.logits, V, S.red, S.redi);
            if (cp.dbg_fine) stamp();
            int pick = am;
            if (p.temperature >= 0.01f) {
                float uu;
                if (p.uniforms) uu = p.uniforms[my_row * 8 + cb];
                else {
                    uint32_t r4[4];
                    philox4x32_10((uint32_t)my_step, (uint32_t)my_utt, (uint32_t)cb, 0u, (uint32_t)p.seed, (uint32_t)(p.seed >> 32), r4);
                    uu = (float)(r4[0] >> 8) * (1.0f / 16777216.0f);
                }
                SampView sv;
                sv.logits = S.logits; sv.sel_v = S.scratch; sv.srt_v = sv.sel_v + kV;
                sv.sel_i = reinterpret_cast<uint16_t *>(sv.srt_v + kV); sv.srt_i = sv.sel_i + kV;
                sv.hist = reinterpret_cast<unsigned *>(sv.srt_i + kV); sv.misc = reinterpret_cast<int *>(sv.hist + 256);
                pick = block_sample_top_k(sv, V, p.temperature, p.top_k, uu);
            }
            if (cp.dbg_fine) stamp();
            hit_eos = hit_eos || pick == p.eos_id || am == p.eos_id;
            if (tid == 0) {
                p.argmax[my_row * 8 + cb] = am;
                p.sampled[my_row * 8 + cb] = pick;
                if (p.next_codes) p.next_codes[my_utt * 8 + cb] = forced ? forced[cb] : pick;
                if (cb == 7) {
                    if (p.eos_flag) p.eos_flag[my_utt] = hit_eos ? 1 : 0;
                    if (p.done_step && hit_eos && p.done_step[my_utt] < 0) p.done_step[my_utt] = my_step;
                }
            }
            if (cb < 7) {
                // seq[cb+1] = in_proj . E_cb[code] + b and its [q | k | vo] row, both tabulated at load (model.cu); no 1/8 scale (magpie.cpp:1285-1291)
                const int fed = forced ? forced[cb] : pick;
                if (tid < L) {
                    if (!forced) {
                        const float * row = cp.qkv_tab + ((size_t)cb * V + fed) * (3 * L);
                        g_seq = p.in_table[cb][(size_t)fed * L + tid]; g_q = row[tid]; g_k = row[L + tid]; g_v = row[2 * L + tid];
                    }
                    S.seq[tid] = g_seq;
                    S.q[tid] = g_q;
                    S.kc[cb + 1][tid] = g_k;
                    S.vc[cb + 1][tid] = g_v;
                }
                __syncthreads();
                if (warp <= cb + 1) {
                    float s = 0.0f;
                    for (int i = lane; i < L; i += 32) s = fmaf(S.kc[warp][i], S.q[i], s);
                    s = warp_sum(s);
                    if (lane == 0) S.scores[warp] = s * att_scale;
                }
                __syncthreads();
                if (tid < L / 4) {
                    float mxs = S.scores[0];
                    for (int j = 1; j <= cb + 1; j++) mxs = fmaxf(mxs, S.scores[j]);
                    float e[8], sum = 0.0f;
#pragma unroll
                    for (int j = 0; j < 8; j++) { e[j] = (j <= cb + 1) ? expf(S.scores[j] - mxs) : 0.0f; sum += e[j]; }
                    const float inv = 1.0f / sum;
                    float4 o = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
                    for (int j = 0; j < 8; j++)
                        if (j <= cb + 1) {
                            const float4 vv = *reinterpret_cast<const float4 *>(&S.vc[j][tid * 4]);
                            const float pj = e[j] * inv;
                            o.x = fmaf(pj, vv.x, o.x); o.y = fmaf(pj, vv.y, o.y); o.z = fmaf(pj, vv.z, o.z); o.w = fmaf(pj, vv.w, o.w);
                        }
                    const float4 sq = *reinterpret_cast<const float4 *>(&S.seq[tid * 4]), ps = *reinterpret_cast<const float4 *>(p.pos + (cb + 1) * L + tid * 4);
                    dsmem_st4_all(&S.x1[rank][tid * 4], make_float4((sq.x + ps.x) + o.x, (sq.y + ps.y) + o.y, (sq.z + ps.z) + o.z, (sq.w + ps.w) + o.w));
                }
            }
        }
        stamp(); cluster_sync_all(); stamp();
    }
    tc_before_sync();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols));
}

template <int U> bool launch_cl(const ClParams & cp, int n_clusters, cudaStream_t stream) {
    static_assert(sizeof(ClSmem<U>) + 1024 <= 227 * 1024, "lt_cluster shared memory");
    static_assert(U <= 10, "hidden-state scratch holds 10 utterances");
    static DeviceOnce attr_done;
    int dev = 0;
    MGB_CUDA_TRY(cudaGetDevice(&dev));
    const size_t smem = sizeof(ClSmem<U>) + 1024;
    if (!attr_done.done(dev)) {
        MGB_CUDA_TRY(cudaFuncSetAttribute(lt_cluster_kernel<U>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        MGB_CUDA_TRY(cudaFuncSetAttribute(lt_cluster_kernel<U>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        attr_done.set(dev);
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n_clusters * kCS); cfg.blockDim = dim3(kLtThreads); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = kCS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    MGB_CUDA_TRY(cudaLaunchKernelEx(&cfg, lt_cluster_kernel<U>, cp));
    MGB_LAUNCH_CHECK();
    return true;
}

// co-resident 16-CTA clusters of the kernel on this device (0 = cannot be scheduled)
template <int U> int max_clusters() {
    int ncl = 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(kCS * 8); cfg.blockDim = dim3(kLtThreads); cfg.dynamicSmemBytes = sizeof(ClSmem<U>) + 1024;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = kCS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    const bool ok = cudaFuncSetAttribute(lt_cluster_kernel<U>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(ClSmem<U>) + 1024)) == cudaSuccess &&
                    cudaFuncSetAttribute(lt_cluster_kernel<U>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
                    cudaOccupancyMaxActiveClusters(&ncl, lt_cluster_kernel<U>, &cfg) == cudaSuccess;
    cudaGetLastError();
    return ok ? ncl : 0;
}

}  // namespace

// utterances per cluster for a batch of B (0 = not supported: the caller keeps lt_batch / the per-utterance kernels)
int lt_cluster_plan(const Model & m, int B) {
    const mgb_hparams & hp = m.hp;
    if (getenv("MGB_NO_LT_CLUSTER") != nullptr || getenv("MGB_LT_STREAM") != nullptr || getenv("MGB_NO_LT_BATCH") != nullptr) return 0;
    if (m.precision != MGB_PREC_BF16 || !m.lt_in_table[0] || !m.lt_qkvo || !m.lt_qkv_tab || !m.lt_ff1.tiles || !m.lt_ff2.tiles || B < 4) return 0;
    if (hp.lt_dim != kL || hp.lt_ffn_dim != kF || hp.d_model != kDm || hp.vocab_per_cb > kV || hp.vocab_per_cb < 16) return 0;
    static std::atomic<int> cap[64];                     // co-resident clusters per device (+1; 0 = not probed yet)
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cap[dev & 63] == 0) cap[dev & 63] = 1 + std::min(max_clusters<10>(), max_clusters<4>());
    const int ncl = cap[dev & 63] - 1;
    if (getenv("MGB_LT_DBG")) fprintf(stderr, "lt_cluster_plan: %d co-resident 16-CTA clusters, B = %d\n", ncl, B);
    for (int U : {1, 2, 3, 4, 5, 6, 7, 8, 10})
        if ((B + U - 1) / U <= ncl) return U;            // the fewest utterances per cluster that still runs as one wave
    return 0;
}

bool launch_lt_cluster(const Model & m, const lt::LtParams & p, int U, cudaStream_t stream) {
    ClParams cp;
    cp.p = p; cp.qkvo = m.lt_qkvo; cp.qkv_tab = m.lt_qkv_tab; cp.ff1_t = m.lt_ff1.tiles; cp.ff2_t = m.lt_ff2.tiles;
    for (int cb = 0; cb < 8; cb++) { cp.out_t[cb] = m.lt_out_w[cb].tiles; if (!cp.out_t[cb]) { set_error("lt_cluster: output projection tiles missing"); return false; } }
    static unsigned long long * dbg = nullptr;
    if (getenv("MGB_LT_DBG") && !dbg) { MGB_CUDA_TRY(cudaMalloc((void **)&dbg, 256 * 8)); MGB_CUDA_TRY(cudaMemset(dbg, 0, 256 * 8)); }
    cp.dbg = dbg; cp.dbg_fine = getenv("MGB_LT_DBG") && atoi(getenv("MGB_LT_DBG")) >= 2;
    if (dbg && getenv("MGB_LT_DBG_DUMP")) {
        unsigned long long h[256];
        cudaDeviceSynchronize();
        cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost);
        for (int i = 1; i < 256 && h[i]; i++) fprintf(stderr, "lt_cluster stamp %3d  +%6llu ns  (%s)\n", i, h[i] - h[i - 1], (i & 1) ? "work" : "sync");
    }
    const int ncl = (p.B + U - 1) / U;
    switch (U) {
        case 1: return launch_cl<1>(cp, ncl, stream);
        case 2: return launch_cl<2>(cp, ncl, stream);
        case 3: return launch_cl<3>(cp, ncl, stream);
        case 4: return launch_cl<4>(cp, ncl, stream);
        case 5: return launch_cl<5>(cp, ncl, stream);
        case 6: return launch_cl<6>(cp, ncl, stream);
        case 7: return launch_cl<7>(cp, ncl, stream);
        case 8: return launch_cl<8>(cp, ncl, stream);
        case 10: return launch_cl<10>(cp, ncl, stream);
        default: set_error("lt_cluster: unsupported utterances per cluster"); return false;
    }
}

}  // namespace mgb
