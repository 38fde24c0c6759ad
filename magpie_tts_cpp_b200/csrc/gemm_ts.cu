// Token-stationary tcgen05 GEMM for the batched decoder step (<= 64 tokens): Y[t][n] = sum_k X[t][k] W[n][k].
//
// gemm_tc.cu makes the weight matrix the MMA "A" operand (128 output features per CTA), which caps a 768-row matrix at 6 CTAs
// (x a 4-way cluster split-K = 24): at 64 tokens such a launch is bound by how fast ONE SM ingests its tiles, with 120+ SMs
// idle (round-1 ncu: 117-320 GB/s per launch).  Here the roles are swapped: the <= 64 TOKENS are the MMA M side
// (tcgen05.mma M = 64: accumulator rows on TMEM lanes 32 (m / 16) + m % 16) and every CTA owns a thin slice of n = 8 / 16 / 32
// weight rows as the MMA N side, so a GEMM spreads over 96-144 CTAs, each streaming only its n x K weight bytes plus the
// (L2-resident) activation image.  The K = 768 GEMMs need no cross-CTA reduction; the FFN's second GEMM (K = 3072) splits k over a
// 4-CTA cluster and reduce-scatters the partial accumulators through DSMEM in rank order: results are deterministic either way.
// Operands are the shared-memory images gemm_tc.cu uses (SWIZZLE_128B K-major tiles of 64 k): an 8-row-aligned run of n rows inside
// a packed 128-row weight tile is itself a valid n-row tile, so no second weight LAYOUT is needed; the activations arrive as ONE f16
// image against f16 twins of the bf16 weight images (default) or as bf16 hi | lo image pairs (MGB_ACT_F16=0), written by the producing
// kernel's epilogue.  What the round-2 timeline (profiles/r2_b64_step_timeline.txt) taught this file:
//   * the producer and MMA roles run their loops warp-uniformly with one elected lane (a divergent single lane pays ~60 cycles per
//     tcgen05.mma / bulk copy for register -> uniform-register moves);
//   * epilogues use all 8 x 32 epilogue lanes (two warps per TMEM lane quarter, lanes 16..31 take half of their row's columns), fetch
//     their global operands before waiting for the accumulator, and store operand images in whole 16 / 32-byte pieces;
//   * a LayerNorm in front of a GEMM is folded THROUGH it: the producer emits x .* w and per-slice row statistics, the epilogue applies
//     (acc - mean * csum[n]) * rstd with the statistics added while the MMAs run.
// Replaces ggml_mul_mat at src/magpie.cpp:3415, 3472, 1796, 1805 for the batched step.
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "gemm_tc.cuh"
#include "kernels.cuh"

namespace mgb {

namespace {

using bf = __nv_bfloat16;

constexpr int kTsThreads = 320;          // warp 0 producer, warp 1 TMEM + MMA issue, warps 2..9 epilogue
constexpr int kXTile = 64 * 128;         // one 64-token x 64-k bf16 image

// MGB_TS_DBG: globaltimer stamps of CTA (0, 0) of every launch (entry | dependency wait over | first stage landed | last MMA issued |
// accumulator complete | epilogue done | kernel id), a ring of 1024 launches; dumped by tools/ts_timeline.py through MGB_TS_DBG_DUMP
__device__ unsigned long long g_ts_dbg[1024 * 8];
__device__ unsigned g_ts_dbg_n;

struct TsEpi {
    int N, M;
    const float * res; int ldr;
    float * Y; int ldy;
    int gelu_f16;
    int n_q, dkv; bf * kdst; bf * vdst; const int32_t * tok_slot;
    bf * pk_hi; bf * pk_lo;
    int dbg;
    int x_f16, pack_f16;                                           // f16 activation images (gemm_tc.cuh pack_act2): operand / epilogue output
    const float * next_w; float * stats_out;                       // TS_RES producer of a folded LayerNorm (kernels.cuh)
    const float * ln_stats; int ln_slices; const float * ln_csum; float eps; int K;     // TS_QKV consumer
};

enum { TS_QKV = 1, TS_RES = 2, TS_GELU_PACK = 3 };

template <int NC> __device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&v)[NC]);
template <> __device__ __forceinline__ void tmem_ld_cols<8>(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
template <> __device__ __forceinline__ void tmem_ld_cols<16>(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// NC = weight rows (output features) per CTA.  SPLIT > 1 (wide K: the FFN's second GEMM, K = 3072): a cluster of SPLIT CTAs
// (grid y) shares one slice of NC rows and divides the k tiles; at 64 tokens a CTA is bound by how fast its SM ingests the
// activation operand (786 KB for K = 3072), so the split shortens the critical path.  Ranks > 0 push their accumulator into rank
// 0's shared memory through DSMEM; rank 0 adds them in rank order (deterministic) and runs the epilogue.
template <int NC, int EPI, int SPLIT, int kTsStages>
__global__ void __launch_bounds__(kTsThreads, 1) ts_linear_kernel(const bf * Wt, const bf * Xhi, const bf * Xlo, int KT_all, const TsEpi e) {
    constexpr int kWTile = NC * 128;
    const int kXBytes = e.x_f16 ? kXTile : 2 * kXTile;                  // activation bytes per stage: one f16 image or bf16 hi | lo
    const int kStage = kXBytes + (kWTile < 1024 ? 1024 : kWTile);
    extern __shared__ unsigned char ts_smem[];
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");      // let the dependent kernel start its own prefetching
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char * tiles = reinterpret_cast<unsigned char *>(((uintptr_t)ts_smem + 1023) & ~(uintptr_t)1023);
    uint64_t * bars = reinterpret_cast<uint64_t *>(tiles + kTsStages * kStage);
    uint64_t * full = bars, * empty = bars + kTsStages, * acc_full = bars + 2 * kTsStages;
    uint32_t * tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * kTsStages + 1);
    __shared__ float2 s_part[4][64];                                     // TS_QKV with a folded LayerNorm: partial row statistics (see the epilogue warps)
    const int n0 = blockIdx.x * NC;                                      // first output feature of this CTA
    // epilogue: two warps per TMEM lane quarter, each takes half of the CTA's columns (NC >= 16; with M = 64 only lanes 0..15 of a
    // warp hold rows, so the epilogues -- GELU + packing of up to 32 columns per row -- were serial tails of 3.5-4.2 us on 64 threads)
    constexpr int NH = NC >= 16 ? NC / 2 : NC;
    const int half = warp >= 6 ? 1 : 0;
    const bool epi_on = NC >= 16 || half == 0;
    const int c0 = NC >= 16 ? half * NH : 0;
    __shared__ unsigned dbg_slot;
    const bool dbg = e.dbg && blockIdx.x == 0 && blockIdx.y == 0;
    auto stamp = [&](int i) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); g_ts_dbg[(size_t)dbg_slot * 8 + i] = t; };
    if (dbg && threadIdx.x == 0) { dbg_slot = atomicAdd(&g_ts_dbg_n, 1u) % 1024u; stamp(0); g_ts_dbg[(size_t)dbg_slot * 8 + 6] = NC * 100 + EPI * 10 + SPLIT; }
    const int rank = SPLIT > 1 ? (int)blockIdx.y : 0;
    constexpr int CR = NC / SPLIT;                                       // split-K: columns a rank finishes (reduce-scatter epilogue)
    constexpr int NE = SPLIT > 1 ? CR : NH;                              // columns per epilogue thread
    const int nb = n0 + (SPLIT > 1 ? CR * rank : c0);                    // first output feature of this thread
    float4 pre0 = make_float4(0.0f, 0.0f, 0.0f, 0.0f), pre1 = pre0;      // epilogue operands fetched ahead of the accumulator (see the epilogue warps)
    int pre_slot = 0;
    static_assert(NE == 8 || EPI == TS_GELU_PACK, "prefetched epilogue operands: 4 columns per lane");
    const int kt_lo = rank * KT_all / SPLIT, KT = (rank + 1) * KT_all / SPLIT - kt_lo;      // this CTA's k tiles [kt_lo, kt_lo + KT)
    float * xbuf = reinterpret_cast<float *>(tiles + kTsStages * kStage + 256);             // rank 0: [SPLIT - 1][64][NC] partial accumulators
    // every CTA reads the SAME activation tiles: with all of them walking k = 0, 1, 2, ... in lock step the 96-144 SMs would hit
    // the same few L2 slices at the same time, so CTA c starts its walk at k tile (5 c) mod KT (a fixed order per CTA: deterministic)
    const int kt_first = (int)((blockIdx.x * 5u) % (unsigned)KT);
    const int KTW = KT_all;                                              // k tiles per 128-row weight tile row
    if (threadIdx.x == 0) {
        for (int s = 0; s < kTsStages; s++) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
        tc::mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(tmem_slot)), "n"(NC < 32 ? 32 : NC));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    // Producer and MMA roles run their loops WARP-UNIFORMLY (all lanes poll the barriers, one elected lane issues): addresses and
    // descriptors then live in uniform registers.  Issued from a divergent single lane every tcgen05.mma cost ~60 cycles of register ->
    // uniform-register moves: 96 MMAs = 3 us of the 4.6 us this loop took per GEMM (profiles/r2_b64_step_timeline.txt).
    if (warp == 0) {
        const bool leader = tc::elect_one();
        // weight rows n0 .. n0 + NC - 1 of k tile kt: an 8-row-aligned run inside the packed 128-row tile (n0 / 128, kt)
        const unsigned char * wsrc = reinterpret_cast<const unsigned char *>(Wt) + ((size_t)(n0 / tc::BM) * KTW + kt_lo) * (tc::BM * 128) + (size_t)(n0 % tc::BM) * 128;
        const unsigned char * hsrc = reinterpret_cast<const unsigned char *>(Xhi) + (size_t)kt_lo * kXTile, * lsrc = reinterpret_cast<const unsigned char *>(Xlo) + (size_t)kt_lo * kXTile;
        // programmatic dependent launch: the weights do not depend on the preceding kernel (which produces the activations)
        const int npre = KT < kTsStages ? KT : kTsStages;
        if (leader)
            for (int kt = 0; kt < npre; kt++) {
                int kk = kt + kt_first; kk = kk >= KT ? kk - KT : kk;
                tc::mbar_expect_tx(&full[kt], kXBytes + kWTile);
                tc::bulk_g2s(tiles + kt * kStage + kXBytes, wsrc + (size_t)kk * (tc::BM * 128), kWTile, &full[kt]);
            }
        asm volatile("griddepcontrol.wait;" ::: "memory");
        if (dbg && leader) stamp(1);
        for (int kt = 0; kt < KT; kt++) {
            const int s = kt % kTsStages;
            int kk = kt + kt_first; kk = kk >= KT ? kk - KT : kk;
            unsigned char * st = tiles + s * kStage;
            if (kt >= npre) tc::mbar_wait(&empty[s], ((kt / kTsStages) & 1) ^ 1);
            if (leader) {
                if (kt >= npre) {
                    tc::mbar_expect_tx(&full[s], kXBytes + kWTile);
                    tc::bulk_g2s(st + kXBytes, wsrc + (size_t)kk * (tc::BM * 128), kWTile, &full[s]);
                }
                tc::bulk_g2s(st, hsrc + (size_t)kk * kXTile, kXTile, &full[s]);
                if (!e.x_f16) tc::bulk_g2s(st + kXTile, lsrc + (size_t)kk * kXTile, kXTile, &full[s]);
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = tc::umma_idesc_bf16(64, NC);          // D[64 tokens x NC] += X[64 x 16] . W[NC x 16]^T
        constexpr uint32_t idesc_h = tc::umma_idesc_f16(64, NC);         // ... activations as one f16 image against the f16 twin of the weights
        const bool leader = tc::elect_one();
        const uint32_t tiles_d = tc::desc_lo(tc::smem_u32(tiles));       // descriptor address units are 16 bytes
        for (int kt = 0; kt < KT; kt++) {
            const int s = kt % kTsStages;
            tc::mbar_wait(&full[s], (kt / kTsStages) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t h0 = tiles_d + (uint32_t)s * (uint32_t)(kStage >> 4), l0 = h0 + (kXTile >> 4), w0 = h0 + (uint32_t)(kXBytes >> 4);
            if (leader) {
                if (dbg && kt == 0) stamp(2);
                if (e.x_f16) {
#pragma unroll
                    for (int j = 0; j < tc::BK / 16; j++) tc::umma_lo(tmem_base, h0 + 2 * j, w0 + 2 * j, idesc_h, (kt | j) != 0);
                } else {
#pragma unroll
                    for (int j = 0; j < tc::BK / 16; j++) {              // 16 k-elements = 32 bytes = 2 descriptor units inside the swizzle row
                        tc::umma_lo(tmem_base, h0 + 2 * j, w0 + 2 * j, idesc, (kt | j) != 0);
                        tc::umma_lo(tmem_base, l0 + 2 * j, w0 + 2 * j, idesc, 1u);
                    }
                }
                if (kt + kTsStages < KT) tc::umma_commit(&empty[s]);      // only a stage that is refilled needs its release
            }
        }
        if (leader) { tc::umma_commit(acc_full); if (dbg) stamp(3); }
        __syncwarp();
    }
    if (warp >= 2) {
        asm volatile("griddepcontrol.wait;" ::: "memory");
        if ((EPI == TS_QKV || EPI == TS_GELU_PACK) && e.ln_stats) {
            // LayerNorm folded through this GEMM: while the MMAs run, the 256 epilogue threads add the row statistics of the e.ln_slices column
            // slices the preceding FF2 wrote: 4 threads per row, a contiguous quarter of the slices each (fixed order: deterministic)
            const int t = (int)threadIdx.x - 64, m = t & 63, part = t >> 6, per = (e.ln_slices + 3) >> 2;
            const int s_lo = part * per, s_hi = min(e.ln_slices, s_lo + per);
            float s1 = 0.0f, s2 = 0.0f;
#pragma unroll 4
            for (int sl = s_lo; sl < s_hi; sl++) {
                const float2 st = *reinterpret_cast<const float2 *>(e.ln_stats + ((size_t)sl * 64 + m) * 2);
                s1 += st.x; s2 += st.y;
            }
            s_part[part][m] = make_float2(s1, s2);
            asm volatile("bar.sync 1, 256;" ::: "memory");
        }
        // the epilogue's global operands (folded-LayerNorm column sums / residual row, next LayerNorm weight, cache slot) are requested
        // BEFORE the wait for the accumulator, so that their L2 round trip is behind the main loop instead of behind the last MMA
        {
            const int q_ = warp & 3, lh_ = lane >> 4, m_ = 16 * q_ + (lane & 15), nc_ = nb + lh_ * (NE / 2);
            if ((SPLIT > 1 ? half == 0 : epi_on) && m_ < e.M && nb < e.N) {
                if constexpr (EPI == TS_QKV) {
                    if (e.ln_stats) pre0 = *reinterpret_cast<const float4 *>(e.ln_csum + nc_);
                    if (nc_ >= e.n_q) pre_slot = e.tok_slot[m_];
                } else if constexpr (EPI == TS_RES) {
                    pre0 = *reinterpret_cast<const float4 *>(e.res + (size_t)m_ * e.ldr + nc_);
                    if (e.next_w) pre1 = *reinterpret_cast<const float4 *>(e.next_w + nc_);
                } else {
                    if (e.ln_stats) { pre0 = *reinterpret_cast<const float4 *>(e.ln_csum + nc_); pre1 = *reinterpret_cast<const float4 *>(e.ln_csum + nc_ + 4); }
                }
            }
        }
        tc::mbar_wait(acc_full, 0);
        if (dbg && threadIdx.x == 64) stamp(4);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3;                      // TMEM lane quarter of this warp; M = 64: rows 16 q .. 16 q + 15 on its lanes 0..15
        const int m = 16 * q + lane;
        if (SPLIT > 1) {
            // split-K: REDUCE-SCATTER.  Rank r finishes columns [CR r, CR r + CR) of the slice: every rank sends the other ranks their
            // column blocks of its partial accumulator through DSMEM (this thread: 16 columns = the blocks of ranks 2 half, 2 half + 1),
            // so the epilogue's work (residual, stores, next operand, statistics) is spread over the cluster instead of rank 0 alone.
            static_assert(SPLIT == 1 || (SPLIT == 4 && NC == 32), "split-K epilogue is written for 4 ranks x 8 columns");
            uint32_t v[16];
            tmem_ld_cols<16>(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(16 * half), v);
            if (lane < 16) {
#pragma unroll
                for (int dd = 0; dd < 2; dd++) {
                    const int dest = 2 * half + dd;
                    if (dest != rank) {
                        const int slot = rank < dest ? rank : rank - 1;          // position among the destination's 3 sources (rank order)
                        const uint32_t local = tc::smem_u32(xbuf + ((size_t)slot * 64 + m) * CR);
                        uint32_t remote;
                        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(dest));
#pragma unroll
                        for (int j = 0; j < CR / 4; j++)
                            asm volatile("st.shared::cluster.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(remote + j * 16), "r"(v[8 * dd + 4 * j]), "r"(v[8 * dd + 4 * j + 1]),
                                         "r"(v[8 * dd + 4 * j + 2]), "r"(v[8 * dd + 4 * j + 3]) : "memory");
                    }
                }
            }
            __syncwarp();
        }
    }
    if (SPLIT > 1) {
        __syncwarp();                                // (the producer / MMA lanes rejoin their warps before the aligned barrier)
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    if (warp >= 2 && (SPLIT > 1 ? half == 0 : epi_on)) {
        // M = 64 puts the accumulator rows on lanes 0..15 of each warp: lanes 16..31 take the upper half of their row's columns over
        // (one shuffle per column), so every lane finishes NE2 columns of row m
        constexpr int NE2 = NE / 2;
        static_assert(NE2 % 4 == 0, "epilogue works in units of 4 columns");
        const int q = warp & 3, lh = lane >> 4, ml = lane & 15;
        const int m = 16 * q + ml;
        const int nc = nb + lh * NE2;                // first output feature of this lane
        uint32_t v[NE];
        tmem_ld_cols<NE>(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(SPLIT > 1 ? CR * rank : c0), v);
        float y[NE2];
#pragma unroll
        for (int j = 0; j < NE2; j++) {
            const uint32_t up = __shfl_sync(0xffffffffu, v[NE2 + j], ml);
            y[j] = __uint_as_float(lh ? up : v[j]);
        }
        if (m < e.M && nb < e.N) {
            if (SPLIT > 1) {                         // partial sums added in rank order (deterministic); this rank's own comes from TMEM
                float own[NE2];
#pragma unroll
                for (int j = 0; j < NE2; j++) { own[j] = y[j]; y[j] = 0.0f; }
#pragma unroll
                for (int r = 0; r < SPLIT; r++) {
                    if (r == rank) {
#pragma unroll
                        for (int j = 0; j < NE2; j++) y[j] = r == 0 ? own[j] : y[j] + own[j];
                    } else {
                        const float4 * pr = reinterpret_cast<const float4 *>(xbuf + ((size_t)(r < rank ? r : r - 1) * 64 + m) * CR + lh * NE2);
#pragma unroll
                        for (int j = 0; j < NE2 / 4; j++) {
                            const float4 t = pr[j];
                            if (r == 0) { y[4 * j] = t.x; y[4 * j + 1] = t.y; y[4 * j + 2] = t.z; y[4 * j + 3] = t.w; }
                            else { y[4 * j] += t.x; y[4 * j + 1] += t.y; y[4 * j + 2] += t.z; y[4 * j + 3] += t.w; }
                        }
                    }
                }
            }
            if constexpr (EPI == TS_QKV) {
                if (e.ln_stats) {                    // LayerNorm folded through this GEMM: y = (acc - mean * csum[n]) * rstd
                    const float s1 = ((s_part[0][m].x + s_part[1][m].x) + s_part[2][m].x) + s_part[3][m].x;
                    const float s2 = ((s_part[0][m].y + s_part[1][m].y) + s_part[2][m].y) + s_part[3][m].y;
                    const float mean = s1 / (float)e.K, var = fmaxf(s2 / (float)e.K - mean * mean, 0.0f);
                    const float rstd = 1.0f / sqrtf(var + e.eps);
                    const float cs[4] = {pre0.x, pre0.y, pre0.z, pre0.w};
#pragma unroll
                    for (int j = 0; j < NE2; j++) y[j] = (y[j] - mean * cs[j]) * rstd;
                }
                if (nc < e.n_q) {
                    float4 * dst = reinterpret_cast<float4 *>(e.Y + (size_t)m * e.ldy + nc);
#pragma unroll
                    for (int j = 0; j < NE2 / 4; j++) dst[j] = make_float4(y[4 * j], y[4 * j + 1], y[4 * j + 2], y[4 * j + 3]);
                } else {
                    const int cc = nc - e.n_q;
                    bf * dst = (cc < e.dkv ? e.kdst + cc : e.vdst + (cc - e.dkv)) + (size_t)pre_slot * e.dkv;
#pragma unroll
                    for (int j = 0; j < NE2 / 4; j++) {
                        uint32_t w[2];
#pragma unroll
                        for (int p = 0; p < 2; p++)
                            w[p] = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(y[4 * j + 2 * p])) |
                                   ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(y[4 * j + 2 * p + 1])) << 16);
                        reinterpret_cast<uint2 *>(dst)[j] = make_uint2(w[0], w[1]);
                    }
                }
            } else if constexpr (EPI == TS_RES) {
                const float4 r[1] = {pre0};
                float4 * dst = reinterpret_cast<float4 *>(e.Y + (size_t)m * e.ldy + nc);
#pragma unroll
                for (int j = 0; j < NE2 / 4; j++) {
                    y[4 * j] += r[j].x; y[4 * j + 1] += r[j].y; y[4 * j + 2] += r[j].z; y[4 * j + 3] += r[j].w;
                    dst[j] = make_float4(y[4 * j], y[4 * j + 1], y[4 * j + 2], y[4 * j + 3]);
                }
                if (e.next_w) {                      // the next GEMM's operand: (y .* w) as operand image(s) + this slice's row statistics
                    float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
                    for (int j = 0; j < NE2; j++) { s1 += y[j]; s2 = fmaf(y[j], y[j], s2); }
                    // (the row's two lanes: lower + upper half of the columns; both lanes of a pair are in this branch)
                    const unsigned am = __activemask();
                    const float o1 = __shfl_xor_sync(am, s1, 16), o2 = __shfl_xor_sync(am, s2, 16);
                    if (lh == 0)
                        *reinterpret_cast<float2 *>(e.stats_out + ((size_t)(SPLIT > 1 ? blockIdx.x * SPLIT + rank : blockIdx.x * (NC / NE) + half) * 64 + m) * 2) = make_float2(s1 + o1, s2 + o2);
                    // operand image stores: a lane's row is 128 bytes away from its neighbour's, so nothing coalesces across the warp and an
                    // 8-byte store leaves a quarter-written sector that the consuming GEMM's bulk copies then read slowly (its main loop:
                    // 2.6 -> 3.7 us); the row's lower lane collects the pair's words and stores whole 16-byte chunks
                    static_assert(NE2 == 4, "residual epilogue with a following operand image: 8 columns per lane pair");
                    {
                        uint32_t h[2], l[2];
#pragma unroll
                        for (int p = 0; p < 2; p++)
                            tc::pack_act2(y[2 * p] * (p ? pre1.z : pre1.x), y[2 * p + 1] * (p ? pre1.w : pre1.y), e.pack_f16 != 0, h[p], l[p]);
                        const uint32_t h2 = __shfl_xor_sync(am, h[0], 16), h3 = __shfl_xor_sync(am, h[1], 16);
                        const uint32_t l2 = __shfl_xor_sync(am, l[0], 16), l3 = __shfl_xor_sync(am, l[1], 16);
                        if (lh == 0) {
                            const size_t off = (size_t)(nb >> 6) * kXTile + tc::swz_offset(m, nb & 63);
                            *reinterpret_cast<uint4 *>(reinterpret_cast<unsigned char *>(e.pk_hi) + off) = make_uint4(h[0], h[1], h2, h3);
                            if (!e.pack_f16) *reinterpret_cast<uint4 *>(reinterpret_cast<unsigned char *>(e.pk_lo) + off) = make_uint4(l[0], l[1], l2, l3);
                        }
                    }
                }
            } else {                                 // GELU + operand image(s) for the next GEMM (k tile n / 64, 16-byte chunk (n % 64) / 8)
                // (whole 32-byte sectors per store, see above: the pair's 16 columns are two 16-byte chunks whose swizzled places c ^ (m & 7)
                //  and (c + 1) ^ (m & 7) share a sector, swapped when m is odd; the lower lane stores both)
                static_assert(NE2 == 8, "GELU epilogue: 16 columns per lane pair");
                if (e.ln_stats) {                    // LayerNorm folded through this GEMM (the folded cross-attention emitted x .* w + statistics)
                    const float s1 = ((s_part[0][m].x + s_part[1][m].x) + s_part[2][m].x) + s_part[3][m].x;
                    const float s2 = ((s_part[0][m].y + s_part[1][m].y) + s_part[2][m].y) + s_part[3][m].y;
                    const float mean = s1 / (float)e.K, var = fmaxf(s2 / (float)e.K - mean * mean, 0.0f);
                    const float rstd = 1.0f / sqrtf(var + e.eps);
                    const float cs[8] = {pre0.x, pre0.y, pre0.z, pre0.w, pre1.x, pre1.y, pre1.z, pre1.w};
#pragma unroll
                    for (int j = 0; j < NE2; j++) y[j] = (y[j] - mean * cs[j]) * rstd;
                }
                {
                    const unsigned am = __activemask();
                    uint32_t h[4], l[4], hp[4], lp[4];
#pragma unroll
                    for (int p = 0; p < 4; p++)
                        tc::pack_act2(gelu_ggml_fast(y[2 * p], e.gelu_f16), gelu_ggml_fast(y[2 * p + 1], e.gelu_f16), e.pack_f16 != 0, h[p], l[p]);
#pragma unroll
                    for (int p = 0; p < 4; p++) { hp[p] = __shfl_xor_sync(am, h[p], 16); lp[p] = __shfl_xor_sync(am, l[p], 16); }
                    if (lh == 0) {
                        const int c = (nb & 63) >> 3, sw = m & 7;                      // chunk index of the pair's lower 8 columns (even)
                        const size_t off = (size_t)(nb >> 6) * kXTile + (size_t)m * 128 + (size_t)(((c ^ sw) & 6) << 4);
                        const bool swp = (sw & 1) != 0;                                // odd rows: the upper chunk comes first in the sector
                        uint4 * dh = reinterpret_cast<uint4 *>(reinterpret_cast<unsigned char *>(e.pk_hi) + off);
                        const uint4 a = make_uint4(h[0], h[1], h[2], h[3]), b = make_uint4(hp[0], hp[1], hp[2], hp[3]);
                        dh[0] = swp ? b : a; dh[1] = swp ? a : b;
                        if (!e.pack_f16) {
                            uint4 * dl = reinterpret_cast<uint4 *>(reinterpret_cast<unsigned char *>(e.pk_lo) + off);
                            const uint4 a2 = make_uint4(l[0], l[1], l[2], l[3]), b2 = make_uint4(lp[0], lp[1], lp[2], lp[3]);
                            dl[0] = swp ? b2 : a2; dl[1] = swp ? a2 : b2;
                        }
                    }
                }
            }
        }
    }
    if (dbg && threadIdx.x == 64) stamp(5);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(NC < 32 ? 32 : NC));
}

template <int NC, int EPI, int SPLIT = 1, int STAGES = 8> bool launch_ts(const bf * Wt, const bf * hi, const bf * lo, int KT, const TsEpi & e, cudaStream_t stream) {
    static DeviceOnce attr_done;
    int dev = 0;
    MGB_CUDA_TRY(cudaGetDevice(&dev));
    constexpr int kWTile = NC * 128;
    constexpr int smem_max = STAGES * (2 * kXTile + (kWTile < 1024 ? 1024 : kWTile)) + 1024 + 256 + (SPLIT - 1) * 64 * NC * 4;
    static_assert(smem_max <= 227 * 1024, "ts_linear shared memory");
    // (the f16 mode needs only half of the activation bytes per stage, but shrinking the allocation to 80-96 KB lets the CTAs of two
    //  consecutive GEMMs share an SM, and the early-resident successor slows its predecessor: 679 vs 664 us per step -- keep the footprint)
    constexpr int smem = smem_max;
    if (!attr_done.done(dev)) {
        MGB_CUDA_TRY(cudaFuncSetAttribute(ts_linear_kernel<NC, EPI, SPLIT, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
        attr_done.set(dev);
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(e.N / NC, SPLIT); cfg.blockDim = dim3(kTsThreads); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    at[1].id = cudaLaunchAttributeClusterDimension;
    at[1].val.clusterDim.x = 1; at[1].val.clusterDim.y = SPLIT; at[1].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = SPLIT > 1 ? 2 : 1;
    MGB_CUDA_TRY(cudaLaunchKernelEx(&cfg, ts_linear_kernel<NC, EPI, SPLIT, STAGES>, Wt, hi, lo, KT, e));
    MGB_LAUNCH_CHECK();
    return true;
}

}  // namespace

static int ts_shape() {
    static const int shaped = getenv("MGB_TS_SHAPE") ? atoi(getenv("MGB_TS_SHAPE")) : 1;
    return shaped;
}

// columns per CTA of the residual-epilogue GEMM (mirrors the dispatch in launch_linear_ts below)
int ts_resid_nc(int /*K*/) {
    return 8;         // 8-row CTAs, and the split-K kernel's four ranks finish 8 columns each
}

__global__ void row_dots_kernel(const bf * __restrict__ W, const float * __restrict__ v, int N, int K, float * __restrict__ out) {
    const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (n >= N) return;
    float s = 0.0f;
    for (int k = lane; k < K; k += 32) s = fmaf(__bfloat162float(W[(size_t)n * K + k]), v[k], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[n] = s;
}
bool launch_row_dots(const void * W, const float * v, int N, int K, float * out, cudaStream_t stream) {
    row_dots_kernel<<<(N + 7) / 8, 256, 0, stream>>>((const bf *)W, v, N, K, out);
    MGB_LAUNCH_CHECK();
    return true;
}

// The three decoder-step uses with their epilogues; everything else stays on gemm_tc.cu.  hi / lo: packed activations (MT = 64).
bool ts_linear_supported(const LinearArgs & a) {
    if (getenv("MGB_NO_TS") != nullptr) return false;
    if (a.M > 64 || a.W.taps != 1 || a.W.K % 64 != 0 || a.bias || a.W.N % 128 != 0) return false;
    if (a.ln_fold_stats && (!a.x_prepacked || !a.ln_fold_csum || a.ln_fold_slices <= 0 || (a.n_q < 0 && !(a.act == ACT_GELU && a.pack_out)))) return false;
    const bool qkv = a.n_q >= 0 && !a.res && a.act == ACT_NONE && !a.pack_out && a.Y && a.n_q % 16 == 0 && a.dkv % 16 == 0 && a.kdst && a.vdst && (a.ldy % 4) == 0;
    const bool resid = a.n_q < 0 && a.res && a.act == ACT_NONE && a.Y && (a.ldr % 4) == 0 && (a.ldy % 4) == 0 &&
                       (a.pack_out ? (a.next_ln_w && a.stats_out && a.W.N % 64 == 0) : !a.next_ln_w);
    const bool gpack = a.n_q < 0 && !a.res && a.act == ACT_GELU && a.pack_out && !a.Y;
    return qkv || resid || gpack;
}

bool launch_linear_ts(const LinearArgs & a, const void * hi, const void * lo, cudaStream_t stream) {
    TsEpi e;
    e.N = a.W.N; e.M = a.M; e.res = a.res; e.ldr = a.ldr; e.Y = a.Y; e.ldy = a.ldy; e.gelu_f16 = a.gelu_f16;
    e.n_q = a.n_q; e.dkv = a.dkv; e.kdst = (bf *)a.kdst; e.vdst = (bf *)a.vdst; e.tok_slot = a.tok_slot;
    e.pk_hi = nullptr; e.pk_lo = nullptr;
    if (a.pack_out) { e.pk_hi = (bf *)a.pack_out; e.pk_lo = e.pk_hi + (size_t)64 * a.W.N; }
    e.next_w = a.next_ln_w; e.stats_out = a.stats_out;
    e.x_f16 = a.act_f16 ? 1 : 0; e.pack_f16 = a.pack_f16 ? 1 : 0;
    static const bool dbg_on = getenv("MGB_TS_DBG") != nullptr;
    e.dbg = dbg_on ? 1 : 0;
    if (dbg_on && getenv("MGB_TS_DBG_DUMP")) {           // dump at the given call of this function (direct launches, not graph replays)
        static int calls = 0;
        cudaStreamCaptureStatus cst = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(stream, &cst);
        static bool dumped = false;
        if (++calls >= atoi(getenv("MGB_TS_DBG_DUMP")) && cst == cudaStreamCaptureStatusNone && !dumped) {
            dumped = true;
            static unsigned long long h[1024 * 8];
            unsigned n = 0;
            cudaDeviceSynchronize();
            cudaMemcpyFromSymbol(h, g_ts_dbg, sizeof(h));
            cudaMemcpyFromSymbol(&n, g_ts_dbg_n, sizeof(n));
            const unsigned first = n > 160 ? n - 160 : 0;
            unsigned long long prev_end = 0;
            for (unsigned i = first; i < n; i++) {
                const unsigned long long * r = h + (size_t)(i % 1024) * 8;
                fprintf(stderr, "ts %4u id %4llu  gap_from_prev_end %6lld | entry->dep %6lld | dep->first_stage %6lld | ->last_mma %6lld | ->acc %6lld | epilogue %6lld | total %6lld ns\n",
                        i, r[6], prev_end ? (long long)(r[0] - prev_end) : 0LL, (long long)(r[1] - r[0]), (long long)(r[2] - r[1]), (long long)(r[3] - r[2]),
                        (long long)(r[4] - r[3]), (long long)(r[5] - r[4]), (long long)(r[5] - r[0]));
                prev_end = r[5];
            }
        }
    }
    e.ln_stats = a.ln_fold_stats; e.ln_slices = a.ln_fold_slices; e.ln_csum = a.ln_fold_csum; e.eps = a.eps; e.K = a.W.K;
    const int KT = a.W.K / 64;
    if (a.act_f16 && !a.W.tiles16) { set_error("linear: f16 activation images need the f16 weight images (MGB_ACT_F16 at model load)"); return false; }
    const bf * W = (const bf *)(a.act_f16 ? a.W.tiles16 : a.W.tiles);
    // One-CTA slices for the K = 768 GEMMs (QKV 144 CTAs, O 96, FF1 96).  The wide-K GEMM (FF2, K = 3072) is bound by how fast ONE
    // SM ingests the 786 KB activation operand (the same 21 us with 24, 48 or 96 one-CTA slices), so its 32-row slices are shared
    // by a 4-CTA cluster that divides the k tiles (10 us).  Measured at 64 utterances, us per step: no split 1331, FF2 split over
    // 2 / 4 CTAs 1241 / 1201; cluster slices for the K = 768 GEMMs as well (64 rows x 4, 32 x 4, 64 x 3): 1343 -- the cluster
    // launch and the DSMEM reduction cost more than the ingest they save (MGB_TS_SHAPE=2 keeps that variant for A/B, 0 = no split).
    const int shaped = ts_shape();
    const bf * h = (const bf *)hi, * l = (const bf *)lo;
    if (a.n_q >= 0) {
        // (16 rows per CTA = 144 CTAs; measured: 32 rows / 72 CTAs the same within 1 %, 8 rows / 288 CTAs = two waves, -11 %)
        return launch_ts<16, TS_QKV>(W, h, l, KT, e, stream);
    }
    if (a.res) {
        if (shaped && KT >= 32 && KT % 4 == 0) return launch_ts<32, TS_RES, 4, 8>(W, h, l, KT, e, stream);        // FF2: 24 slices x 4 CTAs, 12 k tiles each
        return launch_ts<8, TS_RES>(W, h, l, KT, e, stream);
    }
    return launch_ts<32, TS_GELU_PACK>(W, h, l, KT, e, stream);
}

}  // namespace mgb
