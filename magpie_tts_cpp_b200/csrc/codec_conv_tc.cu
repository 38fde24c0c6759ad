// nano-codec residual-block convs as a persistent implicit-GEMM kernel on the 5th-gen tensor cores (tcgen05 + TMEM).
//
// Reference: magpie_codec_build_causal_conv1d (src/nano-codec.cpp:429-466) = ggml_conv_1d, i.e. im2col rounded to f16,
// kernel rounded to f16, f32 accumulation; magpie_codec_build_half_snake (:376-426); residual blocks (:568-641).
//
//   y[co, t] = b[co] + sum_{k, ci} w[co, ci, k] * xa[ci, t - (K-1-k)*dil]
//
// is computed per 128-step time tile as   D[t][co] += A_k[t][ci] * W_k[co][ci]   over the K taps, with
//   * A (MMA "A", M = 128 time steps): the activated f16 input kept in HBM as a TIME-MAJOR image ([t][64 ch] rows of
//     128 bytes, SWIZZLE_128B by row), so ONE cp.async.bulk of rows [t0 - halo, t0 + 128) stages the operand of all K
//     taps: tap k is the same shared-memory image read through a matrix descriptor whose start address is shifted by
//     (halo - (K-1-k)*dil) rows.  No im2col is ever materialised (the reference materialises it per conv).
//   * W (MMA "B", N = output channels padded to 16): per (64-channel chunk, tap) a [npad x 64] K-major f16 tile image,
//     packed once at load, streamed through a shared-memory ring with cp.async.bulk.
//   * D in TMEM, fp32, DOUBLE BUFFERED (2 x npad columns): the epilogue of tile i runs while the MMAs of tile i+1 issue.
//   * epilogue (8 warps, tcgen05.ld): bias, f32 residual add, optional f32 outputs (residual stream / 3-branch mean),
//     HalfSnake of the NEXT conv + f16 rounding written straight into the next conv's time-major image.
//
// Warp roles (320 threads): warp 0 = bulk-copy producer, warp 1 = TMEM allocator + MMA issuer, warps 2..9 = epilogue.
// Persistent: grid = min(#tiles, #SMs); tiles are taken round-robin; all pipelines run across tile boundaries.
#include <algorithm>

#include "codec_tc.h"
#include "common.cuh"
#include "gemm_tc.cuh"

namespace mgb {
namespace ctc {

namespace {

constexpr int kEpiWarps = 16;
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kMaxC = 448;
constexpr int kMaxNA = 4, kMaxNW = 8;
constexpr int kDynSmemMax = 218 * 1024;                 // + 7 KB static epilogue table
constexpr int kSmemBudget = 214 * 1024;

struct KParams {
    ConvArgs a;
    int C, nsplit, nper, npad, nchunk, cs;
    int rb;                 // bytes per image / weight-tile row: 128 (64 channels, SWIZZLE_128B) or 64 (32 channels, SWIZZLE_64B)
    uint32_t desc_hi;       // constant high word of the shared-memory matrix descriptors for this row size
    int hp;                 // staged history rows, (K-1)*dil rounded up to 8
    int tps, ngroups;       // taps per weight stage, stages per chunk
    int NA, NW;             // ring depths
    int wstage_bytes;       // bytes of a full weight stage
    int R;                  // 128-step accumulator tiles per work item (R * npad * 2 TMEM columns)
    int abuf_bytes;         // bytes of one staged activation window: (R * 128 + kHP) rows
    int tiles_per_b, n_tiles;   // work items per utterance / in total
    int tmem_cols;
    long long rows;         // rows per (utterance, chunk) of the activation images
};

__device__ __forceinline__ float f16r(float x) { return __half2float(__float2half_rn(x)); }

// XOR applied to the 16-byte group index of image row r: SWIZZLE_128B (8 groups per 128-byte row) or SWIZZLE_64B (4 groups per
// 64-byte row, address bits [4,5] ^= bits [7,8])
__host__ __device__ __forceinline__ int img_swz(int r, int rb) { return rb == 128 ? (r & 7) : ((r >> 1) & 3); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
// low word of a K-major SWIZZLE_128B shared-memory matrix descriptor (start address >> 4, LBO = 1); the high word is
// constant: SBO = 1024 B >> 4 (bits 32-45), version 1 (bit 46), layout SWIZZLE_128B = 2 (bits 61-63)
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr) { return ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16); }
// (SWIZZLE_64B rows of 64 bytes: SBO = 512 B >> 4 = 32, layout 4)
__host__ __device__ constexpr uint32_t desc_hi_for(int rb) { return rb == 128 ? (64u | (1u << 14) | (2u << 29)) : (32u | (1u << 14) | (4u << 29)); }
__device__ __forceinline__ void umma_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate, uint32_t desc_hi) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "mov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}"
                 ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(desc_hi) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t * bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// 256-bit global accesses (one full 32-byte sector per thread)
__device__ __forceinline__ void ldg256(const float * p, float * v) {
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p) : "memory");
}
__device__ __forceinline__ void stg256f(float * p, const float * v) {
    asm volatile("st.global.v8.f32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};"
                 ::"f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "l"(p) : "memory");
}
__device__ __forceinline__ void stg256(void * p, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e, uint32_t f, uint32_t g, uint32_t h) {
    asm volatile("st.global.v8.b32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};"
                 ::"r"(a), "r"(b), "r"(c), "r"(d), "r"(e), "r"(f), "r"(g), "r"(h), "l"(p) : "memory");
}
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr) : "memory");
}
// x + sin^2(alpha x)/alpha with an explicit two-constant reduction to [-pi, pi] and the SFU sine (abs error ~5e-7, far
// below the f16 rounding applied to the result)
__device__ __forceinline__ float snake_fast(float x, float a, float inv_a) {
    const float z = x * a;
    const float n = rintf(z * 0.15915494309189535f);
    float r = fmaf(n, -6.2831854820251465f, z);
    r = fmaf(n, 1.7484556000744883e-7f, r);
    const float sn = __sinf(r);
    return fmaf(sn * sn, inv_a, x);
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t *>(&h);
}

// MODE 0: activated image only (first conv of a block); 1: + residual, f32 output and activated image (second conv);
// 2: + residual, f32 output only (last block of a branch: the 3-branch mean is taken by the consumer, up_tm / post_tm)
template <int MODE>
__global__ void __launch_bounds__(kThreads, 1) conv_tc_kernel(const KParams p) {
    extern __shared__ unsigned char smem_raw[];
    // per output channel: bias, alpha2 (0 = LeakyReLU) and 1/alpha2; per 8-channel group: 0 = all LeakyReLU, 1 = all snake, 2 = mixed
    __shared__ __align__(16) float s_bias[kMaxC], s_alpha[kMaxC], s_inv[kMaxC];
    __shared__ int s_kind[kMaxC / 8];
    unsigned char * base = reinterpret_cast<unsigned char *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char * abuf = base;
    unsigned char * wbuf = base + p.NA * p.abuf_bytes;
    uint64_t * bars = reinterpret_cast<uint64_t *>(wbuf + (size_t)p.NW * p.wstage_bytes);
    uint64_t * a_full = bars, * a_empty = bars + kMaxNA, * w_full = bars + 2 * kMaxNA, * w_empty = w_full + kMaxNW;
    uint64_t * acc_full = w_empty + kMaxNW, * acc_empty = acc_full + 2;
    uint32_t * tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    for (int c = threadIdx.x; c < p.cs; c += kThreads) {
        const float al = (MODE != 2 && c < p.a.n_alpha2 && c < p.C) ? p.a.alpha2[c] : 0.0f;
        s_bias[c] = c < p.C ? p.a.bias[c] : 0.0f;
        s_alpha[c] = al;
        s_inv[c] = al != 0.0f ? 1.0f / al : 0.0f;
    }
    for (int g8 = threadIdx.x; g8 < p.cs / 8; g8 += kThreads) {
        int n_snake = 0;
        for (int e = 0; e < 8; e++) {
            const int c = g8 * 8 + e;
            n_snake += (MODE != 2 && c < p.a.n_alpha2 && c < p.C && p.a.alpha2[c] != 0.0f) ? 1 : 0;
        }
        s_kind[g8] = n_snake == 0 ? 0 : (n_snake == 8 ? 1 : 2);
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < p.NA; i++) { tc::mbar_init(&a_full[i], 1); tc::mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < p.NW; i++) { tc::mbar_init(&w_full[i], 1); tc::mbar_init(&w_empty[i], 1); }
        for (int i = 0; i < 2; i++) { tc::mbar_init(&acc_full[i], 1); tc::mbar_init(&acc_empty[i], kThreads - 64); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(tmem_slot)), "r"(p.tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const int K = p.a.K, T = p.a.T;
    const int tile_bytes = p.npad * p.rb;
    const int cpr = p.rb >> 1;           // channels per row

    // Producer and MMA roles run their loops warp-uniformly (all lanes poll the barriers; one elected lane issues), so
    // that addresses and descriptors live in uniform registers and the per-MMA issue cost stays a few instructions.
    if (warp == 0) {
        const uint32_t abytes = (uint32_t)(p.R * kTile + p.hp) * p.rb;
        const bool leader = elect_one();
        uint32_t ia = 0, iw = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            const int tb = tile % p.tiles_per_b, hb = tile / p.tiles_per_b;
            const int h = hb % p.nsplit, b = hb / p.nsplit;
            const long long row0 = (long long)kHP + (long long)tb * (p.R * kTile) - p.hp;
            for (int c = 0; c < p.nchunk; c++) {
                const int sa = ia % p.NA;
                tc::mbar_wait(&a_empty[sa], ((ia / p.NA) & 1) ^ 1);
                if (leader) {
                    tc::mbar_expect_tx(&a_full[sa], abytes);
                    tc::bulk_g2s(abuf + (size_t)sa * p.abuf_bytes,
                                 reinterpret_cast<const unsigned char *>(p.a.xa) + (((long long)b * p.nchunk + c) * p.rows + row0) * p.rb,
                                 abytes, &a_full[sa]);
                }
                ia++;
                const unsigned char * wsrc = reinterpret_cast<const unsigned char *>(p.a.w) + ((size_t)(h * p.nchunk + c) * K) * tile_bytes;
                for (int g = 0; g < p.ngroups; g++) {
                    const int sw = iw % p.NW;
                    const int nt = min(p.tps, K - g * p.tps);
                    tc::mbar_wait(&w_empty[sw], ((iw / p.NW) & 1) ^ 1);
                    if (leader) {
                        tc::mbar_expect_tx(&w_full[sw], (uint32_t)(nt * tile_bytes));
                        tc::bulk_g2s(wbuf + (size_t)sw * p.wstage_bytes, wsrc + (size_t)g * p.tps * tile_bytes, (uint32_t)(nt * tile_bytes), &w_full[sw]);
                    }
                    iw++;
                }
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc = tc::umma_idesc_f16(kTile, p.npad);
        const bool leader = elect_one();
        const uint32_t tile_d = (uint32_t)tile_bytes >> 4;             // descriptor address units are 16 bytes
        const uint32_t dil_d = (uint32_t)p.a.dil * (p.rb >> 4);         // one row = rb bytes = rb / 16 units
        const uint32_t tile_rows_d = (uint32_t)kTile * (p.rb >> 4);
        uint32_t ia = 0, iw = 0, it = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, it++) {
            const int buf = it & 1;
            tc::mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t dcol = tmem_base + (uint32_t)(buf * p.R * p.npad);
            // accumulator tiles of this item that hold valid time steps (the last item of an utterance may be partial)
            const int tb = tile % p.tiles_per_b;
            const int rt = min(p.R, (T - tb * (p.R * kTile) + kTile - 1) / kTile);
            uint32_t accum = 0;
            for (int c = 0; c < p.nchunk; c++) {
                const int sa = ia % p.NA;
                tc::mbar_wait(&a_full[sa], (ia / p.NA) & 1);
                // tap 0 reads rows hp - (K-1)*dil ..., every further tap dil rows later
                const uint32_t a_lo0 = desc_lo(tc::smem_u32(abuf + (size_t)sa * p.abuf_bytes) + (uint32_t)(p.hp - (K - 1) * p.a.dil) * p.rb);
                const int nk16 = min(cpr >> 4, (p.C - c * cpr + 15) >> 4);
                for (int g = 0; g < p.ngroups; g++) {
                    const int sw = iw % p.NW;
                    const int nt = min(p.tps, K - g * p.tps);
                    tc::mbar_wait(&w_full[sw], (iw / p.NW) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    uint32_t a_lo = a_lo0 + (uint32_t)(g * p.tps) * dil_d;
                    uint32_t w_lo = desc_lo(tc::smem_u32(wbuf + (size_t)sw * p.wstage_bytes));
                    if (leader) {
                        for (int q = 0; q < nt; q++) {
                            for (int r = 0; r < rt; r++) {
                                const uint32_t ar = a_lo + (uint32_t)r * tile_rows_d, dr = dcol + (uint32_t)(r * p.npad);
                                umma_lo(dr, ar, w_lo, idesc, accum, p.desc_hi);
                                if (nk16 > 1) umma_lo(dr, ar + 2, w_lo + 2, idesc, 1u, p.desc_hi);
                                if (nk16 > 2) umma_lo(dr, ar + 4, w_lo + 4, idesc, 1u, p.desc_hi);
                                if (nk16 > 3) umma_lo(dr, ar + 6, w_lo + 6, idesc, 1u, p.desc_hi);
                            }
                            accum = 1;
                            a_lo += dil_d; w_lo += tile_d;
                        }
                        tc::umma_commit(&w_empty[sw]);
                    }
                    accum = 1;
                    iw++;
                }
                if (leader) tc::umma_commit(&a_empty[sa]);
                ia++;
            }
            if (leader) tc::umma_commit(&acc_full[buf]);
        }
    } else {
        // Epilogue: thread = one time step (TMEM lane), warp = 32 steps x a share of the 16-channel column units.  The
        // residual values of the NEXT unit (possibly of the next tile) are fetched while the current one is processed.
        const int q = warp & 3;                        // TMEM lane quarter this warp may access
        const int part = (warp - 2) >> 2;              // which share of the columns (kEpiWarps / 4 shares)
        const int n16 = p.npad >> 4;
        const int n_units = p.R * n16;                 // (accumulator tile r, 16-column unit u) pairs: w = r * n16 + u
        const int w_lo = n_units * part / (kEpiWarps / 4), w_hi = n_units * (part + 1) / (kEpiWarps / 4);
        const long long hrow_bytes = p.rows * p.rb;
        const size_t tile_stride = (size_t)kTile * p.cs;          // f32 elements between accumulator tiles of an item
        struct TileInfo { int t0, r_img0, cbase, b, rt; size_t row_off; };     // r = 0 row of this thread
        auto decode = [&](int tile) {
            TileInfo ti;
            const int tb = tile % p.tiles_per_b, hb = tile / p.tiles_per_b;
            const int h = hb % p.nsplit;
            ti.b = hb / p.nsplit;
            ti.t0 = tb * (p.R * kTile) + q * 32 + lane;
            ti.rt = tile < p.n_tiles ? min(p.R, (T - tb * (p.R * kTile) + kTile - 1) / kTile) : 0;
            ti.r_img0 = kHP + ti.t0;
            ti.cbase = h * p.nper;
            ti.row_off = ((size_t)ti.b * T + ti.t0) * p.cs + ti.cbase;
            return ti;
        };
        uint32_t it = 0;
        TileInfo cur = decode(blockIdx.x);
        float rn[16];
        int pf_tile = -1, pf_w = -1;                  // which (item, unit) the prefetched residual in rn belongs to
#pragma unroll
        for (int e = 0; e < 16; e++) rn[e] = 0.0f;
        auto res_fetch = [&](const TileInfo & ti, int tile, int w) {
            const int r = w / n16, u = w - r * n16;
            pf_tile = tile; pf_w = w;
            if (ti.t0 + r * kTile < T) {
                const float * src = p.a.res + ti.row_off + r * tile_stride + u * 16;
                ldg256(src, rn); ldg256(src + 8, rn + 8);
            }
        };
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, it++) {
            const int buf = it & 1;
            const TileInfo nxt = decode(tile + gridDim.x);
            const int w_end = min(w_hi, cur.rt * n16);            // units of valid accumulator tiles only
            if (MODE != 0 && w_lo < w_end && !(pf_tile == tile && pf_w == w_lo)) res_fetch(cur, tile, w_lo);
            tc::mbar_wait(&acc_full[buf], (it >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * p.R * p.npad);
            for (int w = w_lo; w < w_end; w++) {
                const int r = w / n16, u = w - r * n16;
                const bool tv = cur.t0 + r * kTile < T;
                const int r_img = cur.r_img0 + r * kTile;
                const size_t row_off = cur.row_off + r * tile_stride;
                float rr[16];
                if (MODE != 0) {
#pragma unroll
                    for (int e = 0; e < 16; e++) rr[e] = rn[e];
                    // prefetch the residual of the next unit (of this item, else the first one of the next item)
                    if (w + 1 < w_end) res_fetch(cur, tile, w + 1);
                    else if (w_lo < min(w_hi, nxt.rt * n16)) res_fetch(nxt, tile + gridDim.x, w_lo);
                }
                uint32_t v[16];
                tmem_ld16(trow + r * p.npad + u * 16, v);
                const int co0 = cur.cbase + u * 16;
                float y[16];
                {
                    const float4 * bp = reinterpret_cast<const float4 *>(s_bias + co0);
#pragma unroll
                    for (int e4 = 0; e4 < 4; e4++) {
                        const float4 bv = bp[e4];
                        y[4 * e4] = __uint_as_float(v[4 * e4]) + bv.x; y[4 * e4 + 1] = __uint_as_float(v[4 * e4 + 1]) + bv.y;
                        y[4 * e4 + 2] = __uint_as_float(v[4 * e4 + 2]) + bv.z; y[4 * e4 + 3] = __uint_as_float(v[4 * e4 + 3]) + bv.w;
                    }
                    if (MODE != 0) {
#pragma unroll
                        for (int e = 0; e < 16; e++) y[e] += rr[e];
                    }
                }
                if (MODE != 2) {
                    uint32_t pk[8];
#pragma unroll
                    for (int hh = 0; hh < 2; hh++) {
                        float act[8];
                        const int kind = s_kind[(co0 >> 3) + hh];         // warp-uniform
                        if (kind == 0) {
#pragma unroll
                            for (int e = 0; e < 8; e++) act[e] = fmaxf(y[hh * 8 + e], 0.01f * y[hh * 8 + e]);
                        } else {
                            const float4 * ap = reinterpret_cast<const float4 *>(s_alpha + co0 + hh * 8);
                            const float4 * ip = reinterpret_cast<const float4 *>(s_inv + co0 + hh * 8);
                            const float4 a0 = ap[0], a1 = ap[1], i0 = ip[0], i1 = ip[1];
                            const float al[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                            const float iv[8] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w};
#pragma unroll
                            for (int e = 0; e < 8; e++) act[e] = snake_fast(y[hh * 8 + e], al[e], iv[e]);     // branch-free: independent chains
                            if (kind == 2) {
#pragma unroll
                                for (int e = 0; e < 8; e++) act[e] = al[e] != 0.0f ? act[e] : fmaxf(y[hh * 8 + e], 0.01f * y[hh * 8 + e]);
                            }
                        }
#pragma unroll
                        for (int e2 = 0; e2 < 4; e2++) pk[hh * 4 + e2] = pack_h2(act[2 * e2], act[2 * e2 + 1]);
                    }
                    if (tv) {
                        // the unit's two 16-byte groups share one 32-byte sector of the swizzled row; odd rows swap them
                        const int sw = img_swz(r_img, p.rb);
                        unsigned char * dst = reinterpret_cast<unsigned char *>(p.a.ya) + ((long long)cur.b * p.nchunk + co0 / cpr) * hrow_bytes +
                                              (long long)r_img * p.rb + (((((co0 % cpr) >> 3) ^ sw) & 6) << 4);
                        const bool swap = sw & 1;
                        stg256(dst, swap ? pk[4] : pk[0], swap ? pk[5] : pk[1], swap ? pk[6] : pk[2], swap ? pk[7] : pk[3],
                                    swap ? pk[0] : pk[4], swap ? pk[1] : pk[5], swap ? pk[2] : pk[6], swap ? pk[3] : pk[7]);
                    }
                }
                if (MODE != 0 && tv) {
                    float * yp = p.a.y + row_off + u * 16;
                    stg256f(yp, y); stg256f(yp + 8, y + 8);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(&acc_empty[buf]);
            cur = nxt;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols));
}

// f32 (C, C, K) -> f16 tile images [nsplit][nchunk][K][npad x 64]; one thread per 16-byte group of 8 input channels
__global__ void pack_w_kernel(const float * w, int C, int K, int nsplit, int nper, int npad, int nchunk, int rb, __half * out) {
    const int gpr = rb >> 4, cpr = rb >> 1;             // 16-byte groups / channels per row
    const size_t total = (size_t)nsplit * nchunk * K * npad * gpr;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int kc = (int)(i % gpr);
        size_t r = i / gpr;
        const int n = (int)(r % npad); r /= npad;
        const int k = (int)(r % K); r /= K;
        const int c = (int)(r % nchunk); const int h = (int)(r / nchunk);
        const int co = h * nper + n;
        __half hv[8];
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const int ci = c * cpr + kc * 8 + e;
            hv[e] = __float2half_rn((n < nper && ci < C) ? w[((size_t)co * C + ci) * K + k] : 0.0f);
        }
        unsigned char * dst = reinterpret_cast<unsigned char *>(out) + (((size_t)(h * nchunk + c) * K + k) * npad) * rb + (size_t)n * rb +
                              ((kc ^ img_swz(n, rb)) << 4);
        *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(hv);
    }
}

// HalfSnake -> grouped ConvTranspose1d (groups = Cout, 2 inputs per group, K = 2*stride, first T*stride samples kept;
// nano-codec.cpp:481-565) on time-major rows, producing the stage input `up` AND the activated images of the three
// residual branches (their first HalfSnake) in one pass.  grid (time blocks of 128, 8-channel groups, B).
struct UpKParams { UpArgs a; int cs_in, cs_out, nchunk, rb; long long rows; };
__device__ __forceinline__ float half_snake_fast(float x, int c, const float * alpha, int n_alpha) {
    if (c < n_alpha) {
        const float a = __ldg(alpha + c);
        return snake_fast(x, a, 1.0f / a);
    }
    return fmaxf(x, 0.01f * x);
}
// thread = one INPUT time step of one 8-output-channel group: the 2 x 16 activated inputs (this step and the previous
// one, taken from the neighbouring lane) are computed once and reused for all s output steps; the per-channel constants
// of the CTA's channel group live in shared memory.
__global__ void __launch_bounds__(128) up_tm_kernel(const UpKParams p) {
    __shared__ float s_ia[16], s_iinv[16], s_w[16 * 16], s_b[8], s_ba[3][8], s_binv[3][8];
    const int g0 = blockIdx.y * 8, b = blockIdx.z;
    const int Cout = p.a.Cin / 2, s = p.a.s, K = 2 * s;
    const int tid = threadIdx.x;
    if (tid < 16) {
        const int ci = 2 * g0 + tid;
        const float al = (ci < p.a.n_alpha && ci < p.a.Cin) ? p.a.alpha[ci] : 0.0f;
        s_ia[tid] = al; s_iinv[tid] = al != 0.0f ? 1.0f / al : 0.0f;
    }
    for (int i = tid; i < 16 * K; i += 128) {
        const int ci = 2 * g0 + i / K;
        s_w[i] = ci < p.a.Cin ? p.a.w[(size_t)ci * K + i % K] : 0.0f;
    }
    if (tid < 8) s_b[tid] = g0 + tid < Cout ? p.a.bias[g0 + tid] : 0.0f;
    if (tid >= 32 && tid < 56) {
        const int j = (tid - 32) / 8, e = (tid - 32) % 8, c = g0 + e;
        const float al = (c < p.a.n_br_alpha && c < Cout) ? p.a.br_alpha[j][c] : 0.0f;
        s_ba[j][e] = al; s_binv[j][e] = al != 0.0f ? 1.0f / al : 0.0f;
    }
    __syncthreads();
    const int ti = blockIdx.x * 128 + tid;
    const bool tv = ti < p.a.T;
    // activated inputs of this step; the previous step's come from lane - 1 (lane 0 computes them itself)
    float a_cur[16], a_prev[16];
    auto load_act = [&](int t, float * out) {
        const bool ok = t >= 0 && t < p.a.T && g0 < Cout;
        const size_t xo = ((size_t)b * p.a.T + (ok ? t : 0)) * p.cs_in + 2 * g0;
#pragma unroll
        for (int q4 = 0; q4 < 4; q4++) {
            float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok && 2 * g0 + 4 * q4 < p.cs_in) {
                f = *reinterpret_cast<const float4 *>(p.a.x[0] + xo + 4 * q4);
                if (p.a.n_x == 3) {        // input = mean of the previous stage's three residual branches (nano-codec.cpp:601-641)
                    const float4 g = *reinterpret_cast<const float4 *>(p.a.x[1] + xo + 4 * q4), h = *reinterpret_cast<const float4 *>(p.a.x[2] + xo + 4 * q4);
                    f.x = ((f.x + g.x) + h.x) * (1.0f / 3.0f); f.y = ((f.y + g.y) + h.y) * (1.0f / 3.0f);
                    f.z = ((f.z + g.z) + h.z) * (1.0f / 3.0f); f.w = ((f.w + g.w) + h.w) * (1.0f / 3.0f);
                }
            }
            out[4 * q4] = f.x; out[4 * q4 + 1] = f.y; out[4 * q4 + 2] = f.z; out[4 * q4 + 3] = f.w;
        }
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const float al = s_ia[i];                      // warp-uniform
            out[i] = al != 0.0f ? snake_fast(out[i], al, s_iinv[i]) : fmaxf(out[i], 0.01f * out[i]);
        }
    };
    load_act(ti, a_cur);
#pragma unroll
    for (int i = 0; i < 16; i++) a_prev[i] = __shfl_up_sync(0xffffffffu, a_cur[i], 1);
    if ((tid & 31) == 0) load_act(ti - 1, a_prev);        // zero history for ti == 0 (ok == false -> act(0) = 0)
    if (!tv) return;
    const int To = p.a.T * s;
    for (int r = 0; r < s; r++) {
        const int to = ti * s + r;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const float * w0 = s_w + (2 * e) * K + r, * w1 = w0 + K;
            v[e] = s_b[e] + w0[0] * a_cur[2 * e] + w1[0] * a_cur[2 * e + 1] + w0[s] * a_prev[2 * e] + w1[s] * a_prev[2 * e + 1];
        }
        stg256f(p.a.up + ((size_t)b * To + to) * p.cs_out + g0, v);
        const int r_img = kHP + to;
        const int cpr = p.rb >> 1;
        const long long off = (((long long)b * p.nchunk + g0 / cpr) * p.rows + r_img) * p.rb + ((((g0 % cpr) >> 3) ^ img_swz(r_img, p.rb)) << 4);
#pragma unroll
        for (int j = 0; j < 3; j++) {
            float act[8];
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const float al = s_ba[j][e];               // warp-uniform
                act[e] = al != 0.0f ? snake_fast(v[e], al, s_binv[j][e]) : fmaxf(v[e], 0.01f * v[e]);
            }
            uint4 pk;
            pk.x = pack_h2(act[0], act[1]); pk.y = pack_h2(act[2], act[3]);
            pk.z = pack_h2(act[4], act[5]); pk.w = pack_h2(act[6], act[7]);
            *reinterpret_cast<uint4 *>(reinterpret_cast<unsigned char *>(p.a.img[j]) + off) = pk;
        }
    }
}

// Same operation for stages with at most kUpNG 8-channel output groups (every stage but the first): lanes run ACROSS the
// channel groups of a row, so every warp access is a run of whole rows (f32 rows in, f32 rows + swizzled f16 rows out, no
// partially written sectors), and a thread walks kUpNT consecutive input steps carrying the previous step's activated
// inputs in registers (no neighbour exchange; the next step's rows are loaded before the current one is computed).
constexpr int kUpNG = 28, kUpNT = 8;
template <int S>
__global__ void __launch_bounds__(128) up_rows_kernel(const UpKParams p) {
    constexpr int K = 2 * S, WS = 16 * K + 1;
    __shared__ float s_w[kUpNG * WS], s_ia[kUpNG * 17], s_iinv[kUpNG * 17], s_b[kUpNG * 9], s_ba[3][kUpNG * 9], s_binv[3][kUpNG * 9];
    const int ng = p.cs_out >> 3, b = blockIdx.y, tid = threadIdx.x;
    const int Cout = p.a.Cin / 2;
    for (int i = tid; i < ng * 16; i += 128) {
        const float al = (i < p.a.n_alpha && i < p.a.Cin) ? p.a.alpha[i] : 0.0f;
        s_ia[(i >> 4) * 17 + (i & 15)] = al; s_iinv[(i >> 4) * 17 + (i & 15)] = al != 0.0f ? 1.0f / al : 0.0f;
    }
    for (int i = tid; i < ng * 16 * K; i += 128) {
        const int ci = i / K;
        s_w[(ci >> 4) * WS + (ci & 15) * K + i % K] = ci < p.a.Cin ? p.a.w[i] : 0.0f;
    }
    for (int i = tid; i < ng * 8; i += 128) {
        const int o = (i >> 3) * 9 + (i & 7);
        s_b[o] = i < Cout ? p.a.bias[i] : 0.0f;
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const float al = (i < p.a.n_br_alpha && i < Cout) ? p.a.br_alpha[j][i] : 0.0f;
            s_ba[j][o] = al; s_binv[j][o] = al != 0.0f ? 1.0f / al : 0.0f;
        }
    }
    __syncthreads();
    const int cpb = 128 / ng, chunk = tid / ng, gi = tid - chunk * ng;
    const int t0 = (blockIdx.x * cpb + chunk) * kUpNT;
    if (chunk >= cpb || t0 >= p.a.T) return;
    const int g0 = gi * 8;
    const bool in_ok[4] = {2 * g0 < p.cs_in, 2 * g0 + 4 < p.cs_in, 2 * g0 + 8 < p.cs_in, 2 * g0 + 12 < p.cs_in};
    auto load_raw = [&](int t, float * out) {          // mean of the branch outputs (nano-codec.cpp:601-641) or the single input
        const size_t xo = ((size_t)b * p.a.T + t) * p.cs_in + 2 * g0;
#pragma unroll
        for (int q4 = 0; q4 < 4; q4++) {
            float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
            if (in_ok[q4]) {
                f = *reinterpret_cast<const float4 *>(p.a.x[0] + xo + 4 * q4);
                if (p.a.n_x == 3) {
                    const float4 g = *reinterpret_cast<const float4 *>(p.a.x[1] + xo + 4 * q4), h = *reinterpret_cast<const float4 *>(p.a.x[2] + xo + 4 * q4);
                    f.x = ((f.x + g.x) + h.x) * (1.0f / 3.0f); f.y = ((f.y + g.y) + h.y) * (1.0f / 3.0f);
                    f.z = ((f.z + g.z) + h.z) * (1.0f / 3.0f); f.w = ((f.w + g.w) + h.w) * (1.0f / 3.0f);
                }
            }
            out[4 * q4] = f.x; out[4 * q4 + 1] = f.y; out[4 * q4 + 2] = f.z; out[4 * q4 + 3] = f.w;
        }
    };
    auto activate = [&](float * v) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const float al = s_ia[gi * 17 + i];
            v[i] = al != 0.0f ? snake_fast(v[i], al, s_iinv[gi * 17 + i]) : fmaxf(v[i], 0.01f * v[i]);
        }
    };
    float a_prev[16], a_cur[16], nxt[16];
#pragma unroll
    for (int i = 0; i < 16; i++) a_prev[i] = 0.0f;        // zero history in front of step 0: act(0) = 0
    if (t0 > 0) { load_raw(t0 - 1, a_prev); activate(a_prev); }
    load_raw(t0, nxt);
    const int To = p.a.T * S;
    const int cpr = p.rb >> 1;
    const long long img_base = ((long long)b * p.nchunk + g0 / cpr) * p.rows;
    const int kc = (g0 % cpr) >> 3;
    const float * wg = s_w + gi * WS;
    for (int i = 0; i < kUpNT; i++) {
        const int ti = t0 + i;
        if (ti >= p.a.T) break;
#pragma unroll
        for (int e = 0; e < 16; e++) a_cur[e] = nxt[e];
        if (i + 1 < kUpNT && ti + 1 < p.a.T) load_raw(ti + 1, nxt);
        activate(a_cur);
#pragma unroll
        for (int r = 0; r < S; r++) {
            const int to = ti * S + r;
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const float * w0 = wg + (2 * e) * K + r, * w1 = w0 + K;
                v[e] = s_b[gi * 9 + e] + w0[0] * a_cur[2 * e] + w1[0] * a_cur[2 * e + 1] + w0[S] * a_prev[2 * e] + w1[S] * a_prev[2 * e + 1];
            }
            stg256f(p.a.up + ((size_t)b * To + to) * p.cs_out + g0, v);
            const int r_img = kHP + to;
            const long long off = (img_base + r_img) * p.rb + ((kc ^ img_swz(r_img, p.rb)) << 4);
#pragma unroll
            for (int j = 0; j < 3; j++) {
                float act[8];
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    const float al = s_ba[j][gi * 9 + e];
                    act[e] = al != 0.0f ? snake_fast(v[e], al, s_binv[j][gi * 9 + e]) : fmaxf(v[e], 0.01f * v[e]);
                }
                uint4 pk;
                pk.x = pack_h2(act[0], act[1]); pk.y = pack_h2(act[2], act[3]);
                pk.z = pack_h2(act[4], act[5]); pk.w = pack_h2(act[6], act[7]);
                *reinterpret_cast<uint4 *>(reinterpret_cast<unsigned char *>(p.a.img[j]) + off) = pk;
            }
        }
#pragma unroll
        for (int e = 0; e < 16; e++) a_prev[e] = a_cur[e];
    }
}

// HalfSnake -> causal conv (C -> 1, K taps, operands rounded to f16 as ggml_conv_1d does) -> tanh  (nano-codec.cpp:703-712)
struct PostKParams { PostArgs a; int cs; };
// block = 256 consecutive time steps of one utterance (+ K-1 history rows): every thread activates ITS row once
// (3-branch mean -> HalfSnake -> f16 rounding) into shared memory, then takes the K-tap dot product from there.
constexpr int kPostCS = 32;               // channels per row (27 used)
__global__ void __launch_bounds__(256) post_tm_kernel(const PostKParams p) {
    __shared__ float sa[(256 + 8) * (kPostCS + 1)];
    __shared__ float sw[kPostCS * 8];
    const int b = blockIdx.y, t0 = blockIdx.x * 256, tid = threadIdx.x;
    const int H = p.a.K - 1;              // history rows
    for (int i = tid; i < p.a.C * p.a.K; i += 256) sw[i] = f16r(p.a.w[i]);
    // staging: consecutive lanes take consecutive 16-byte pieces of the rows (whole-row runs per warp access)
    const int q4n = p.cs >> 2;
    for (int i = tid; i < (256 + H) * q4n; i += 256) {
        const int r = i / q4n, c0 = (i - r * q4n) * 4;
        const int t = t0 - H + r;
        float * dst = sa + r * (kPostCS + 1) + c0;
        float xv[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        const bool ok = t >= 0 && t < p.a.T;
        if (ok) {
            const size_t xo = ((size_t)b * p.a.T + t) * p.cs + c0;
            const float4 f = *reinterpret_cast<const float4 *>(p.a.x[0] + xo), g = *reinterpret_cast<const float4 *>(p.a.x[1] + xo),
                         h = *reinterpret_cast<const float4 *>(p.a.x[2] + xo);
            xv[0] = ((f.x + g.x) + h.x) * (1.0f / 3.0f); xv[1] = ((f.y + g.y) + h.y) * (1.0f / 3.0f);
            xv[2] = ((f.z + g.z) + h.z) * (1.0f / 3.0f); xv[3] = ((f.w + g.w) + h.w) * (1.0f / 3.0f);
        }
#pragma unroll
        for (int e = 0; e < 4; e++)
            if (c0 + e < p.a.C) dst[e] = ok ? f16r(half_snake_fast(xv[e], c0 + e, p.a.alpha, p.a.n_alpha)) : 0.0f;
    }
    __syncthreads();
    const int t = t0 + tid;
    if (t >= p.a.T) return;
    float acc = 0.0f;
    for (int k = 0; k < p.a.K; k++) {
        const float * row = sa + (tid + k) * (kPostCS + 1);       // input step t - (K-1-k)
        for (int c = 0; c < p.a.C; c++) acc = fmaf(sw[c * p.a.K + k], row[c], acc);
    }
    p.a.pcm[(size_t)b * p.a.T + t] = tanhf(acc + __ldg(p.a.bias));
}

}  // namespace

Geom geom_for(int C) {
    Geom g;
    g.C = C;
    if (C <= 0) return g;
    // output channels in `nsplit` equal groups of at most 224; a split group must be a whole number of 16-channel units
    for (g.nsplit = 1; g.nsplit <= 8; g.nsplit++) {
        if (C % g.nsplit != 0) continue;
        const int nper = C / g.nsplit;
        if (nper > 224) continue;
        if (g.nsplit > 1 && nper % 16 != 0) continue;
        g.nper = nper;
        break;
    }
    if (!g.nper) return g;
    g.npad = (g.nper + 15) / 16 * 16;
    // 64-byte rows (32 channels) for narrow stages: the 27-channel stage would otherwise move 128-byte rows with 54 useful bytes
    g.rb = g.npad <= 32 ? 64 : 128;
    g.nchunk = (C + g.rb / 2 - 1) / (g.rb / 2);
    g.ok = true;
    return g;
}

size_t weight_image_bytes(const Geom & g, int K) { return (size_t)g.nsplit * g.nchunk * K * g.npad * g.rb; }

bool pack_weights(const float * w, const Geom & g, int K, void * img, cudaStream_t stream) {
    pack_w_kernel<<<296, 256, 0, stream>>>(w, g.C, K, g.nsplit, g.nper, g.npad, g.nchunk, g.rb, (__half *)img);
    MGB_LAUNCH_CHECK();
    return true;
}

bool launch_conv(const Geom & g, const ConvArgs & a, cudaStream_t stream) {
    static std::atomic<int> n_sm[64];
    static DeviceOnce attr_done;
    int dev = 0;
    MGB_CUDA_TRY(cudaGetDevice(&dev));
    if (!n_sm[dev & 63]) { int n = 0; MGB_CUDA_TRY(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev)); n_sm[dev & 63] = n; }
    KParams p = {};
    p.a = a;
    p.C = g.C; p.nsplit = g.nsplit; p.nper = g.nper; p.npad = g.npad; p.nchunk = g.nchunk; p.cs = row_stride(g.C);
    const int halo = (a.K - 1) * a.dil;
    if (!g.ok || halo > kHP) { set_error("codec: conv shape not supported by the tensor-core path"); return false; }
    p.hp = (halo + 7) / 8 * 8;
    const int tile_bytes = g.npad * g.rb;
    p.rb = g.rb; p.desc_hi = desc_hi_for(g.rb);
    p.tps = std::max(1, std::min(a.K, (32 * 1024) / tile_bytes));
    p.ngroups = (a.K + p.tps - 1) / p.tps;
    p.wstage_bytes = p.tps * tile_bytes;
    // accumulator tiles per work item: as many as TMEM (2 buffers x R x npad <= 512 columns) and shared memory allow;
    // small channel counts amortise the per-item pipeline hand-offs and the weight stream over up to 512 time steps
    p.R = std::max(1, std::min(4, 256 / g.npad));
    if (p.R == 3) p.R = 2;
    p.abuf_bytes = ((p.R * kTile + kHP) * g.rb + 1023) / 1024 * 1024;
    p.NA = p.R >= 4 ? 2 : 3;
    p.NW = std::max(2, std::min(kMaxNW, (kSmemBudget - p.NA * p.abuf_bytes) / p.wstage_bytes));
    p.tiles_per_b = (a.T + p.R * kTile - 1) / (p.R * kTile);
    p.n_tiles = g.nsplit * a.B * p.tiles_per_b;
    p.tmem_cols = 32;
    while (p.tmem_cols < 2 * p.R * g.npad) p.tmem_cols *= 2;
    p.rows = (long long)act_rows(a.T);
    const size_t smem = 1024 + (size_t)p.NA * p.abuf_bytes + (size_t)p.NW * p.wstage_bytes + 512;
    if (smem > (size_t)kDynSmemMax) { set_error("codec: conv tile configuration exceeds shared memory"); return false; }
    if (!attr_done.done(dev)) {
        MGB_CUDA_TRY(cudaFuncSetAttribute(conv_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDynSmemMax));
        MGB_CUDA_TRY(cudaFuncSetAttribute(conv_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDynSmemMax));
        MGB_CUDA_TRY(cudaFuncSetAttribute(conv_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDynSmemMax));
        attr_done.set(dev);
    }
    if (g.C > kMaxC) { set_error("codec: too many channels for the tensor-core path"); return false; }
    const int grid = std::min(p.n_tiles, n_sm[dev & 63].load());
    const int mode = a.res ? (a.ya ? 1 : 2) : 0;
    if (mode == 0 && !a.ya) { set_error("codec: conv without an output"); return false; }
    if (mode == 1 && (!a.y || !a.ya)) { set_error("codec: residual conv needs y and ya"); return false; }
    if (mode == 2 && !a.y) { set_error("codec: residual conv needs y"); return false; }
    if (mode == 0) conv_tc_kernel<0><<<grid, kThreads, smem, stream>>>(p);
    else if (mode == 1) conv_tc_kernel<1><<<grid, kThreads, smem, stream>>>(p);
    else conv_tc_kernel<2><<<grid, kThreads, smem, stream>>>(p);
    MGB_LAUNCH_CHECK();
    return true;
}

bool launch_up(const Geom & g, const UpArgs & a, cudaStream_t stream) {
    UpKParams p = {};
    p.a = a; p.cs_in = row_stride(a.Cin); p.cs_out = row_stride(a.Cin / 2); p.nchunk = g.nchunk; p.rb = g.rb; p.rows = (long long)act_rows(a.T * a.s);
    if (a.s > 8) { set_error("codec: up-sampling stride > 8"); return false; }
    static const bool lane_per_step = getenv("MGB_UP_LANE_PER_STEP") != nullptr;      // diagnostic: the first-stage kernel everywhere
    const int ng = p.cs_out / 8;
    if (!lane_per_step && ng <= kUpNG && (a.s == 2 || a.s == 4 || a.s == 8)) {
        const int cpb = 128 / ng;
        dim3 grid(((a.T + kUpNT - 1) / kUpNT + cpb - 1) / cpb, a.B);
        if (a.s == 2) up_rows_kernel<2><<<grid, 128, 0, stream>>>(p);
        else if (a.s == 4) up_rows_kernel<4><<<grid, 128, 0, stream>>>(p);
        else up_rows_kernel<8><<<grid, 128, 0, stream>>>(p);
    } else {
        dim3 grid((a.T + 127) / 128, p.cs_out / 8, a.B);
        up_tm_kernel<<<grid, 128, 0, stream>>>(p);
    }
    MGB_LAUNCH_CHECK();
    return true;
}

bool launch_post(const PostArgs & a, cudaStream_t stream) {
    PostKParams p = {};
    p.a = a; p.cs = row_stride(a.C);
    if (a.C > kPostCS || a.K > 8) { set_error("codec: post conv shape not supported"); return false; }
    dim3 grid((a.T + 255) / 256, a.B);
    post_tm_kernel<<<grid, 256, 0, stream>>>(p);
    MGB_LAUNCH_CHECK();
    return true;
}

}  // namespace ctc
}  // namespace mgb
