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

constexpr int kThreads = 320;
constexpr int kABytes = (kTile + kHP) * 128;      // one staged activation window (23 KB)
constexpr int kMaxNA = 4, kMaxNW = 8;
constexpr int kSmemBudget = 220 * 1024;

struct KParams {
    ConvArgs a;
    int C, nsplit, nper, npad, nchunk;
    int hp;                 // staged history rows, (K-1)*dil rounded up to 8
    int tps, ngroups;       // taps per weight stage, stages per chunk
    int NA, NW;             // ring depths
    int wstage_bytes;       // bytes of a full weight stage
    int tiles_per_b, n_tiles;
    int tmem_cols;
    long long rows;         // rows per (utterance, chunk) of the activation images
};

__device__ __forceinline__ float f16r(float x) { return __half2float(__float2half_rn(x)); }

// x + sin^2(alpha x)/alpha for c < n_alpha, LeakyReLU(0.01) otherwise (nano-codec.cpp:386-417)
__device__ __forceinline__ float half_snake(float x, int c, const float * alpha, int n_alpha) {
    if (c < n_alpha) {
        const float a = __ldg(alpha + c);
        const float sn = sinf(x * a);
        return x + (sn * sn) / a;
    }
    return x > 0.0f ? x : 0.01f * x;
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
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t *>(&h);
}

__global__ void __launch_bounds__(kThreads, 1) conv_tc_kernel(const KParams p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char * base = reinterpret_cast<unsigned char *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char * abuf = base;
    unsigned char * wbuf = base + p.NA * kABytes;
    uint64_t * bars = reinterpret_cast<uint64_t *>(wbuf + (size_t)p.NW * p.wstage_bytes);
    uint64_t * a_full = bars, * a_empty = bars + kMaxNA, * w_full = bars + 2 * kMaxNA, * w_empty = w_full + kMaxNW;
    uint64_t * acc_full = w_empty + kMaxNW, * acc_empty = acc_full + 2;
    uint32_t * tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

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
    const int tile_bytes = p.npad * 128;

    if (warp == 0) {
        if (lane == 0) {
            const uint32_t abytes = (uint32_t)(kTile + p.hp) * 128;
            uint32_t ia = 0, iw = 0;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                const int tb = tile % p.tiles_per_b, hb = tile / p.tiles_per_b;
                const int h = hb % p.nsplit, b = hb / p.nsplit;
                const long long row0 = (long long)kHP + (long long)tb * kTile - p.hp;
                for (int c = 0; c < p.nchunk; c++) {
                    const int sa = ia % p.NA;
                    tc::mbar_wait(&a_empty[sa], ((ia / p.NA) & 1) ^ 1);
                    tc::mbar_expect_tx(&a_full[sa], abytes);
                    tc::bulk_g2s(abuf + sa * kABytes,
                                 reinterpret_cast<const unsigned char *>(p.a.xa) + (((long long)b * p.nchunk + c) * p.rows + row0) * 128,
                                 abytes, &a_full[sa]);
                    ia++;
                    const unsigned char * wsrc = reinterpret_cast<const unsigned char *>(p.a.w) + ((size_t)(h * p.nchunk + c) * K) * tile_bytes;
                    for (int g = 0; g < p.ngroups; g++) {
                        const int sw = iw % p.NW;
                        const int nt = min(p.tps, K - g * p.tps);
                        tc::mbar_wait(&w_empty[sw], ((iw / p.NW) & 1) ^ 1);
                        tc::mbar_expect_tx(&w_full[sw], (uint32_t)(nt * tile_bytes));
                        tc::bulk_g2s(wbuf + (size_t)sw * p.wstage_bytes, wsrc + (size_t)g * p.tps * tile_bytes, (uint32_t)(nt * tile_bytes), &w_full[sw]);
                        iw++;
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = tc::umma_idesc_f16(kTile, p.npad);
            uint32_t ia = 0, iw = 0, it = 0;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, it++) {
                const int buf = it & 1;
                tc::mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t dcol = tmem_base + (uint32_t)(buf * p.npad);
                uint32_t first = 1;
                for (int c = 0; c < p.nchunk; c++) {
                    const int sa = ia % p.NA;
                    tc::mbar_wait(&a_full[sa], (ia / p.NA) & 1);
                    const uint32_t a0 = tc::smem_u32(abuf + sa * kABytes);
                    const int nk16 = min(4, (p.C - c * 64 + 15) >> 4);
                    for (int g = 0; g < p.ngroups; g++) {
                        const int sw = iw % p.NW;
                        const int nt = min(p.tps, K - g * p.tps);
                        tc::mbar_wait(&w_full[sw], (iw / p.NW) & 1);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t w0 = tc::smem_u32(wbuf + (size_t)sw * p.wstage_bytes);
                        for (int q = 0; q < nt; q++) {
                            const int k = g * p.tps + q;
                            const uint32_t ak = a0 + (uint32_t)(p.hp - (K - 1 - k) * p.a.dil) * 128;
                            const uint32_t wk = w0 + (uint32_t)(q * tile_bytes);
                            for (int j = 0; j < nk16; j++) {
                                tc::umma_bf16(dcol, tc::umma_desc_sw128(ak + j * 32), tc::umma_desc_sw128(wk + j * 32), idesc, first ^ 1u);
                                first = 0;
                            }
                        }
                        tc::umma_commit(&w_empty[sw]);
                        iw++;
                    }
                    tc::umma_commit(&a_empty[sa]);
                    ia++;
                }
                tc::umma_commit(&acc_full[buf]);
            }
        }
    } else {
        const int q = warp & 3;                        // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;              // column half
        const int ngrp = p.npad >> 4;                  // 16-column groups
        const int g_lo = half == 0 ? 0 : (ngrp + 1) / 2, g_hi = half == 0 ? (ngrp + 1) / 2 : ngrp;
        const long long hrow_bytes = p.rows * 128;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, it++) {
            const int buf = it & 1;
            const int tb = tile % p.tiles_per_b, hb = tile / p.tiles_per_b;
            const int h = hb % p.nsplit, b = hb / p.nsplit;
            const int t = tb * kTile + q * 32 + lane;
            const bool tv = t < T;
            tc::mbar_wait(&acc_full[buf], (it >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * p.npad);
            const int r_img = kHP + t;
            for (int g = g_lo; g < g_hi; g++) {
                uint32_t v[16];
                tmem_ld16(trow + g * 16, v);
#pragma unroll
                for (int s = 0; s < 2; s++) {
                    const int n0 = g * 16 + s * 8;
                    const int co0 = h * p.nper + n0;
                    float act[8];
#pragma unroll
                    for (int e = 0; e < 8; e++) {
                        const int co = co0 + e;
                        const bool cv = (n0 + e) < p.nper;
                        float y = 0.0f;
                        if (cv) {
                            y = __uint_as_float(v[s * 8 + e]) + __ldg(p.a.bias + co);
                            if (tv) {
                                const size_t o = ((size_t)b * p.C + co) * T + t;
                                if (p.a.res) y = p.a.res[o] + y;
                                if (p.a.y) p.a.y[o] = y;
                                if (p.a.sum_mode == 1) p.a.sum_out[o] = y;
                                else if (p.a.sum_mode == 2) p.a.sum_out[o] = p.a.sum_in[o] + y;
                                else if (p.a.sum_mode == 3) p.a.sum_out[o] = (p.a.sum_in[o] + y) * (1.0f / 3.0f);
                            }
                        }
                        act[e] = (cv && p.a.ya) ? half_snake(y, co, p.a.alpha2, p.a.n_alpha2) : 0.0f;
                    }
                    const bool store_grp = p.nsplit == 1 ? true : n0 < p.nper;
                    if (p.a.ya && tv && store_grp) {
                        uint4 pk;
                        pk.x = pack_h2(act[0], act[1]); pk.y = pack_h2(act[2], act[3]);
                        pk.z = pack_h2(act[4], act[5]); pk.w = pack_h2(act[6], act[7]);
                        unsigned char * dst = reinterpret_cast<unsigned char *>(p.a.ya) + ((long long)b * p.nchunk + (co0 >> 6)) * hrow_bytes +
                                              (long long)r_img * 128 + ((((co0 & 63) >> 3) ^ (r_img & 7)) << 4);
                        *reinterpret_cast<uint4 *>(dst) = pk;
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(&acc_empty[buf]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols));
}

// f32 (C, C, K) -> f16 tile images [nsplit][nchunk][K][npad x 64]; one thread per 16-byte group of 8 input channels
__global__ void pack_w_kernel(const float * w, int C, int K, int nsplit, int nper, int npad, int nchunk, __half * out) {
    const size_t total = (size_t)nsplit * nchunk * K * npad * 8;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int kc = (int)(i % 8);
        size_t r = i / 8;
        const int n = (int)(r % npad); r /= npad;
        const int k = (int)(r % K); r /= K;
        const int c = (int)(r % nchunk); const int h = (int)(r / nchunk);
        const int co = h * nper + n;
        __half hv[8];
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const int ci = c * 64 + kc * 8 + e;
            hv[e] = __float2half_rn((n < nper && ci < C) ? w[((size_t)co * C + ci) * K + k] : 0.0f);
        }
        unsigned char * dst = reinterpret_cast<unsigned char *>(out) + (((size_t)(h * nchunk + c) * K + k) * npad) * 128 + tc::swz_offset(n, kc * 8);
        *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(hv);
    }
}

struct SnakeKParams { SnakeArgs a; int nchunk; long long rows; };
// grid (time blocks of 128, 8-channel groups, B); thread = one time step
__global__ void __launch_bounds__(128) snake_images_kernel(const SnakeKParams p) {
    const int t = blockIdx.x * 128 + threadIdx.x;
    if (t >= p.a.T) return;
    const int g8 = blockIdx.y, b = blockIdx.z;
    const int c0 = g8 * 8;
    float x[8];
#pragma unroll
    for (int e = 0; e < 8; e++) x[e] = (c0 + e) < p.a.C ? p.a.x[((size_t)b * p.a.C + c0 + e) * p.a.T + t] : 0.0f;
    const int r_img = kHP + t;
    const long long off = ((long long)b * p.nchunk + (c0 >> 6)) * p.rows * 128 + (long long)r_img * 128 + ((((c0 & 63) >> 3) ^ (r_img & 7)) << 4);
    for (int j = 0; j < p.a.n_out; j++) {
        float act[8];
#pragma unroll
        for (int e = 0; e < 8; e++) act[e] = (c0 + e) < p.a.C ? half_snake(x[e], c0 + e, p.a.alpha[j], p.a.n_alpha) : 0.0f;
        uint4 pk;
        pk.x = pack_h2(act[0], act[1]); pk.y = pack_h2(act[2], act[3]);
        pk.z = pack_h2(act[4], act[5]); pk.w = pack_h2(act[6], act[7]);
        *reinterpret_cast<uint4 *>(reinterpret_cast<unsigned char *>(p.a.out[j]) + off) = pk;
    }
}

}  // namespace

Geom geom_for(int C) {
    Geom g;
    g.C = C;
    g.nsplit = (C + 223) / 224;
    if (C <= 0 || C % g.nsplit != 0) return g;
    g.nper = C / g.nsplit;
    if (g.nsplit > 1 && (g.nper % 8 != 0 || C % 16 != 0)) return g;
    g.npad = (g.nper + 15) / 16 * 16;
    if (g.npad < 16 || g.npad > 256) return g;
    g.nchunk = (C + 63) / 64;
    g.ok = true;
    return g;
}

size_t weight_image_bytes(const Geom & g, int K) { return (size_t)g.nsplit * g.nchunk * K * g.npad * 128; }

bool pack_weights(const float * w, const Geom & g, int K, void * img, cudaStream_t stream) {
    pack_w_kernel<<<296, 256, 0, stream>>>(w, g.C, K, g.nsplit, g.nper, g.npad, g.nchunk, (__half *)img);
    MGB_LAUNCH_CHECK();
    return true;
}

bool launch_conv(const Geom & g, const ConvArgs & a, cudaStream_t stream) {
    static int n_sm[64] = {};
    static uint64_t attr_done = 0;
    int dev = 0;
    MGB_CUDA_TRY(cudaGetDevice(&dev));
    if (!n_sm[dev]) MGB_CUDA_TRY(cudaDeviceGetAttribute(&n_sm[dev], cudaDevAttrMultiProcessorCount, dev));
    KParams p = {};
    p.a = a;
    p.C = g.C; p.nsplit = g.nsplit; p.nper = g.nper; p.npad = g.npad; p.nchunk = g.nchunk;
    const int halo = (a.K - 1) * a.dil;
    if (!g.ok || halo > kHP) { set_error("codec: conv shape not supported by the tensor-core path"); return false; }
    p.hp = (halo + 7) / 8 * 8;
    const int tile_bytes = g.npad * 128;
    p.tps = std::max(1, std::min(a.K, (32 * 1024) / tile_bytes));
    p.ngroups = (a.K + p.tps - 1) / p.tps;
    p.wstage_bytes = p.tps * tile_bytes;
    p.NA = 3;
    p.NW = std::max(2, std::min(kMaxNW, (kSmemBudget - p.NA * kABytes) / p.wstage_bytes));
    p.tiles_per_b = (a.T + kTile - 1) / kTile;
    p.n_tiles = g.nsplit * a.B * p.tiles_per_b;
    p.tmem_cols = 32;
    while (p.tmem_cols < 2 * g.npad) p.tmem_cols *= 2;
    p.rows = (long long)act_rows(a.T);
    const size_t smem = 1024 + (size_t)p.NA * kABytes + (size_t)p.NW * p.wstage_bytes + 512;
    if (!(attr_done >> dev & 1)) {
        MGB_CUDA_TRY(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_done |= 1ull << dev;
    }
    const int grid = std::min(p.n_tiles, n_sm[dev]);
    conv_tc_kernel<<<grid, kThreads, smem, stream>>>(p);
    MGB_LAUNCH_CHECK();
    return true;
}

bool launch_snake_images(const Geom & g, const SnakeArgs & a, cudaStream_t stream) {
    SnakeKParams p = {};
    p.a = a; p.nchunk = g.nchunk; p.rows = (long long)act_rows(a.T);
    dim3 grid((a.T + 127) / 128, g.nchunk * 8, a.B);
    snake_images_kernel<<<grid, 128, 0, stream>>>(p);
    MGB_LAUNCH_CHECK();
    return true;
}

}  // namespace ctc
}  // namespace mgb
