// Batched linear layers on the 5th-generation tensor cores: LayerNorm / bf16 hi-lo packing of the activations into
// shared-memory tile images, tcgen05.mma mainloop (gemm_tc.cuh), fused bias / GELU / residual / KV-cache epilogue.
// Replaces linear_kernel (transformer_kernels.cu) for bf16 models when a launch carries >= 16 tokens: batched
// decoder steps, the 110-frame context prefill and the cross-attention K/V precompute
// (reference: ggml_mul_mat call sites src/magpie.cpp:3415, 3464-3479, 1733, 1764, 1796, 1805).
#include <cstdlib>

#include <cuda_fp16.h>
#include "common.cuh"
#include "gemm_tc.cuh"
#include "kernels.cuh"

namespace mgb {

namespace {

using bf = __nv_bfloat16;

// row-major bf16 [N][K] -> [ceil(N/128)][K/64] tiles of 128 x 64 in the SWIZZLE_128B K-major image; one thread per 16 bytes
// taps > 1 (causal conv weights stored tap-major [taps][N][K]): the taps are concatenated along k, k' = tap * K + k
__global__ void pack_w_kernel(const bf * W, int N, int Ktap, int taps, bf * Wt, int as_f16) {
    const int K = Ktap * taps;
    const int KT = K / 64, NT = (N + tc::BM - 1) / tc::BM;
    const size_t total = (size_t)NT * tc::BM * (K / 8);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / (K / 8)), kc = (int)(i % (K / 8));
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        const int tap = (kc * 8) / Ktap, kk = kc * 8 - tap * Ktap;
        if (r < N) v = *reinterpret_cast<const uint4 *>(W + ((size_t)tap * N + r) * Ktap + kk);
        if (as_f16) {                               // bf16 -> f16: exact for normal f16 magnitudes (8 mantissa bits fit into 11)
            uint32_t * w = reinterpret_cast<uint32_t *>(&v);
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const float lo = __uint_as_float(w[q] << 16), hi = __uint_as_float(w[q] & 0xffff0000u);
                const __half2 h = __floats2half2_rn(lo, hi);
                w[q] = *reinterpret_cast<const uint32_t *>(&h);
            }
        }
        const size_t tile = (size_t)(r / tc::BM) * KT + kc / 8;
        unsigned char * dst = reinterpret_cast<unsigned char *>(Wt) + tile * (tc::BM * 128) + tc::swz_offset(r % tc::BM, (kc % 8) * 8);
        *reinterpret_cast<uint4 *>(dst) = v;
    }
}

// activations f32 [M][ldx] (optionally LayerNorm'd, magpie.cpp:2237-2259) -> hi / lo bf16 tile images [ceil(M/MT)][K/64][MT x 64].
// One CTA per row of the padded token range; rows >= M are written as zeros.
// taps > 1: row m of the packed operand is [x(m - (taps-1)) | ... | x(m)] (each LayerNorm'd on its own), a shifted row being
// zero when it would cross the start of the utterance (tok_pos[m] < shift): the k = 3 causal conv as ONE GEMM over 3 K.
__global__ void __launch_bounds__(256) pack_x_kernel(const float * X, int ldx, int M, int K, const float * ln_w, float eps, int MT,
                                                     bf * hi, bf * lo, int taps, const int32_t * tok_pos, int f16) {
    __shared__ float red[32];
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");     // let the GEMM's CTAs start prefetching weights
    asm volatile("griddepcontrol.wait;" ::: "memory");                  // (itself launched as a programmatic dependent of the previous kernel)
    const int m = blockIdx.x, tid = threadIdx.x;
    const int KT = K * taps / 64;
    const size_t tile_row = (size_t)(m / MT) * KT;
    for (int tap = 0; tap < taps; tap++) {
        const int shift = taps - 1 - tap;
        const bool valid = m < M && (shift == 0 || tok_pos[m] >= shift);
        const float * xr = X + (size_t)(valid ? m - shift : 0) * ldx;
        float mean = 0.0f, scale = 1.0f;
        if (ln_w) {
            float s = 0.0f;
            for (int k = tid; k < K; k += 256) s += valid ? xr[k] : 0.0f;
            mean = block_sum(s, red) / (float)K;
            float s2 = 0.0f;
            for (int k = tid; k < K; k += 256) { const float v = (valid ? xr[k] : 0.0f) - mean; s2 += v * v; }
            const float var = block_sum(s2, red) / (float)K;
            scale = 1.0f / sqrtf(var + eps);
        }
        for (int kc = tid; kc < K / 8; kc += 256) {
            float v[8];
            if (valid) {
                const float4 a = *reinterpret_cast<const float4 *>(xr + kc * 8), b = *reinterpret_cast<const float4 *>(xr + kc * 8 + 4);
                v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
                if (ln_w) {
#pragma unroll
                    for (int q = 0; q < 8; q++) v[q] = ((v[q] - mean) * scale) * ln_w[kc * 8 + q];
                }
            } else {
#pragma unroll
                for (int q = 0; q < 8; q++) v[q] = 0.0f;
            }
            uint32_t h[4], l[4];
#pragma unroll
            for (int q = 0; q < 4; q++) tc::pack_act2(v[2 * q], v[2 * q + 1], f16 != 0, h[q], l[q]);
            const int kg = tap * (K / 8) + kc;                       // 8-column group of the concatenated row
            const size_t off = (tile_row + kg / 8) * ((size_t)MT * 128) + tc::swz_offset(m % MT, (kg % 8) * 8);
            *reinterpret_cast<uint4 *>(reinterpret_cast<unsigned char *>(hi) + off) = make_uint4(h[0], h[1], h[2], h[3]);
            if (!f16) *reinterpret_cast<uint4 *>(reinterpret_cast<unsigned char *>(lo) + off) = make_uint4(l[0], l[1], l[2], l[3]);
        }
    }
}

struct TcEpi {
    int N, M;
    const float * bias;
    const float * res; int ldr;
    float * Y; int ldy;
    int act, gelu_f16;
    int n_q, dkv; bf * kdst; bf * vdst; const int32_t * tok_slot;
    bf * pk_hi; bf * pk_lo;          // optional: the output as hi / lo tile images [N/64][64 x 64] for a following GEMM over K' = N (M <= 64)
};

// SPLIT > 1: split-K over a thread-block cluster of SPLIT CTAs (cluster dims 1 x 1 x SPLIT, rank = blockIdx.z).  Every CTA
// accumulates its share of the k tiles in its own TMEM; ranks > 0 then push their accumulator into rank 0's shared memory
// through DSMEM, and rank 0 adds them in rank order (deterministic) and runs the epilogue.  At M <= 64 a GEMM launch is
// bound by how fast ONE SM can ingest its CTA's weight + activation tiles, so the split shortens the critical path.
template <int SPLIT> struct StageCap { static constexpr int value = SPLIT == 1 ? 8 : (SPLIT == 2 ? 5 : 3); };

// EPI specialises the epilogue so that a decoder-step GEMM carries only the code it runs (these launches are latency
// bound and start with cold instruction caches: code size is time): 0 = generic, 1 = q | K | V split store (no
// activation / residual), 2 = residual add into an f32 row, 3 = GELU + packed hi | lo output for the next GEMM.
enum { EPI_GENERIC = 0, EPI_QKV = 1, EPI_RES = 2, EPI_GELU_PACK = 3 };

template <int MT, int SPLIT, int EPI>
__global__ void __launch_bounds__(tc::kThreads, 1) tc_linear_kernel(const bf * Wt, const bf * Xhi, const bf * Xlo, int KT, const TcEpi e) {
    constexpr bool kRes = EPI == EPI_GENERIC || EPI == EPI_RES;
    constexpr bool kAct = EPI == EPI_GENERIC || EPI == EPI_GELU_PACK;
    constexpr bool kPack = EPI == EPI_GENERIC || EPI == EPI_GELU_PACK;
    constexpr bool kSplitStore = EPI == EPI_GENERIC || EPI == EPI_QKV;
    constexpr bool kPlainStore = EPI != EPI_GELU_PACK;
    extern __shared__ unsigned char tc_smem[];
    constexpr int SCAP = StageCap<SPLIT>::value;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");      // let a dependent kernel start its own prefetching
    const int nt = blockIdx.x, mt = blockIdx.y;
    const int rank = SPLIT > 1 ? (int)blockIdx.z : 0;
    const int kt0 = rank * KT / SPLIT, kt1 = (rank + 1) * KT / SPLIT;
    const uint32_t tmem = tc::mainloop<MT, 2, false, true, SCAP>(tc_smem, Wt, Xhi, Xlo, KT, nt, mt, kt0, kt1 - kt0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // exchange buffer of rank 0: [SPLIT-1][MT columns][128 rows] f32, behind the stage ring and the barriers
    unsigned char * tiles = reinterpret_cast<unsigned char *>(((uintptr_t)tc_smem + 1023) & ~(uintptr_t)1023);
    float * xbuf = reinterpret_cast<float *>(tiles + tc::Smem<MT, 2, SCAP>::kStages * tc::Smem<MT, 2, SCAP>::kStageBytes + 256);
    if (SPLIT > 1) {
        if (rank > 0 && warp >= 2) {
            const int q = warp & 3;
            for (int c = 0; c < MT; c += 32) {
                uint32_t v[32];
                tc::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + c, v);
                const uint32_t local = tc::smem_u32(xbuf + ((size_t)(rank - 1) * MT + c) * 128 + q * 32 + lane);
                uint32_t remote;
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(0));
#pragma unroll
                for (int j = 0; j < 32; j++) asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(remote + j * 128 * 4), "r"(v[j]) : "memory");
            }
        }
        __syncwarp();
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    if (warp >= 2 && rank == 0) {
        const int q = warp & 3;                      // TMEM lane quarter this warp may access
        const int n = nt * tc::BM + q * 32 + lane;
        const float bias = (e.bias && n < e.N) ? e.bias[n] : 0.0f;
        for (int c = 0; c < MT; c += 32) {
            if (mt * MT + c >= e.M) break;           // warp-uniform
            uint32_t v[32];
            tc::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + c, v);
            if (SPLIT > 1) {
#pragma unroll
                for (int r = 1; r < SPLIT; r++)
#pragma unroll
                    for (int j = 0; j < 32; j++)
                        v[j] = __float_as_uint(__uint_as_float(v[j]) + xbuf[((size_t)(r - 1) * MT + c + j) * 128 + q * 32 + lane]);
            }
            if (n < e.N) {
                // residual values first (Y may alias res for the in-place x += W h updates, which would otherwise
                // serialise 32 dependent load -> store round trips)
                float r[kRes ? 32 : 1];
                if (kRes && e.res) {
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        const int m = mt * MT + c + j;
                        r[j] = m < e.M ? e.res[(size_t)m * e.ldr + n] : 0.0f;
                    }
                }
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    const int m = mt * MT + c + j;
                    if (m < e.M) {
                        float y = __uint_as_float(v[j]) + bias;
                        if (kAct && (EPI == EPI_GELU_PACK || e.act == ACT_GELU)) y = EPI == EPI_GELU_PACK ? gelu_ggml_fast(y, e.gelu_f16) : gelu_ggml(y, e.gelu_f16);
                        if (kRes && e.res) y += r[j];
                        if (kPack && (EPI == EPI_GELU_PACK || e.pk_hi)) {
                            const bf h = __float2bfloat16_rn(y);
                            const size_t off = (size_t)(n >> 6) * (64 * 128) + tc::swz_offset(m, n & 63);
                            *reinterpret_cast<bf *>(reinterpret_cast<unsigned char *>(e.pk_hi) + off) = h;
                            *reinterpret_cast<bf *>(reinterpret_cast<unsigned char *>(e.pk_lo) + off) = __float2bfloat16_rn(y - __bfloat162float(h));
                        }
                        if (!kPlainStore || !e.Y) continue;
                        if (!kSplitStore || e.n_q < 0 || n < e.n_q) e.Y[(size_t)m * e.ldy + n] = y;
                        else {
                            const size_t slot = (size_t)e.tok_slot[m] * e.dkv;
                            const int cc = n - e.n_q;
                            if (cc < e.dkv) e.kdst[slot + cc] = __float2bfloat16_rn(y);
                            else e.vdst[slot + (cc - e.dkv)] = __float2bfloat16_rn(y);
                        }
                    }
                }
            }
        }
    }
    tc::finish<MT>(tmem);
}

template <int MT, int SPLIT, int EPI> bool launch_tc(const bf * Wt, const bf * hi, const bf * lo, int KT, const TcEpi & e, cudaStream_t stream) {
    static DeviceOnce attr_done;
    int dev = 0;
    MGB_CUDA_TRY(cudaGetDevice(&dev));
    constexpr int SCAP = StageCap<SPLIT>::value;
    constexpr int smem = tc::Smem<MT, 2, SCAP>::kBytes + (SPLIT - 1) * MT * 128 * 4;
    static_assert(smem <= 227 * 1024, "tc_linear shared memory");
    if (!attr_done.done(dev)) {
        MGB_CUDA_TRY(cudaFuncSetAttribute(tc_linear_kernel<MT, SPLIT, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_done.set(dev);
    }
    // launched as a programmatic dependent of the activation-packing kernel: CTAs start (and prefetch weight tiles)
    // while pack_x_kernel is still running, and wait for it with griddepcontrol.wait before touching its output
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((e.N + tc::BM - 1) / tc::BM, (e.M + MT - 1) / MT, SPLIT); cfg.blockDim = dim3(tc::kThreads);
    cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    at[1].id = cudaLaunchAttributeClusterDimension;
    at[1].val.clusterDim.x = 1; at[1].val.clusterDim.y = 1; at[1].val.clusterDim.z = SPLIT;
    cfg.attrs = at; cfg.numAttrs = SPLIT > 1 ? 2 : 1;
    MGB_CUDA_TRY(cudaLaunchKernelEx(&cfg, tc_linear_kernel<MT, SPLIT, EPI>, Wt, hi, lo, KT, e));
    MGB_LAUNCH_CHECK();
    return true;
}

}  // namespace

size_t tc_weight_tile_bytes(int N, int K) { return (size_t)((N + tc::BM - 1) / tc::BM) * tc::BM * K * sizeof(bf); }      // K = taps * K_tap

bool tc_pack_weights(const void * W, int N, int K, void * Wt, cudaStream_t stream, int taps, bool as_f16) {
    pack_w_kernel<<<592, 256, 0, stream>>>((const bf *)W, N, K, taps, (bf *)Wt, as_f16 ? 1 : 0);
    MGB_LAUNCH_CHECK();
    return true;
}

size_t tc_scratch_bytes(int M, int K) { return 2 * (size_t)((M + 127) / 128) * 128 * K * sizeof(bf); }

bool tc_linear_supported(const LinearArgs & a) {
    // fewer tokens: CUDA-core skinny GEMM (linear_kernel).  Measured per decoder step, CUDA-core vs tensor-core chain: 2 utterances
    // 676 vs 704 us, 4: 834 vs 704, 8: 1467 vs 824, 12: 1341 vs 773 (round 1 switched at 16)
    static const int min_m = getenv("MGB_TC_MIN_M") ? atoi(getenv("MGB_TC_MIN_M")) : 4;
    return a.precision == MGB_PREC_BF16 && a.W.tiles != nullptr && (a.W.taps == 1 || a.tok_pos != nullptr) && a.M >= min_m && a.W.K % 64 == 0 &&
           a.tc_scratch != nullptr && tc_scratch_bytes(a.M, a.W.K * a.W.taps) <= a.tc_scratch_bytes && (a.ldx % 4) == 0;
}

bool launch_linear_tc(const LinearArgs & a, cudaStream_t stream) {
    const int Ktap = a.W.K, taps = a.W.taps, K = Ktap * taps, M = a.M;       // taps > 1: one GEMM over the concatenated taps
    static const int mt64_max = [] { const char * e = getenv("MGB_TC_MT64_MAX"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 128; }();
    const int MT = M <= mt64_max ? 64 : 128;     // up to two 64-token tiles with cluster split-K: shorter per-SM ingest chains than one 128-token tile
    const int Mpad = (M + MT - 1) / MT * MT;
    bf * hi = (bf *)a.tc_scratch;
    bf * lo = hi + (size_t)Mpad * K;
    if (!a.x_prepacked) {
        cudaLaunchConfig_t pc = {};
        pc.gridDim = dim3(Mpad); pc.blockDim = dim3(256); pc.dynamicSmemBytes = 0; pc.stream = stream;
        cudaLaunchAttribute pa[1];
        pa[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        pa[0].val.programmaticStreamSerializationAllowed = 1;
        pc.attrs = pa; pc.numAttrs = 1;
        MGB_CUDA_TRY(cudaLaunchKernelEx(&pc, pack_x_kernel, a.X, a.ldx, M, Ktap, a.ln_w, a.eps, MT, hi, lo, taps, a.tok_pos, a.act_f16 ? 1 : 0));
        MGB_LAUNCH_CHECK();
    }
    TcEpi e;
    e.pk_hi = nullptr; e.pk_lo = nullptr;
    if (a.pack_out) {
        if (M > 64 || a.W.N % 64 != 0 || a.n_q >= 0) { set_error("linear: packed output needs one token tile and N % 64 == 0"); return false; }
        e.pk_hi = (bf *)a.pack_out; e.pk_lo = e.pk_hi + (size_t)64 * a.W.N;
    }
    e.N = a.W.N; e.M = M; e.bias = a.bias; e.res = a.res; e.ldr = a.ldr; e.Y = a.Y; e.ldy = a.ldy; e.act = a.act; e.gelu_f16 = a.gelu_f16;
    e.n_q = a.n_q; e.dkv = a.dkv; e.kdst = (bf *)a.kdst; e.vdst = (bf *)a.vdst; e.tok_slot = a.tok_slot;
    if (MT == 64 && ts_linear_supported(a)) return launch_linear_ts(a, hi, lo, stream);      // decoder-step GEMMs: token-stationary kernel (gemm_ts.cu)
    if (a.ln_fold_stats || a.next_ln_w || a.act_f16 || a.pack_f16) { set_error("linear: a folded LayerNorm / f16 activation images need the token-stationary GEMM (gemm_ts.cu)"); return false; }
    if (MT == 64) {
        // one token tile: split K over a cluster (deterministic DSMEM reduction) to shorten the per-SM ingest chain
        const int KT = K / 64;
        static const bool no_split = getenv("MGB_NO_SPLITK") != nullptr;
        const bf * W = (const bf *)a.W.tiles;
        const bool qkv = a.n_q >= 0 && !a.res && a.act == ACT_NONE && !a.pack_out;
        const bool resid = a.n_q < 0 && a.res && a.act == ACT_NONE && !a.pack_out && a.Y;
        const bool gpack = a.n_q < 0 && !a.res && a.act == ACT_GELU && a.pack_out && !a.Y;
        if (!no_split && KT >= 32) {
            if (resid) return launch_tc<64, 4, EPI_RES>(W, hi, lo, KT, e, stream);
            return launch_tc<64, 4, EPI_GENERIC>(W, hi, lo, KT, e, stream);
        }
        static const bool mid2 = getenv("MGB_SPLIT_MID2") != nullptr;     // diagnostic: 2-way instead of 4-way split for K = 512..1984
        if (!no_split && !mid2 && KT >= 8 && KT % 4 == 0) {
            if (qkv) return launch_tc<64, 4, EPI_QKV>(W, hi, lo, KT, e, stream);
            if (resid) return launch_tc<64, 4, EPI_RES>(W, hi, lo, KT, e, stream);
            if (gpack) return launch_tc<64, 4, EPI_GELU_PACK>(W, hi, lo, KT, e, stream);
        }
        if (!no_split && KT >= 8) {
            if (qkv) return launch_tc<64, 2, EPI_QKV>(W, hi, lo, KT, e, stream);
            if (resid) return launch_tc<64, 2, EPI_RES>(W, hi, lo, KT, e, stream);
            if (gpack) return launch_tc<64, 2, EPI_GELU_PACK>(W, hi, lo, KT, e, stream);
            return launch_tc<64, 2, EPI_GENERIC>(W, hi, lo, KT, e, stream);
        }
        return launch_tc<64, 1, EPI_GENERIC>(W, hi, lo, KT, e, stream);
    }
    return launch_tc<128, 1, EPI_GENERIC>((const bf *)a.W.tiles, hi, lo, K / 64, e, stream);
}

}  // namespace mgb
