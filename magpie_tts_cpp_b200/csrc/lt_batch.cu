// Local transformer + sampler for MANY utterances at once: one persistent cooperative kernel, weight-stationary.
//
// Same arithmetic as lt_kernel.cu (reference src/magpie.cpp:946-1048 LT layer, 1072-1109 sample_top_k, 1113-1317
// magpie_local_transformer_sample_all), different parallelisation.  lt_kernel runs one cluster per utterance and
// re-streams the 10 MB of LT weights from L2 for every utterance (64 utterances: 0.66 GB per step, 764 us).  Here every
// CTA keeps a fixed slice of the rows of EVERY LT matrix in shared memory for the whole launch (10.3 MB bf16 / 148 CTAs,
// kept as f32: 141 KB), and each phase computes those rows for ALL utterances from an activation tile staged through shared memory:
// each weight element is read from HBM/L2 once per step, the 8 x 7 dependent phases are separated by grid barriers.
//
//   per codebook:  QKV (LN prologue) | attention + folded O-projection + residual (utterance-owner CTAs) | FF1 (LN prologue, GELU) |
//                  FF2 + residual | out-projection + bias | mask / argmax / top-k sample / feedback gather (owner CTAs)
#include <cooperative_groups.h>
#include <cstdlib>

#include "lt_common.cuh"

namespace cg = cooperative_groups;

namespace mgb {

namespace {

using namespace lt;
using bf = __nv_bfloat16;

constexpr int kTileFloats = 64 * 260;          // activation tile staged per pass (65 KB): 64 utterances x 256 columns, padded rows

struct BParams {
    LtParams p;
    float * seq, * q, * kc, * vc, * att, * x1, * ffh, * hout, * logits;     // [B][..] f32 scratch
    const float * qkv_tab;           // [7][V][3L] f32: [q | k | vo] of position cb+1 per fed code of codebook cb (model.cu)
    const void * qkvo;               // [4L][L] bf16: [Wq; Wk; hi(Wo Wv); lo(Wo Wv)]
    unsigned long long * dbg;        // MGB_LT_DBG: globaltimer stamps of CTA 0 at every phase boundary
};

struct Slice { int n0, n1; };
__device__ __forceinline__ Slice slice_of(int N, int c, int G) { return {(int)((long long)c * N / G), (int)((long long)(c + 1) * N / G)}; }

// sampler scratch, aliased onto the activation tile
struct SampSmem {
    float logits[kV];
    float sel_v[kV]; int sel_i[kV];
    float srt_v[kV]; int srt_i[kV];
    float red[32]; int redi[32];
    unsigned hist[256];
    int misc[8];
};
static_assert(sizeof(SampSmem) <= kTileFloats * 4, "sampler scratch must fit the activation tile");

// rows [n0, n1) of W (bf16 [N][K], row-major) -> shared memory as f32 (converted once per launch: the inner loops then
// issue one 16-byte broadcast load per 4 weights and no conversions)
__device__ __forceinline__ void load_rows(float * dst, const void * W, int K, Slice s) {
    const uint4 * src = reinterpret_cast<const uint4 *>(reinterpret_cast<const bf *>(W) + (size_t)s.n0 * K);
    const int n16 = (s.n1 - s.n0) * K / 8;
    for (int i = threadIdx.x; i < n16; i += kLtThreads) {
        const uint4 v = src[i];
        reinterpret_cast<float4 *>(dst)[2 * i] = make_float4(bf16lo(v.x), bf16hi(v.x), bf16lo(v.y), bf16hi(v.y));
        reinterpret_cast<float4 *>(dst)[2 * i + 1] = make_float4(bf16lo(v.z), bf16hi(v.z), bf16lo(v.w), bf16hi(v.w));
    }
}

// One GEMV phase: for every utterance u and every row n of this CTA's slice, epi(u, n, sum_k W[n][k] * X[u][k]).
// 64 utterances x 256 input columns are staged per pass (row stride 260 floats: conflict-free 16-byte reads with one
// utterance per lane); warp w owns utterance group (w & 1) and the slice rows (w >> 1), (w >> 1) + 8: no shuffles, the
// weights are shared-memory broadcasts.  `prep(u, row)` may transform a staged row in place (position add / LayerNorm;
// K == 256 only) and is executed by one warp per row.
constexpr int kKC = 256, kKS = kKC + 4, kUB = 64;
static_assert(kUB * kKS <= kTileFloats, "stage tile");
template <bool PREP, typename Prep, typename Epi>
__device__ __forceinline__ void gemv_phase(const float * Ws, Slice s, int K, const float * X, int ldx, int B, float * tile, Prep prep, Epi epi) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ug = warp & 1, rw = warp >> 1;
    const int R = s.n1 - s.n0;
    for (int ub = 0; ub < B; ub += kUB) {
        const int nu = min(kUB, B - ub);
        float acc[2] = {0.0f, 0.0f};
        for (int kc = 0; kc < K; kc += kKC) {
            __syncthreads();
            for (int i = threadIdx.x; i < nu * (kKC / 4); i += kLtThreads) {
                const int u = i / (kKC / 4), k = (i % (kKC / 4)) * 4;
                *reinterpret_cast<float4 *>(tile + u * kKS + k) = *reinterpret_cast<const float4 *>(X + (size_t)(ub + u) * ldx + kc + k);
            }
            __syncthreads();
            if (PREP) {
                for (int u = warp; u < nu; u += kLtWarps) prep(ub + u, tile + u * kKS);
                __syncthreads();
            }
            if (rw < R && ug * 32 < nu) {
                const float * xr = tile + (ug * 32 + min(lane, nu - ug * 32 - 1)) * kKS;      // lanes past the batch repeat the last row
                const float * w0 = Ws + (size_t)rw * K + kc;
                const bool two = rw + 8 < R;
                const float * w1 = two ? w0 + (size_t)8 * K : w0;
#pragma unroll 4
                for (int k = 0; k < kKC; k += 4) {
                    const float4 x = *reinterpret_cast<const float4 *>(xr + k);
                    const float4 a = *reinterpret_cast<const float4 *>(w0 + k), b = *reinterpret_cast<const float4 *>(w1 + k);
                    acc[0] = fmaf(a.x, x.x, acc[0]); acc[0] = fmaf(a.y, x.y, acc[0]);
                    acc[0] = fmaf(a.z, x.z, acc[0]); acc[0] = fmaf(a.w, x.w, acc[0]);
                    acc[1] = fmaf(b.x, x.x, acc[1]); acc[1] = fmaf(b.y, x.y, acc[1]);
                    acc[1] = fmaf(b.z, x.z, acc[1]); acc[1] = fmaf(b.w, x.w, acc[1]);
                }
            }
        }
        const int u = ub + ug * 32 + lane;
        if (rw < R && u < B) {
            epi(u, s.n0 + rw, acc[0]);
            if (rw + 8 < R) epi(u, s.n0 + rw + 8, acc[1]);
        }
    }
}

// in-place LayerNorm of one staged row of kKC = 256 values by one warp (magpie.cpp:2237-2259: mean, centred variance,
// * weight).  Lane l owns elements l, l + 32, ...; `add` (position embedding or zeros) and the weights `w` are the lane's
// 8 values, fetched ONCE per phase by the caller so that no global load sits on the per-row dependency chain.
__device__ __forceinline__ void warp_layer_norm(float * row, const float (&add)[8], const float (&w)[8], float eps) {
    const int lane = threadIdx.x & 31;
    float v[8], s = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; k++) { v[k] = row[lane + 32 * k] + add[k]; s += v[k]; }
    const float mean = warp_sum(s) * (1.0f / 256.0f);
    float s2 = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; k++) { v[k] -= mean; s2 += v[k] * v[k]; }
    const float scale = 1.0f / sqrtf(warp_sum(s2) * (1.0f / 256.0f) + eps);
#pragma unroll
    for (int k = 0; k < 8; k++) row[lane + 32 * k] = (v[k] * scale) * w[k];
    __syncwarp();
}

__global__ void __launch_bounds__(kLtThreads, 1) lt_batch_kernel(const BParams bp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const LtParams & p = bp.p;
    cg::grid_group grid = cg::this_grid();
    const int G = gridDim.x, c = blockIdx.x, tid = threadIdx.x;
    const int B = p.B, d = p.d, L = p.L, F = p.F, V = p.V;
    float * tile = reinterpret_cast<float *>(smem_raw);
    SampSmem & S = *reinterpret_cast<SampSmem *>(smem_raw);
    float * wbase = reinterpret_cast<float *>(smem_raw + (size_t)kTileFloats * 4);

    // ---- this CTA's row slices of every matrix, resident for the whole launch ----
    const Slice s_in = slice_of(L, c, G), s_qkv = slice_of(3 * L, c, G), s_f1 = slice_of(F, c, G),
                s_f2 = slice_of(L, c, G), s_out = slice_of(V, c, G);
    float * w_in = wbase;
    float * w_qkv = w_in + (size_t)(s_in.n1 - s_in.n0) * d;
    float * w_f1 = w_qkv + (size_t)(s_qkv.n1 - s_qkv.n0) * L;
    float * w_f2 = w_f1 + (size_t)(s_f1.n1 - s_f1.n0) * L;
    float * w_out0 = w_f2 + (size_t)(s_f2.n1 - s_f2.n0) * F;
    const size_t out_elems = (size_t)(s_out.n1 - s_out.n0) * L;
    load_rows(w_in, p.in_w, d, s_in);
    // q and k rows as they are; value rows carry the folded output projection vo = (Wo Wv) n, stored as a bf16 hi + lo pair
    // (model.cu): rows [2L, 3L) = hi, [3L, 4L) = lo, summed here into one f32 row
    for (int r = s_qkv.n0; r < s_qkv.n1; r++) {
        const bf * src = reinterpret_cast<const bf *>(bp.qkvo) + (size_t)r * L;
        float * dst = w_qkv + (size_t)(r - s_qkv.n0) * L;
        for (int k = tid; k < L; k += kLtThreads)
            dst[k] = r < 2 * L ? __bfloat162float(src[k]) : __bfloat162float(src[k]) + __bfloat162float(src[(size_t)L * L + k]);
    }
    load_rows(w_f1, p.ff1_w, L, s_f1);
    load_rows(w_f2, p.ff2_w, F, s_f2);
    for (int cb = 0; cb < 8; cb++) load_rows(w_out0 + cb * out_elems, p.out_w[cb], L, s_out);

    int n_stamp = 0;
    auto stamp = [&]() {
        if (bp.dbg && c == 0 && tid == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); bp.dbg[n_stamp] = t; }
        n_stamp++;
    };
    stamp();
    const bool loop = p.d_step != nullptr || p.utt_step != nullptr;
    const int step0 = loop && p.d_step ? *p.d_step : (int)p.step;
    auto step_of = [&](size_t u) { return p.utt_step ? p.utt_step[u] : step0; };
    const float att_scale = 1.0f / sqrtf((float)L);
    auto no_prep = [](int, float *) {};

    // seq[0] = in_proj . hidden + b   (magpie.cpp:1153-1185)
    gemv_phase<false>(w_in, s_in, d, p.hidden, d, B, tile, no_prep, [&](int u, int n, float v) { bp.seq[(size_t)u * L + n] = v + p.in_b[n]; });
    if (p.hidden_hist)
        for (size_t i = (size_t)c * kLtThreads + tid; i < (size_t)B * d; i += (size_t)G * kLtThreads) {
            const size_t u = i / d, k = i % d;
            p.hidden_hist[((loop ? u * p.T_total + step_of(u) : u)) * d + k] = p.hidden[i];
        }
    stamp(); grid.sync(); stamp();

    bool hit_eos[8];                       // per owned utterance (utterances c, c + G, ...), thread-uniform
#pragma unroll
    for (int i = 0; i < 8; i++) hit_eos[i] = false;

    // attention of LT position `cbq` of utterance u by its owner CTA (magpie.cpp:946-1013): q from bp.q, keys / folded values of
    // positions 0..cbq from bp.kc / bp.vc; with the output projection folded into the value rows,
    // x1 = (seq + pos[cbq]) + sum_j p_j vo_j comes out directly.  `sc` = 8 floats of shared scratch.
    auto attend = [&](int u, int cbq, float * sc) {
        const int warp = tid >> 5, lane = tid & 31;
        const float * posq = p.pos + cbq * L;
        if (warp <= cbq) {
            float s = 0.0f;
            for (int i = lane; i < L; i += 32) s = fmaf(bp.kc[((size_t)u * 8 + warp) * L + i], bp.q[(size_t)u * L + i], s);
            s = warp_sum(s);
            if (lane == 0) sc[warp] = s * att_scale;
        }
        __syncthreads();
        if (tid < L) {
            float mxs = sc[0];
            for (int j = 1; j <= cbq; j++) mxs = fmaxf(mxs, sc[j]);
            float e[8], sum = 0.0f;
#pragma unroll
            for (int j = 0; j < 8; j++) { e[j] = (j <= cbq) ? expf(sc[j] - mxs) : 0.0f; sum += e[j]; }
            const float inv = 1.0f / sum;
            float o = 0.0f;
#pragma unroll
            for (int j = 0; j < 8; j++) if (j <= cbq) o = fmaf(e[j] * inv, bp.vc[((size_t)u * 8 + j) * L + tid], o);
            bp.x1[(size_t)u * L + tid] = (bp.seq[(size_t)u * L + tid] + posq[tid]) + o;
        }
        __syncthreads();
    };

    for (int cb = 0; cb < 8; cb++) {
        if (cb == 0) {
            // position 0: q | k | vo = [Wq; Wk; Wo Wv] . LN(seq + pos[0])   (magpie.cpp:1026-1030, 1501-1503).  Positions 1..7 take
            // their rows from the (codebook, fed code) table in the sampling phase below: no QKV phase, no attention phase.
            const float * pos = p.pos;
            float ln_add[8], ln_w[8];
#pragma unroll
            for (int k = 0; k < 8; k++) { ln_add[k] = pos[(tid & 31) + 32 * k]; ln_w[k] = p.norm_self[(tid & 31) + 32 * k]; }
            gemv_phase<true>(w_qkv, s_qkv, L, bp.seq, L, B, tile,
                       [&](int, float * row) { warp_layer_norm(row, ln_add, ln_w, p.eps); },
                       [&](int u, int n, float v) {
                           if (n < L) bp.q[(size_t)u * L + n] = v;
                           else if (n < 2 * L) bp.kc[((size_t)u * 8) * L + (n - L)] = v;
                           else bp.vc[((size_t)u * 8) * L + (n - 2 * L)] = v;
                       });
            stamp(); grid.sync(); stamp();
            for (int u = c; u < B; u += G) { __syncthreads(); attend(u, 0, tile); }
            stamp(); grid.sync(); stamp();
        }
        float ln_add[8], ln_w[8];
        // ffh = gelu(ff1 . LN(x1))
#pragma unroll
        for (int k = 0; k < 8; k++) { ln_add[k] = 0.0f; ln_w[k] = p.norm_ff[(tid & 31) + 32 * k]; }
        gemv_phase<true>(w_f1, s_f1, L, bp.x1, L, B, tile,
                   [&](int, float * row) { warp_layer_norm(row, ln_add, ln_w, p.eps); },
                   [&](int u, int n, float v) { bp.ffh[(size_t)u * F + n] = gelu_ggml(v, p.gelu_f16); });
        stamp(); grid.sync(); stamp();
        // hout = x1 + ff2 . ffh
        gemv_phase<false>(w_f2, s_f2, F, bp.ffh, F, B, tile, no_prep,
                   [&](int u, int n, float v) { bp.hout[(size_t)u * L + n] = v + bp.x1[(size_t)u * L + n]; });
        stamp(); grid.sync(); stamp();
        // logits = out_proj[cb] . hout + b   (magpie.cpp:1037-1048)
        {
            const float * ob = p.out_b[cb];
            gemv_phase<false>(w_out0 + cb * out_elems, s_out, L, bp.hout, L, B, tile, no_prep,
                       [&](int u, int n, float v) { bp.logits[(size_t)u * V + n] = v + ob[n]; });
        }
        stamp(); grid.sync(); stamp();
        // masking, argmax, top-k sampling, feedback embedding: owner CTA per utterance
        int oi = 0;
        for (int u = c; u < B; u += G, oi++) {
            __syncthreads();
            const int step = step_of(u);
            const size_t row = loop ? (size_t)u * p.T_total + step : (size_t)u;
            const int32_t * forced = p.forced ? p.forced + row * 8 : nullptr;
            const bool forbid_eos = p.forbid_eos_all || (p.forbid_eos && p.forbid_eos[u]) || (loop && step < p.min_frames);
            for (int i = tid; i < V; i += kLtThreads) S.logits[i] = bp.logits[(size_t)u * V + i];
            __syncthreads();
            // forbidden ids: BOS, BOS+2..BOS+7, and EOS while forbid_eos (magpie.cpp:1131-1145, 1243-1248)
            if (tid < 8) {
                const int id = tid == 0 ? p.bos_id : (tid < 7 ? p.bos_id + 1 + tid : (forbid_eos ? p.eos_id : -1));
                if (id >= 0 && id < V) S.logits[id] = -INFINITY;
            }
            __syncthreads();
            if (p.logits)
                for (int i = tid; i < V; i += kLtThreads) p.logits[(row * 8 + cb) * V + i] = S.logits[i];
            const int am = block_argmax(S.logits, V, S.red, S.redi);
            int pick = am;
            if (p.temperature >= 0.01f) {
                float uu;
                if (p.uniforms) uu = p.uniforms[row * 8 + cb];
                else {
                    uint32_t r[4];
                    philox4x32_10((uint32_t)step, (uint32_t)u, (uint32_t)cb, 0u, (uint32_t)p.seed, (uint32_t)(p.seed >> 32), r);
                    uu = (float)(r[0] >> 8) * (1.0f / 16777216.0f);
                }
                pick = block_sample_top_k(S, V, p.temperature, p.top_k, uu);
            }
            if (oi < 8) hit_eos[oi] = hit_eos[oi] || pick == p.eos_id || am == p.eos_id;
            if (tid == 0) {
                p.argmax[row * 8 + cb] = am;
                p.sampled[row * 8 + cb] = pick;
                if (p.next_codes) p.next_codes[u * 8 + cb] = forced ? forced[cb] : pick;
                if (cb == 7) {
                    const bool he = oi < 8 ? hit_eos[oi] : false;
                    if (p.eos_flag) p.eos_flag[u] = he ? 1 : 0;
                    if (p.done_step && he && p.done_step[u] < 0) p.done_step[u] = step;
                }
            }
            if (cb < 7) {
                // seq[cb+1] = in_proj . E_cb[code] + b and its [q | k | vo] row, both tabulated at load (model.cu); the owner CTA
                // then runs the attention of position cb+1 right here
                const int fed = forced ? forced[cb] : pick;
                if (tid < L) {
                    const float * row = bp.qkv_tab + ((size_t)cb * V + fed) * (3 * L);
                    bp.seq[(size_t)u * L + tid] = p.in_table[cb][(size_t)fed * L + tid];
                    bp.q[(size_t)u * L + tid] = row[tid];
                    bp.kc[((size_t)u * 8 + cb + 1) * L + tid] = row[L + tid];
                    bp.vc[((size_t)u * 8 + cb + 1) * L + tid] = row[2 * L + tid];
                }
                __syncthreads();
                attend(u, cb + 1, reinterpret_cast<float *>(S.hist));      // (the sampler's histogram is free again)
            }
        }
        if (cb < 7) { stamp(); grid.sync(); stamp(); }
    }
}

size_t slice_smem_bytes(const Model & m, int G) {
    const mgb_hparams & hp = m.hp;
    auto rows = [&](int N) { return (size_t)((N + G - 1) / G); };         // largest balanced slice
    const int L = hp.lt_dim, F = hp.lt_ffn_dim, d = hp.d_model, V = hp.vocab_per_cb;
    size_t e = rows(L) * d + rows(3 * L) * L + rows(F) * L + rows(L) * F + 8 * rows(V) * L;
    return e * sizeof(float);
}

}  // namespace

size_t lt_batch_scratch_bytes(const Model & m, int B) {
    const mgb_hparams & hp = m.hp;
    const size_t L = hp.lt_dim, F = hp.lt_ffn_dim, V = hp.vocab_per_cb;
    return (size_t)B * (L * 5 + 2 * 8 * L + F + V) * sizeof(float);
}

// bf16 models with the folded feedback table, >= 16 utterances, at most 8 utterances per owner CTA
bool lt_batch_supported(const Model & m, int B) {
    if (getenv("MGB_NO_LT_BATCH") != nullptr) return false;
    const mgb_hparams & hp = m.hp;
    if (m.precision != MGB_PREC_BF16 || !m.lt_in_table[0] || !m.lt_qkvo || !m.lt_qkv_tab || B < 16) return false;
    if (hp.lt_dim != 256 || hp.lt_ffn_dim > kF || hp.lt_ffn_dim % 256 != 0 || hp.d_model % 256 != 0 || hp.d_model > kD || hp.vocab_per_cb > kV) return false;
    return B <= 8 * 132;
}

bool launch_lt_batch(const Model & m, const LtParams & p, void * scratch, size_t scratch_bytes, cudaStream_t stream) {
    static std::atomic<int> n_sm[64];
    static DeviceOnce attr_done;
    int dev = 0;
    MGB_CUDA_TRY(cudaGetDevice(&dev));
    if (!n_sm[dev & 63]) { int n = 0; MGB_CUDA_TRY(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev)); n_sm[dev & 63] = n; }
    const int G = n_sm[dev & 63];
    const size_t smem = (size_t)kTileFloats * 4 + slice_smem_bytes(m, G);
    if (smem > 227 * 1024) { set_error("lt_batch: weight slices do not fit shared memory"); return false; }
    if (scratch_bytes < lt_batch_scratch_bytes(m, p.B) || !scratch) { set_error("lt_batch: scratch too small"); return false; }
    if (!attr_done.done(dev)) {
        MGB_CUDA_TRY(cudaFuncSetAttribute(lt_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_done.set(dev);
    }
    BParams bp;
    bp.p = p;
    bp.qkvo = m.lt_qkvo; bp.qkv_tab = m.lt_qkv_tab;
    const size_t B = p.B, L = p.L, F = p.F, V = p.V;
    float * f = (float *)scratch;
    bp.seq = f; f += B * L; bp.q = f; f += B * L; bp.kc = f; f += B * 8 * L; bp.vc = f; f += B * 8 * L;
    bp.att = f; f += B * L; bp.x1 = f; f += B * L; bp.ffh = f; f += B * F; bp.hout = f; f += B * L; bp.logits = f; f += B * V;
    static unsigned long long * dbg = nullptr;
    if (getenv("MGB_LT_DBG") && !dbg) { MGB_CUDA_TRY(cudaMalloc((void **)&dbg, 256 * 8)); MGB_CUDA_TRY(cudaMemset(dbg, 0, 256 * 8)); }
    bp.dbg = dbg;
    if (dbg && getenv("MGB_LT_DBG_DUMP")) {
        unsigned long long h[256];
        cudaDeviceSynchronize();
        cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost);
        for (int i = 1; i < 256 && h[i]; i++) fprintf(stderr, "lt_batch stamp %3d  +%6llu ns  (%s)\n", i, h[i] - h[i - 1], (i & 1) ? "work" : "sync");
    }
    void * args[] = {&bp};
    MGB_CUDA_TRY(cudaLaunchCooperativeKernel((void *)lt_batch_kernel, dim3(G), dim3(kLtThreads), args, smem, stream));
    MGB_LAUNCH_CHECK();
    return true;
}

}  // namespace mgb
