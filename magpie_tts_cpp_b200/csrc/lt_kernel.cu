// Local transformer + sampler: all 8 codebooks of a frame in ONE persistent kernel.
//
// Restates magpie_local_transformer_sample_all (reference src/magpie.cpp:1113-1317), the LT layer
// (magpie.cpp:946-1013), the per-codebook output projection (magpie.cpp:1037-1048) and
// sample_top_k (magpie.cpp:1072-1109).  The reference re-runs the whole prefix for every codebook
// (36 token passes, 16 graphs, 8 D2H syncs per frame); the layer is causal and single, so a K/V
// cache of the <= 8 positions kept in shared memory gives the identical result with 8 passes.
//
// Parallelisation: one thread-block CLUSTER per utterance.  Every GEMV is row-sliced over the CTAs
// of the cluster; each CTA broadcasts its output slice into every CTA's shared memory through
// DSMEM, followed by one cluster barrier.  LayerNorm, the <= 8-key attention, masking, argmax and
// the top-k sampler are cheap and run redundantly (bit-identically) in every CTA, so no gather is
// needed.  Weights stream from L2/HBM with 16-byte loads, 4 rows in flight per warp.
#include <cooperative_groups.h>

#include "common.cuh"
#include "kernels.cuh"

namespace cg = cooperative_groups;

namespace mgb {

namespace {

constexpr int kLtThreads = 512;
constexpr int kLtWarps = kLtThreads / 32;
constexpr int kL = 256;        // lt_dim (max)
constexpr int kF = 1024;       // lt_ffn_dim (max)
constexpr int kD = 1024;       // d_model (max)
constexpr int kV = 2048;       // vocab_per_cb (max)

struct LtParams {
    int B, d, L, F, V;
    float eps; int gelu_f16;
    const float * hidden;
    const void * in_w; const float * in_b;
    const float * pos;
    const float * norm_self; const float * norm_ff;
    const void * qkv_w; const void * o_w; const void * ff1_w; const void * ff2_w;
    const void * out_w[8]; const float * out_b[8];
    const float * audio_emb[8];
    float temperature; int top_k;
    const uint8_t * forbid_eos; int forbid_eos_all;
    const int32_t * forced; const float * uniforms;
    uint64_t seed; uint32_t step;
    int bos_id, eos_id;
    int32_t * sampled; int32_t * argmax; int32_t * next_codes; float * logits; int32_t * eos_flag;
    const int32_t * d_step; int T_total, min_frames; int32_t * done_step; float * hidden_hist;
};

struct LtSmem {
    float hid[kD];
    float seq[kL];          // s[cb] (in-projection output), broadcast
    float x[kL];            // s[cb] + pos[cb]
    float nrm[kL];
    float q[kL];
    float kc[8][kL];
    float vc[8][kL];
    float att[kL];
    float x1[kL];
    float ffh[kF];
    float hout[kL];
    float logits[kV];
    float sel_v[kV]; int sel_i[kV];
    float srt_v[kV]; int srt_i[kV];
    float red[32]; int redi[32];
    unsigned hist[256];
    float scores[8];
    int   misc[8];
};

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t (&out)[4]) {
#pragma unroll
    for (int i = 0; i < 10; i++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ unsigned order_key(float f) {      // larger float -> larger key
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

template <int CS>
__device__ __forceinline__ void cluster_barrier(cg::cluster_group & cluster) {
    if constexpr (CS > 1) cluster.sync(); else __syncthreads();
}

// dst[n] (in every CTA of the cluster) = epi(n, sum_k W[n][k] x[k]) for the rows of this CTA's slice.
template <typename T, int CS, typename Epi>
__device__ __forceinline__ void gemv_bcast(cg::cluster_group & cluster, const T * __restrict__ W, int N, int K,
                                           const float * x, float * dst, int rank, Epi epi) {
    constexpr int VEC = WT<T>::VEC;
    constexpr int RW = (CS >= 8) ? 32 / CS : 4;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gw = rank * kLtWarps + warp;
    for (int n0 = gw * RW; n0 < N; n0 += CS * kLtWarps * RW) {
        float acc[RW];
#pragma unroll
        for (int r = 0; r < RW; r++) acc[r] = 0.0f;
        for (int k = lane * VEC; k < K; k += 32 * VEC) {
            float w[RW][VEC];
#pragma unroll
            for (int r = 0; r < RW; r++) {
                if (n0 + r < N) WT<T>::load(W + (size_t)(n0 + r) * K + k, w[r]);
                else {
#pragma unroll
                    for (int v = 0; v < VEC; v++) w[r][v] = 0.0f;
                }
            }
            float xv[VEC];
#pragma unroll
            for (int v4 = 0; v4 < VEC; v4 += 4) {
                float4 t4 = *reinterpret_cast<const float4 *>(x + k + v4);
                xv[v4] = t4.x; xv[v4 + 1] = t4.y; xv[v4 + 2] = t4.z; xv[v4 + 3] = t4.w;
            }
#pragma unroll
            for (int r = 0; r < RW; r++)
#pragma unroll
                for (int v = 0; v < VEC; v++) acc[r] = fmaf(w[r][v], xv[v], acc[r]);
        }
        float mine = 0.0f;
#pragma unroll
        for (int r = 0; r < RW; r++) {
            float v = warp_sum(acc[r]);
            if (lane / CS == r) mine = v;
        }
        const int r = lane / CS, dr = lane % CS;
        if (r < RW && n0 + r < N) {
            const float v = epi(n0 + r, mine);
            if constexpr (CS > 1) cluster.map_shared_rank(dst, dr)[n0 + r] = v;
            else dst[n0 + r] = v;
        }
    }
}

__device__ __forceinline__ void block_layer_norm(const float * x, const float * w, float * y, int n, float eps, float * red) {
    const int tid = threadIdx.x;
    float v = tid < n ? x[tid] : 0.0f;
    const float mean = block_sum(v, red) / (float)n;
    const float c = tid < n ? v - mean : 0.0f;
    const float var = block_sum(c * c, red) / (float)n;
    const float scale = 1.0f / sqrtf(var + eps);
    if (tid < n) y[tid] = (c * scale) * w[tid];
    __syncthreads();
}

// argmax with "first max wins" (strict >, lowest index on ties)  -- magpie.cpp:1250-1259
__device__ __forceinline__ int block_argmax(const float * v, int n, float * red, int * redi) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    float bv = -INFINITY; int bi = 0x7fffffff;
    for (int i = tid; i < n; i += kLtThreads) { float f = v[i]; if (f > bv || (f == bv && i < bi)) { bv = f; bi = i; } }
    // NB: all -inf rows never happen (only 8 ids are masked)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, bv, o); int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    __syncthreads();
    if (lane == 0) { red[wid] = bv; redi[wid] = bi; }
    __syncthreads();
    if (wid == 0) {
        bv = lane < kLtWarps ? red[lane] : -INFINITY; bi = lane < kLtWarps ? redi[lane] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float ov = __shfl_xor_sync(0xffffffffu, bv, o); int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) redi[0] = bi;
    }
    __syncthreads();
    const int r = redi[0];
    __syncthreads();
    return r;
}

// sample_top_k (magpie.cpp:1072-1109): k largest (value desc, index asc), softmax((l - max)/T) with
// sequential float accumulation, inverse CDF with draw u; fallback = last of the k.
__device__ int block_sample_top_k(LtSmem & S, int V, float temperature, int top_k, float u) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    int k = top_k < V ? top_k : V;
    if (k < 1) k = 1;                              // top_k = 0 is UB in the reference: clamp
    // ---- radix select: key of the k-th largest element ----
    unsigned prefix = 0, pmask = 0; int want = k;
    for (int pass = 0; pass < 4; pass++) {
        const int shift = 24 - 8 * pass;
        if (tid < 256) S.hist[tid] = 0;
        __syncthreads();
        for (int i = tid; i < V; i += kLtThreads) {
            const unsigned key = order_key(S.logits[i]);
            if ((key & pmask) == prefix) atomicAdd(&S.hist[(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            int acc = 0, b = 255;
            for (; b > 0; b--) { if (acc + (int)S.hist[b] >= want) break; acc += (int)S.hist[b]; }
            S.misc[0] = b; S.misc[1] = want - acc;
        }
        __syncthreads();
        prefix |= (unsigned)S.misc[0] << shift; pmask |= 255u << shift; want = S.misc[1];
        __syncthreads();
    }
    const unsigned thr = prefix;                   // k-th largest key; `want` ties to take (lowest indices)
    // ---- compaction in index order by warp 0 ----
    if (wid == 0) {
        int cnt = 0, eq_taken = 0;
        for (int i0 = 0; i0 < V; i0 += 32) {
            const int i = i0 + lane;
            const unsigned key = i < V ? order_key(S.logits[i]) : 0u;
            const bool gt = i < V && key > thr, eq = i < V && key == thr;
            const unsigned eqm = __ballot_sync(0xffffffffu, eq);
            const int eq_rank = eq_taken + __popc(eqm & ((1u << lane) - 1u));
            const bool take = gt || (eq && eq_rank < want);
            const unsigned tm = __ballot_sync(0xffffffffu, take);
            if (take) { const int ppos = cnt + __popc(tm & ((1u << lane) - 1u)); S.sel_v[ppos] = S.logits[i]; S.sel_i[ppos] = i; }
            cnt += __popc(tm); eq_taken += __popc(eqm);
        }
    }
    __syncthreads();
    // ---- rank by counting -> sorted (value desc, index asc) ----
    for (int a = tid; a < k; a += kLtThreads) {
        const float va = S.sel_v[a]; const int ia = S.sel_i[a];
        int r = 0;
        for (int b = 0; b < k; b++) { const float vb = S.sel_v[b]; r += (vb > va || (vb == va && S.sel_i[b] < ia)) ? 1 : 0; }
        S.srt_v[r] = va; S.srt_i[r] = ia;
    }
    __syncthreads();
    const float mx = S.srt_v[0];
    for (int a = tid; a < k; a += kLtThreads) S.sel_v[a] = expf((S.srt_v[a] - mx) / temperature);
    __syncthreads();
    if (tid == 0) {
        float sum = 0.0f;
        for (int a = 0; a < k; a++) sum += S.sel_v[a];
        float cum = 0.0f; int pick = S.srt_i[k - 1];
        for (int a = 0; a < k; a++) { cum += S.sel_v[a] / sum; if (u < cum) { pick = S.srt_i[a]; break; } }
        S.misc[2] = pick;
    }
    __syncthreads();
    const int r = S.misc[2];
    __syncthreads();
    return r;
}

template <typename T, int CS>
__global__ void __launch_bounds__(kLtThreads, 1) lt_kernel(const LtParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LtSmem & S = *reinterpret_cast<LtSmem *>(smem_raw);
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = CS > 1 ? (int)cluster.block_rank() : 0;
    const int utt = blockIdx.x / CS;
    const int tid = threadIdx.x;
    const int d = p.d, L = p.L, F = p.F, V = p.V;
    const float att_scale = 1.0f / sqrtf((float)L);

    // loop mode: per-step arrays are indexed by the device-side step counter (CUDA-graph replayable)
    const bool loop = p.d_step != nullptr;
    const int step = loop ? *p.d_step : (int)p.step;
    const size_t row = loop ? (size_t)utt * p.T_total + step : (size_t)utt;
    const int32_t * forced = p.forced ? p.forced + row * 8 : nullptr;
    const float * uniforms = p.uniforms ? p.uniforms + row * 8 : nullptr;
    float * logits_out = p.logits ? p.logits + row * 8 * V : nullptr;
    int32_t * sampled_out = p.sampled + row * 8;
    int32_t * argmax_out = p.argmax + row * 8;

    for (int i = tid; i < d; i += kLtThreads) {
        const float h = p.hidden[(size_t)utt * d + i];
        S.hid[i] = h;
        if (p.hidden_hist && rank == 0) p.hidden_hist[row * d + i] = h;
    }
    __syncthreads();
    // seq[0] = in_proj . hidden + b   (magpie.cpp:1153-1185)
    gemv_bcast<T, CS>(cluster, (const T *)p.in_w, L, d, S.hid, S.seq, rank,
                      [&](int n, float v) { return v + p.in_b[n]; });
    cluster_barrier<CS>(cluster);

    const bool forbid_eos = p.forbid_eos_all || (p.forbid_eos && p.forbid_eos[utt]) || (loop && step < p.min_frames);
    bool hit_eos = false;

    for (int cb = 0; cb < 8; cb++) {
        if (tid < L) S.x[tid] = S.seq[tid] + p.pos[cb * L + tid];        // magpie.cpp:1026-1030
        __syncthreads();
        block_layer_norm(S.x, p.norm_self, S.nrm, L, p.eps, S.red);
        // q | k | v rows of qkv_net (magpie.cpp:1501-1503)
        {
            float * kdst = S.kc[cb], * vdst = S.vc[cb];
            constexpr int RWQ = (CS >= 8) ? 32 / CS : 4;
            (void)RWQ;
            // three slices so that each lands in its own buffer
            gemv_bcast<T, CS>(cluster, (const T *)p.qkv_w, L, L, S.nrm, S.q, rank, [](int, float v) { return v; });
            gemv_bcast<T, CS>(cluster, (const T *)p.qkv_w + (size_t)L * L, L, L, S.nrm, kdst, rank, [](int, float v) { return v; });
            gemv_bcast<T, CS>(cluster, (const T *)p.qkv_w + (size_t)2 * L * L, L, L, S.nrm, vdst, rank, [](int, float v) { return v; });
        }
        cluster_barrier<CS>(cluster);
        // single-head causal attention over positions 0..cb (magpie.cpp:946-1013)
        {
            const int warp = tid >> 5, lane = tid & 31;
            if (warp <= cb) {
                float s = 0.0f;
                for (int i = lane; i < L; i += 32) s = fmaf(S.kc[warp][i], S.q[i], s);
                s = warp_sum(s);
                if (lane == 0) S.scores[warp] = s * att_scale;
            }
            __syncthreads();
            if (tid < L) {
                float mxs = S.scores[0];
                for (int j = 1; j <= cb; j++) mxs = fmaxf(mxs, S.scores[j]);
                float e[8], sum = 0.0f;
#pragma unroll
                for (int j = 0; j < 8; j++) { e[j] = (j <= cb) ? expf(S.scores[j] - mxs) : 0.0f; sum += e[j]; }
                const float inv = 1.0f / sum;
                float o = 0.0f;
#pragma unroll
                for (int j = 0; j < 8; j++) if (j <= cb) o = fmaf(e[j] * inv, S.vc[j][tid], o);
                S.att[tid] = o;
            }
            __syncthreads();
        }
        gemv_bcast<T, CS>(cluster, (const T *)p.o_w, L, L, S.att, S.x1, rank, [&](int n, float v) { return v + S.x[n]; });
        cluster_barrier<CS>(cluster);
        block_layer_norm(S.x1, p.norm_ff, S.nrm, L, p.eps, S.red);
        gemv_bcast<T, CS>(cluster, (const T *)p.ff1_w, F, L, S.nrm, S.ffh, rank,
                          [&](int, float v) { return gelu_ggml(v, p.gelu_f16); });
        cluster_barrier<CS>(cluster);
        gemv_bcast<T, CS>(cluster, (const T *)p.ff2_w, L, F, S.ffh, S.hout, rank, [&](int n, float v) { return v + S.x1[n]; });
        cluster_barrier<CS>(cluster);
        const float * ob = p.out_b[cb];
        gemv_bcast<T, CS>(cluster, (const T *)p.out_w[cb], V, L, S.hout, S.logits, rank,
                          [&](int n, float v) { return v + ob[n]; });
        cluster_barrier<CS>(cluster);
        // forbidden ids: BOS, BOS+2..BOS+7, and EOS while forbid_eos (magpie.cpp:1131-1145, 1243-1248)
        if (tid < 8) {
            int id = tid == 0 ? p.bos_id : (tid < 7 ? p.bos_id + 1 + tid : (forbid_eos ? p.eos_id : -1));
            if (id >= 0 && id < V) S.logits[id] = -INFINITY;
        }
        __syncthreads();
        if (logits_out && rank == 0)
            for (int i = tid; i < V; i += kLtThreads) logits_out[(size_t)cb * V + i] = S.logits[i];
        const int am = block_argmax(S.logits, V, S.red, S.redi);
        int pick = am;
        if (p.temperature >= 0.01f) {
            float u;
            if (uniforms) u = uniforms[cb];
            else {
                uint32_t r[4];
                philox4x32_10((uint32_t)step, (uint32_t)utt, (uint32_t)cb, 0u, (uint32_t)p.seed, (uint32_t)(p.seed >> 32), r);
                u = (float)(r[0] >> 8) * (1.0f / 16777216.0f);
            }
            pick = block_sample_top_k(S, V, p.temperature, p.top_k, u);
        }
        hit_eos = hit_eos || pick == p.eos_id || am == p.eos_id;
        if (rank == 0 && tid == 0) {
            argmax_out[cb] = am;
            sampled_out[cb] = pick;
            if (p.next_codes) p.next_codes[utt * 8 + cb] = forced ? forced[cb] : pick;
            if (cb == 7) {
                if (p.eos_flag) p.eos_flag[utt] = hit_eos ? 1 : 0;
                if (p.done_step && hit_eos && p.done_step[utt] < 0) p.done_step[utt] = step;
            }
        }
        if (cb < 7) {
            // seq[cb+1] = in_proj . E_cb[code] + b, embedding NOT scaled by 1/8 (magpie.cpp:1274-1313)
            const int fed = forced ? forced[cb] : pick;
            const float * er = p.audio_emb[cb] + (size_t)fed * d;
            for (int i = tid; i < d; i += kLtThreads) S.hid[i] = er[i];
            __syncthreads();
            gemv_bcast<T, CS>(cluster, (const T *)p.in_w, L, d, S.hid, S.seq, rank,
                              [&](int n, float v) { return v + p.in_b[n]; });
            cluster_barrier<CS>(cluster);
        }
    }
    // p.next_codes = what the next decoder step consumes (the forced codes under teacher forcing)
    cluster_barrier<CS>(cluster);      // no CTA may exit while peers can still write into its smem
}

template <typename T, int CS>
bool launch_lt_t(const LtParams & p, cudaStream_t stream) {
    const size_t smem = sizeof(LtSmem);
    static uint64_t attr_done = 0;
    int dev = 0;
    MGB_CUDA_TRY(cudaGetDevice(&dev));
    if (!(attr_done >> dev & 1)) {
        MGB_CUDA_TRY(cudaFuncSetAttribute(lt_kernel<T, CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (CS > 8) MGB_CUDA_TRY(cudaFuncSetAttribute(lt_kernel<T, CS>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        attr_done |= 1ull << dev;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(p.B * CS); cfg.blockDim = dim3(kLtThreads); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    MGB_CUDA_TRY(cudaLaunchKernelEx(&cfg, lt_kernel<T, CS>, p));
    MGB_LAUNCH_CHECK();
    return true;
}

}  // namespace

bool launch_local_transformer(const Model & m, const LtArgs & a, cudaStream_t stream) {
    if (a.B <= 0) return true;
    LtParams p;
    p.B = a.B; p.d = m.hp.d_model; p.L = m.hp.lt_dim; p.F = m.hp.lt_ffn_dim; p.V = m.hp.vocab_per_cb;
    p.eps = m.hp.eps; p.gelu_f16 = m.gelu_f16;
    p.hidden = a.hidden; p.in_w = m.lt_in_w.w; p.in_b = m.lt_in_b; p.pos = m.lt_pos;
    p.norm_self = m.lt_norm_self; p.norm_ff = m.lt_norm_ff;
    p.qkv_w = m.lt_qkv.w; p.o_w = m.lt_o.w; p.ff1_w = m.lt_ff1.w; p.ff2_w = m.lt_ff2.w;
    for (int cb = 0; cb < 8; cb++) { p.out_w[cb] = m.lt_out_w[cb].w; p.out_b[cb] = m.lt_out_b[cb]; p.audio_emb[cb] = m.audio_emb[cb]; }
    p.temperature = a.temperature; p.top_k = a.top_k; p.forbid_eos = a.forbid_eos; p.forbid_eos_all = a.forbid_eos_all;
    p.forced = a.forced; p.uniforms = a.uniforms; p.seed = a.seed; p.step = a.step;
    p.bos_id = m.hp.audio_bos_id; p.eos_id = m.hp.audio_eos_id;
    p.sampled = a.sampled; p.argmax = a.argmax; p.next_codes = a.next_codes; p.logits = a.logits; p.eos_flag = a.eos_flag;
    p.d_step = a.d_step; p.T_total = a.T_total; p.min_frames = a.min_frames; p.done_step = a.done_step; p.hidden_hist = a.hidden_hist;
    if (m.precision == MGB_PREC_F32) return launch_lt_t<float, 8>(p, stream);
    return launch_lt_t<__nv_bfloat16, 8>(p, stream);
}

}  // namespace mgb
