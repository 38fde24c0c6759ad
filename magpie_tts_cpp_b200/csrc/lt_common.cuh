// Shared device helpers of the local-transformer kernels (lt_kernel.cu: streaming weights; lt_resident.cu:
// shared-memory-resident weights).  See lt_kernel.cu for the reference citations.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"
#include "kernels.cuh"

namespace mgb {
namespace lt {
namespace cg = cooperative_groups;


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
    const int32_t * utt_step;       // optional [B]: per-utterance step counters instead of *d_step (continuous batching: utterances start at different times)
    int stream_feedback = 0;        // 1: feedback in-projection as a GEMV (lt_kernel); 0: row gather from in_table
    const float * in_table[8];      // P_cb = E_cb . Win^T + b, [V][L] f32 (resident kernel: feedback is a row gather)
};

}  // namespace lt
// lt_cluster.cu: utterances per 16-CTA cluster for a batch of B (0 = not supported) / launch
int  lt_cluster_plan(const Model & m, int B);
bool launch_lt_cluster(const Model & m, const lt::LtParams & p, int U, cudaStream_t stream);
namespace lt {

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
template <typename SM>
__device__ int block_sample_top_k(SM & S, int V, float temperature, int top_k, float u) {
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


}  // namespace lt
}  // namespace mgb
