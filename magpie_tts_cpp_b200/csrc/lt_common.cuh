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
// Round 2: the first version took ~40 us per call (877 vs 556 us per batched step with sampling on): its histogram atomics all hit the
// two or three bins the logits' exponent byte falls into, one thread walked the 256 bins of every radix pass, one warp compacted the
// 2024 candidates chunk by chunk.  Now: warp-aggregated atomics (one per distinct bin and warp), the bin of a pass found by a parallel
// suffix sum, compaction by all warps around a prefix over the 64 index chunks; the arithmetic that decides the result (order, sum,
// CDF) is unchanged.
template <typename SM>
__device__ int block_sample_top_k(SM & S, int V, float temperature, int top_k, float u) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    int k = top_k < V ? top_k : V;
    if (k < 1) k = 1;                              // top_k = 0 is UB in the reference: clamp
    // ---- radix select: key of the k-th largest element ----
    unsigned prefix = 0, pmask = 0; int want = k;
    if (tid < 256) S.hist[tid] = 0;
    __syncthreads();
    for (int pass = 0; pass < 4; pass++) {
        const int shift = 24 - 8 * pass;
        for (int i0 = 0; i0 < V; i0 += kLtThreads) {
            const int i = i0 + tid;
            const unsigned key = i < V ? order_key(S.logits[i]) : 0u;
            const bool valid = i < V && (key & pmask) == prefix;
            const unsigned bin = (key >> shift) & 255u;
            const unsigned vm = __ballot_sync(0xffffffffu, valid);
            if (valid) {
                if (pass == 0) {                     // all V keys, two or three hot bins (sign + exponent byte): one atomic per bin and warp
                    const unsigned peers = __match_any_sync(vm, bin);
                    if (lane == __ffs(peers) - 1) atomicAdd(&S.hist[bin], (unsigned)__popc(peers));
                } else atomicAdd(&S.hist[bin], 1u);  // later passes: only the keys of one bin of the previous pass
            }
        }
        __syncthreads();
        // the bin b >= 1 with the most keys at or above it still >= want (else 0): suffix sums over the 256 bins, one bin per thread
        unsigned c = 0, sfx = 0;
        if (tid < 256) {
            c = S.hist[tid]; sfx = c;
            S.hist[tid] = 0;                       // (this thread's bin, ready for the next pass: no separate clearing step)
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_down_sync(0xffffffffu, sfx, o); if (lane + o < 32) sfx += t; }
        }
        if (tid < 256 && lane == 0) reinterpret_cast<unsigned *>(S.sel_v)[wid] = sfx;      // the warp's 32 bins together (sel_v is free until the compaction)
        __syncthreads();
        if (tid < 256) {
            unsigned off = 0;
            for (int w = wid + 1; w < 8; w++) off += reinterpret_cast<const unsigned *>(S.sel_v)[w];
            const unsigned suf = sfx + off;                          // keys in bins >= tid
            if (suf - c < (unsigned)want && (tid == 0 || suf >= (unsigned)want)) { S.misc[0] = tid; S.misc[1] = want - (int)(suf - c); }
        }
        __syncthreads();
        prefix |= (unsigned)S.misc[0] << shift; pmask |= 255u << shift; want = S.misc[1];
        // (no barrier here: the next pass writes misc only after three more barriers, and hist was last read before the previous one)
    }
    __syncthreads();
    const unsigned thr = prefix;                   // k-th largest key; `want` ties to take (lowest indices)
    // ---- compaction in index order: per 32-index chunk counts, a prefix over the <= 64 chunks, then every warp places its chunks ----
    const int nchunk = (V + 31) >> 5;              // <= 64 (V <= 2048): hist = [greater | equal | base | equal before], 64 entries each
    for (int ch = wid; ch < nchunk; ch += kLtWarps) {
        const int i = ch * 32 + lane;
        const unsigned key = i < V ? order_key(S.logits[i]) : 0u;
        const unsigned gm = __ballot_sync(0xffffffffu, i < V && key > thr), em = __ballot_sync(0xffffffffu, i < V && key == thr);
        if (lane == 0) { S.hist[ch] = (unsigned)__popc(gm); S.hist[64 + ch] = (unsigned)__popc(em); }
    }
    __syncthreads();
    if (wid == 0) {
        int eqb[2], base[2], carry_e = 0, carry_t = 0;
#pragma unroll
        for (int h = 0; h < 2; h++) {              // chunks lane and lane + 32
            const int ch = lane + 32 * h;
            const int e = ch < nchunk ? (int)S.hist[64 + ch] : 0, g = ch < nchunk ? (int)S.hist[ch] : 0;
            int pe = e;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, pe, o); if (lane >= o) pe += t; }
            eqb[h] = carry_e + pe - e;             // equal keys in earlier chunks
            const int te = min(max(want - eqb[h], 0), e), tk = g + te;
            int pt = tk;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, pt, o); if (lane >= o) pt += t; }
            base[h] = carry_t + pt - tk;
            carry_e += __shfl_sync(0xffffffffu, pe, 31); carry_t += __shfl_sync(0xffffffffu, pt, 31);
            if (ch < nchunk) { S.hist[128 + ch] = (unsigned)base[h]; S.hist[192 + ch] = (unsigned)eqb[h]; }
        }
    }
    __syncthreads();
    for (int ch = wid; ch < nchunk; ch += kLtWarps) {
        const int i = ch * 32 + lane;
        const unsigned key = i < V ? order_key(S.logits[i]) : 0u;
        const bool gt = i < V && key > thr, eq = i < V && key == thr;
        const unsigned eqm = __ballot_sync(0xffffffffu, eq);
        const int eq_rank = (int)S.hist[192 + ch] + __popc(eqm & ((1u << lane) - 1u));
        const bool take = gt || (eq && eq_rank < want);
        const unsigned tm = __ballot_sync(0xffffffffu, take);
        if (take) { const int ppos = (int)S.hist[128 + ch] + __popc(tm & ((1u << lane) - 1u)); S.sel_v[ppos] = S.logits[i]; S.sel_i[ppos] = i; }
    }
    __syncthreads();
    // ---- rank by counting -> sorted (value desc, index asc) ----
    // (all threads: element a = tid mod 128, a quarter of the k comparisons each, partial ranks added in shared memory; the first
    //  version ran k threads x k dependent iterations: 9 000 cycles of the sampler's 28 000)
    if (tid < 128) S.hist[tid] = 0;
    __syncthreads();
    {
        const int a = tid & 127, part = tid >> 7, per = (k + 3) >> 2;          // kLtThreads = 512 = 4 parts x 128 elements (k <= 128 here)
        if (a < k && k <= 128) {
            const float va = S.sel_v[a]; const int ia = S.sel_i[a];
            int r = 0;
            const int b1 = min(k, (part + 1) * per);
#pragma unroll 4
            for (int b = part * per; b < b1; b++) { const float vb = S.sel_v[b]; r += (vb > va || (vb == va && S.sel_i[b] < ia)) ? 1 : 0; }
            if (r) atomicAdd(&S.hist[a], (unsigned)r);
        }
    }
    __syncthreads();
    if (k <= 128) {
        if (tid < k) { const int r = (int)S.hist[tid]; S.srt_v[r] = S.sel_v[tid]; S.srt_i[r] = S.sel_i[tid]; }
    } else {
        for (int a = tid; a < k; a += kLtThreads) {
            const float va = S.sel_v[a]; const int ia = S.sel_i[a];
            int r = 0;
            for (int b = 0; b < k; b++) { const float vb = S.sel_v[b]; r += (vb > va || (vb == va && S.sel_i[b] < ia)) ? 1 : 0; }
            S.srt_v[r] = va; S.srt_i[r] = ia;
        }
    }
    __syncthreads();
    const float mx = S.srt_v[0];
    for (int a = tid; a < k; a += kLtThreads) S.sel_v[a] = expf((S.srt_v[a] - mx) / temperature);
    __syncthreads();
    if (tid == 0) {                                // sequential float accumulation, as the reference
        float sum = 0.0f;
#pragma unroll 8
        for (int a = 0; a < k; a++) sum += S.sel_v[a];
        S.misc[3] = __float_as_int(sum);
    }
    __syncthreads();
    {
        const float sum = __int_as_float(S.misc[3]);
        for (int a = tid; a < k; a += kLtThreads) S.srt_v[a] = S.sel_v[a] / sum;       // (the sorted values are no longer needed)
    }
    __syncthreads();
    if (tid == 0) {
        float cum = 0.0f; int pa = k - 1; bool found = false;      // (no early exit: the loads of the unrolled loop go out ahead of the sums)
#pragma unroll 8
        for (int a = 0; a < k; a++) { cum += S.srt_v[a]; if (!found && u < cum) { pa = a; found = true; } }
        S.misc[2] = S.srt_i[pa];
    }
    __syncthreads();
    const int r = S.misc[2];
    __syncthreads();
    return r;
}


}  // namespace lt
}  // namespace mgb
