// Launch wrappers of the sm_100a kernels (host-callable). All launches are asynchronous on
// `stream`; every wrapper returns false (with the thread error set) if the launch failed.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "model.h"

namespace mgb {

// ---- generic "tokens" view ---------------------------------------------------------------------
// A launch processes M tokens (rows). tok_utt[t] = utterance, tok_pos[t] = position inside that
// utterance's sequence, tok_slot[t] = row of the token in the per-layer KV storage.
struct Tokens {
    int M = 0;
    const int32_t * utt = nullptr;
    const int32_t * pos = nullptr;
    const int32_t * slot = nullptr;
};

enum { ACT_NONE = 0, ACT_GELU = 1 };
constexpr int kKvPageRows = 128;      // rows (cache positions) per page of the paged decoder self-attention K / V cache

struct LinearArgs {
    DevMat W;                       // [taps][N][K]
    int precision = 0;
    const float * X = nullptr; int ldx = 0;      // [M][ldx] (compact rows)
    const float * ln_w = nullptr; float eps = 1e-5f;   // LayerNorm prologue (no bias) if non-null
    const float * bias = nullptr;
    const float * res = nullptr; int ldr = 0;    // residual added in the epilogue
    float * Y = nullptr; int ldy = 0;
    int act = ACT_NONE, gelu_f16 = 1;
    const int32_t * tok_pos = nullptr;           // taps > 1: shifted rows valid iff tok_pos[t] >= shift
    // split store (QKV / cross KV): outputs [0,n_q) -> Y, [n_q, n_q+dkv) -> kdst, rest -> vdst,
    // at row tok_slot[t] of the KV storage (element type = model weight dtype)
    int n_q = -1, dkv = 0;
    void * kdst = nullptr; void * vdst = nullptr;
    const int32_t * tok_slot = nullptr;
    int M = 0;
    void * tc_scratch = nullptr; size_t tc_scratch_bytes = 0;     // activation tile images of the tensor-core path (gemm_tc.cu)
    bool x_prepacked = false;        // tc_scratch already holds the hi | lo tile images of X (written by the producing kernel)
    void * pack_out = nullptr;       // tensor-core path, M <= 64: write the output as hi | lo tile images here (Y may be null)
    // LayerNorm folded THROUGH the next GEMM (batched decoder step, gemm_ts.cu; removes the LN + pack launch between two layers):
    //   producer (res + pack_out + next_ln_w + stats_out): besides Y = acc + res it writes (Y .* next_ln_w) as hi | lo images into
    //     pack_out and every row's (sum, sum of squares) over the CTA's column slice into stats_out [N / ts_resid_nc(K)][64][2];
    //   consumer (x_prepacked + ln_fold_*): W . LN(x) = ((W . (x .* w)) - mean * csum) * rstd with csum[n] = sum_k W[n][k] w[k]
    //     (model.cu), mean / rstd from the slices' sums added in slice order (deterministic).
    const float * next_ln_w = nullptr; float * stats_out = nullptr;
    const float * ln_fold_stats = nullptr; int ln_fold_slices = 0; const float * ln_fold_csum = nullptr;
    // batched decoder step with f16 activation images (gemm_tc.cuh pack_act2): act_f16 = the X images of this GEMM are ONE f16 image
    // (its packing kernel writes that, its MMAs read A as f16); pack_f16 = the epilogue writes pack_out that way
    bool act_f16 = false, pack_f16 = false;
};
bool launch_linear(const LinearArgs & a, cudaStream_t stream);
// tcgen05 path (bf16, >= 16 tokens): gemm_tc.cu
size_t tc_weight_tile_bytes(int N, int K);
bool   tc_pack_weights(const void * W, int N, int K, void * Wt, cudaStream_t stream, int taps = 1, bool as_f16 = false);     // W: [taps][N][K]
size_t tc_scratch_bytes(int M, int K);
bool   tc_linear_supported(const LinearArgs & a);
bool   launch_linear_tc(const LinearArgs & a, cudaStream_t stream);
// token-stationary variant for <= 64 tokens (gemm_ts.cu): tokens on the MMA M side, 8..32 weight rows per CTA, 96-144 CTAs per GEMM
bool   ts_linear_supported(const LinearArgs & a);
bool   launch_linear_ts(const LinearArgs & a, const void * hi, const void * lo, cudaStream_t stream);
int    ts_resid_nc(int K);          // columns per CTA slice of the residual-epilogue GEMM with contraction width K (the granularity of stats_out)
bool   launch_row_dots(const void * W_bf16, const float * v, int N, int K, float * out, cudaStream_t stream);    // out[n] = sum_k W[n][k] v[k]

struct AttnArgs {
    int precision = 0;
    const float * q = nullptr; int ldq = 0;      // [M][ldq], head h at column h*dh
    const void * K = nullptr; const void * V = nullptr;   // [utt][rows_per_utt][H*dh]
    int rows_per_utt = 0;
    int H = 1, dh = 64;
    int causal = 1;                              // keys 0..tok_pos[t]; else keys 0..n_ctx[utt]-1
    const int32_t * n_ctx = nullptr;
    Tokens tok;
    float * out = nullptr; int ldo = 0;
    void * pack_out = nullptr;                   // dh == 64, <= 64 tokens: write hi | lo tile images for the next GEMM instead of `out`
    bool pack_f16 = false;                       // ... as ONE f16 image (LinearArgs::act_f16 of the consumer)
    int prefill_len = 0;                         // > 0: tokens are utterance-major runs of positions 0..prefill_len-1 (context prefill)
    const int32_t * page_table = nullptr; int max_pages = 0;   // paged K / V: [utterances][max_pages] page ids (pages of kKvPageRows rows); null = contiguous
    int kv_split = 0;                            // packed-output decoder step only: >= 1 = long-KV kernel, keys of a (head, token) divided over a cluster of this many CTAs
};
bool launch_attention(const AttnArgs & a, cudaStream_t stream);
int attention_plan_kv_split(int items, int max_keys);   // cluster size for a decoder step that will reach max_keys keys
// batched decoder step: folded cross-attention x += softmax(M LN(x)) N (tables from launch_xattn_fold, frame_loop.cu)
// pack_ln_w / pack_out (optional, B <= 64): additionally emit LN(x_new; pack_ln_w) as hi | lo tile images for the next GEMM
bool launch_xattn_folded(float * x, const float * ln_w, float eps, const float * xm, const float * xn, const int32_t * n_ctx, int B, int d,
                         int max_text, const float * pack_ln_w, void * pack_out, cudaStream_t stream, bool pack_f16 = false,
                         float * fold_stats = nullptr);      // fold_stats: LayerNorm folded through the consumer GEMM (LinearArgs::ln_fold_*), 4 slices

// x[t] = (sum_cb E_cb[codes[utt][cb]]) * 1/8 + dec_pos[pos]        (magpie.cpp:2746-2787, 4376-4379)
bool launch_audio_embed(const Model & m, const int32_t * codes /*[B][8] device*/, const int32_t * pos /*[B]*/,
                        int B, float * x, cudaStream_t stream);
// x[t] = baked_ctx[speaker[utt]][pos] + dec_pos[pos]               (magpie.cpp:4139-4165, 4228)
bool launch_context_embed(const Model & m, const int32_t * speakers, Tokens tok, float * x, cudaStream_t stream);
// x[t] = text_emb[token[t]] + enc_pos[pos]                         (magpie.cpp:1319-1345, 1974-1975)
bool launch_text_embed(const Model & m, const int32_t * tokens /*[M] compact*/, Tokens tok, float * x, cudaStream_t stream);
// y[t] = LN(x[t]) * w                                              (magpie.cpp:2237-2259)
bool launch_layer_norm(const float * x, const float * w, float eps, int M, int d, float * y, cudaStream_t stream);
bool launch_add_one(int32_t * v, int n, cudaStream_t stream);
// bf16 qkv [3L][L], o [L][L] -> bf16 [4L][L] = [Wq; Wk; hi(Wo Wv); lo(Wo Wv)]
// out[code] = [q | k | vo](LN(in_table[code] + pos; ln_w)) with the folded matrix qkvo ([Wq; Wk; hi; lo], bf16 [4L][L])
bool launch_lt_qkv_table(const float * in_table, const float * pos, const float * ln_w, float eps, const void * qkvo, int V, int L,
                         float * out, cudaStream_t stream);
bool launch_lt_fold_ov(const void * qkv, const void * o, int L, void * out, cudaStream_t stream);

// ---- local transformer + sampler (magpie.cpp:946-1048, 1072-1317) ------------------------------
struct LtArgs {
    int B = 0;
    const float * hidden = nullptr;      // [B][d] device
    float temperature = 0.0f; int top_k = 80;
    const uint8_t * forbid_eos = nullptr;   // [B] device or null
    int forbid_eos_all = 0;                 // applies to every utterance (generation loop: step < 4)
    const int32_t * forced = nullptr;    // [B][8] device or null
    const float * uniforms = nullptr;    // [B][8] device or null
    uint64_t seed = 0; uint32_t step = 0;
    int32_t * sampled = nullptr;         // [B][8] device: the model's picks
    int32_t * next_codes = nullptr;      // [B][8] device: codes the next decoder step consumes (forced or picked)
    int32_t * argmax = nullptr;          // [B][8] device
    float * logits = nullptr;            // [B][8][V] device or null
    int32_t * eos_flag = nullptr;        // [B] device or null: set to 1 when sampled/argmax hits EOS
    // loop mode (d_step != null): step = *d_step; per-step arrays are laid out [B][T_total][...] and
    // forced / uniforms / logits / sampled / argmax are indexed at (utt*T_total + step); `next_codes`
    // stays [B][8].  forbid_eos applies while step < min_frames.  done_step[utt] (init -1) records the
    // first step at which EOS was hit; hidden_hist (optional) receives the hidden state [B][T_total][d].
    const int32_t * d_step = nullptr;
    const int32_t * utt_step = nullptr;  // loop mode with per-utterance step counters [B] (continuous batching); takes precedence over d_step
    int T_total = 0, min_frames = 0;
    int32_t * done_step = nullptr;
    float * hidden_hist = nullptr;
    void * lt_scratch = nullptr; size_t lt_scratch_bytes = 0;     // activation scratch of the batched kernel (lt_batch.cu)
};
size_t lt_batch_scratch_bytes(const Model & m, int B);
bool launch_local_transformer(const Model & m, const LtArgs & a, cudaStream_t stream);

// ---- nano-codec ---------------------------------------------------------------------------------
bool codec_decode_device(Codec & c, const int32_t * d_codes, int B, int T, float * d_pcm, cudaStream_t stream);
bool codec_fsq_device(const int32_t * d_codes, int B, int T, float * d_latent, cudaStream_t stream);

}  // namespace mgb
