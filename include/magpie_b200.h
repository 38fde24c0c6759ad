/*
 * magpie_b200.h -- C-ABI of the B200-native (sm_100a) implementation of magpie-tts.cpp's
 * per-frame synthesis hot path.  Plain pointers and sizes only; no C++/torch/ggml types.
 *
 * This is the layer the reference's host code binds to instead of ggml: each entry point names
 * the reference function(s) it replaces (paths relative to the reference repo).  The C++ shim
 * `include/magpie.h` (same names/signatures as the reference's src/magpie.h) is written on top
 * of exactly these calls; INTEGRATION.md shows the binding a maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, a negative MGB_E* code on failure, unless it returns
 *     a handle (NULL on failure).  mgb_last_error() gives the message of the calling thread's
 *     last failure.  There is NO CPU fallback: without a CUDA device every call fails loudly.
 *   - host pointers are borrowed for the duration of the call; results are copied to host
 *     buffers supplied by the caller.  "[B][8]" means row-major, last index fastest.
 *   - a session owns the device state of B independent utterances (KV caches, encoder output,
 *     current position).  Utterances never interact: sessions on different devices need no
 *     collective.
 */
#ifndef MAGPIE_B200_H
#define MAGPIE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define MGB_API __attribute__((visibility("default")))
#else
#define MGB_API
#endif

enum {
    MGB_OK = 0,
    MGB_EINVAL = -1,   /* bad argument (reference: returns nullptr / false / empty vector) */
    MGB_EIO = -2,      /* file could not be read / parsed */
    MGB_ECUDA = -3,    /* CUDA runtime error or no device */
    MGB_ERANGE = -4    /* sequence/table capacity exceeded */
};

/* arithmetic the device path computes in */
enum {
    MGB_PREC_F32 = 0,   /* f32 weights + f32 KV cache: parity mode (rtol 1e-4 vs ggml CPU) */
    MGB_PREC_BF16 = 1   /* bf16 weights + bf16 KV cache, f32 activations/accumulation */
};

/* same fields, order and defaults as reference `struct magpie_hparams` (src/magpie.h:35-80) */
typedef struct mgb_hparams {
    int32_t d_model, d_ffn, d_head;
    int32_t enc_layers, enc_heads, enc_kernel;
    int32_t dec_layers, dec_sa_heads, dec_xa_heads, dec_xa_d_head, dec_kernel;
    int32_t lt_dim, lt_ffn_dim, lt_layers, lt_heads;
    int32_t text_vocab_size, num_codebooks, codebook_size, vocab_per_cb;
    int32_t num_speakers, context_frames;
    int32_t text_bos_id, text_eos_id, audio_bos_id, audio_eos_id;
    int32_t max_dec_steps, sample_rate;
    float   eps;
} mgb_hparams;

/* same fields as reference `struct magpie_codec_hparams` (src/magpie.h:655-664) */
typedef struct mgb_codec_hparams {
    int32_t sample_rate, num_codebooks, codebook_size, hop_length, latent_dim;
} mgb_codec_hparams;

typedef struct mgb_model   mgb_model;
typedef struct mgb_session mgb_session;
typedef struct mgb_codec   mgb_codec;

MGB_API const char * mgb_last_error(void);
MGB_API int          mgb_device_count(void);                 /* 0 if no CUDA device */
MGB_API const char * mgb_version(void);
/* Utterance -> device assignment of the multi-GPU path: independent utterances, weights
 * replicated per device, utterance i runs on device i mod n_devices; no collective on the data path. */
MGB_API int          mgb_shard_device(int64_t utterance_index, int n_devices);

/* ---- model -------------------------------------------------------------------------------
 * Replaces magpie_init / magpie_init_with_backend / magpie_free (src/magpie.cpp:777-915):
 * GGUF open, hparams (magpie.cpp:73-121), tensor name mapping (magpie.cpp:501-667), upload.
 * gelu_f16 != 0 reproduces ggml-CPU's f16 GELU table (default of the parity tests). */
MGB_API mgb_model * mgb_model_load(const char * gguf_path, int device, int precision);
MGB_API void        mgb_model_free(mgb_model * m);
MGB_API int         mgb_model_get_hparams(const mgb_model * m, mgb_hparams * out);
MGB_API int         mgb_model_set_max_dec_steps(mgb_model * m, int32_t n);   /* tests write hparams.max_dec_steps */
MGB_API int         mgb_model_set_gelu_f16(mgb_model * m, int on);
MGB_API int         mgb_model_precision(const mgb_model * m);
MGB_API int         mgb_model_device(const mgb_model * m);
/* bytes of unique weights one decoder step + local-transformer pass reads (roofline numerator) */
MGB_API int64_t     mgb_model_step_weight_bytes(const mgb_model * m);
/* tokenizer metadata strings (magpie.cpp:353-398); returns NULL if the key is absent */
MGB_API const char * mgb_model_meta_str(const mgb_model * m, const char * key);
MGB_API int32_t      mgb_model_meta_u32(const mgb_model * m, const char * key, int32_t def);

/* ---- session: B utterances -----------------------------------------------------------------
 * Replaces magpie_kv_cache_init_gpu / magpie_kv_cache_free_gpu (magpie.cpp:3315-3391).
 * max_text: capacity of encoder tokens per utterance; max_seq: KV slots per utterance
 * (reference uses context_frames + max_dec_steps + 16, magpie.cpp:4077; pass 0 for that). */
MGB_API mgb_session * mgb_session_new(mgb_model * m, int batch, int max_text, int max_seq);
/* Same, with an explicit page pool for the decoder self-attention cache: the cache is PAGED (pages of 128 positions, one page
 * table per utterance read by the attention kernels; the reference allocates one contiguous [12][max_seq][768] slab per call,
 * magpie.cpp:3315-3376).  kv_pages = 0 reserves ceil(max_seq / 128) pages per utterance up front (what mgb_session_new does);
 * kv_pages > 0 sizes the pool to that many pages shared by all utterances, handed out on demand as utterances grow and returned
 * when they finish (mgb_generate_queue) -- ragged batches then need memory for the frames they really produce. */
MGB_API mgb_session * mgb_session_new_paged(mgb_model * m, int batch, int max_text, int max_seq, int kv_pages);
MGB_API int           mgb_session_kv_pages(const mgb_session * s, int32_t * total, int32_t * in_use);
MGB_API void          mgb_session_free(mgb_session * s);
MGB_API int           mgb_session_batch(const mgb_session * s);
MGB_API int           mgb_session_max_seq(const mgb_session * s);
MGB_API int           mgb_session_positions(const mgb_session * s, int32_t * pos_out /*[B]*/);

/* magpie_encode_text (magpie.cpp:2284-2374): tokens [B][max_text] (padded), n_tokens [B].
 * enc_out (optional, host) receives [B][max_text][d_model]; rows >= n_tokens[b] are zero. */
MGB_API int mgb_encode_text(mgb_session * s, const int32_t * tokens, const int32_t * n_tokens, float * enc_out);

/* Steps 2-5 of magpie_synthesize_codes_graph_reuse (magpie.cpp:4089-4243): per-layer
 * cross-attention K/V (magpie.cpp:1663-1711) and the batched 110-frame context prefill
 * (magpie.cpp:3911-4060) for speaker ids [B].  Leaves position = context_frames. */
MGB_API int mgb_prefill(mgb_session * s, const int32_t * speakers);

/* One autoregressive decoder step for all B utterances (magpie.cpp:4366-4405, 3484-3528):
 * x = (sum_cb E_cb[code])/8 + pos[p]; 12 layers; final LayerNorm.  codes [B][8] host, or NULL
 * to consume the codes the previous mgb_lt_sample left on the device.  hidden_out optional
 * host [B][d_model].  Advances every utterance's position by one. */
MGB_API int mgb_decoder_step(mgb_session * s, const int32_t * codes, float * hidden_out);

/* magpie_build_final_proj (magpie.cpp:2261-2282): logits [B][8*vocab_per_cb] from the session's
 * current hidden state (hidden == NULL) or from host hidden [B][d_model]. */
MGB_API int mgb_final_proj(mgb_session * s, const float * hidden, float * logits_out);

/* magpie_local_transformer_sample_all (magpie.cpp:1113-1317) + sample_top_k (magpie.cpp:1072-1109),
 * all 8 codebooks of a frame in one kernel, for all B utterances.
 *   hidden       host [B][d_model] or NULL = session's current hidden state
 *   temperature  < 0.01 => greedy (sampled = argmax), as the reference
 *   forbid_eos   host [B] bytes or NULL (= none)
 *   forced_codes host [B][8] teacher forcing (code fed back for codebook cb), or NULL
 *   uniforms     host [B][8] uniform draws in [0,1) or NULL = Philox(seed, utterance, step, cb)
 *   sampled/argmax host [B][8] out (either may be NULL); logits_out host [B][8][V] or NULL
 * The sampled (or forced) codes stay on the device for the next mgb_decoder_step(codes=NULL). */
MGB_API int mgb_lt_sample(mgb_session * s, const float * hidden, float temperature, int top_k,
                          const uint8_t * forbid_eos, const int32_t * forced_codes,
                          const float * uniforms, uint64_t seed,
                          int32_t * sampled, int32_t * argmax, float * logits_out);

/* The generation loop of magpie_synthesize_codes_graph_reuse (magpie.cpp:4243-4432) after
 * mgb_encode_text + mgb_prefill: BOS step, then up to max_steps frames; per utterance stops at
 * EOS (sampled or argmax == audio_eos_id in any codebook; that frame is not emitted;
 * forbid_eos for the first min_frames=4 steps).  ignore_eos != 0 runs exactly max_steps frames
 * (benchmark mode).  codes_out host [B][max_steps][8] frame-major; n_frames_out host [B].
 * hidden_out optional host [B][max_steps][d_model] (hidden fed to the LT at each step). */
MGB_API int mgb_generate(mgb_session * s, int max_steps, float temperature, int top_k,
                         const float * uniforms /*[B][max_steps][8] or NULL*/, uint64_t seed,
                         int ignore_eos, int32_t * codes_out, int32_t * n_frames_out, float * hidden_out);

/* Continuous batching (BASELINE configs[3]/[4] with EOS enabled): n_utt utterances (n_utt >= the session's batch B) are streamed
 * through the B slots of the session.  Every step runs all B slots; a slot whose utterance hit EOS (magpie.cpp:4341-4352) or its
 * frame limit is retired at the next poll (every 8 steps): its codes are copied out, its cache pages go back to the pool and the
 * slot is refilled from the queue (text encoder + cross K/V + 110-frame prefill of that slot only), so the batch stays full and
 * the loop runs about sum(frames) / B steps instead of (n_utt / B) x max(frames).
 *   tokens [n_utt][max_text] (padded; max_text <= the session's), n_tokens / speakers [n_utt],
 *   max_steps_per_utt [n_utt] or NULL (= max_steps): per-utterance frame limits (ragged requests),
 *   codes_out [n_utt][max_steps][8], n_frames_out [n_utt], steps_run_out (optional): decoder steps the loop executed. */
MGB_API int mgb_generate_queue(mgb_session * s, int n_utt, const int32_t * tokens, const int32_t * n_tokens, int max_text,
                               const int32_t * speakers, const int32_t * max_steps_per_utt, int max_steps, float temperature,
                               int top_k, uint64_t seed, int32_t * codes_out, int32_t * n_frames_out, int64_t * steps_run_out);

/* Teacher-forced run (BASELINE config 2): feeds BOS then codes_in [B][T][8]; outputs per step the
 * decoder hidden [B][T][d_model] (optional), the 8 masked LT logit vectors [B][T][8][V] (optional)
 * and the greedy codes [B][T][8] (optional).  Step t consumes frame t-1 (BOS for t = 0). */
MGB_API int mgb_teacher_forced(mgb_session * s, const int32_t * codes_in, int T,
                               float * hidden_out, float * lt_logits_out, int32_t * greedy_out);

/* device time (ms) of the last mgb_generate / mgb_teacher_forced loop, measured with CUDA events
 * on the session's stream, and the number of kernel launches it issued */
MGB_API float   mgb_session_last_loop_ms(const mgb_session * s);
MGB_API int64_t mgb_session_last_loop_launches(const mgb_session * s);
/* profiling aid (sessions created with MGB_MEGA_DBG=1): clock64() stamps CTA 0 of the batch-1 decoder
 * megakernel took around every grid barrier of the last step (16 per layer + 2) */
MGB_API int mgb_session_debug_stamps(mgb_session * s, uint64_t * out, int n);

/* ---- batched streaming synthesis --------------------------------------------------------------------
 * magpie_synthesize_sentence_streaming (magpie.cpp:4502-4829; chunk decode 4483-4500) for all B utterances of a prefilled
 * session at once (BASELINE configs[4]: long-form streaming): the device loop runs frames_per_chunk frames (reference default 4),
 * the chunks of all utterances are decoded by the codec in ONE batched launch sequence and every utterance's new samples are
 * handed to the callback (utterance, pcm, n_samples, frames so far, is_last, user); a non-zero return stops the call.
 * codec_context_frames = 0 decodes every chunk with zero causal history, as the reference does (audible seam per chunk);
 * N > 0 decodes each chunk together with the utterance's previous N frames and emits only the new samples (overlap-save;
 * N >= 25 covers the codec's 24.8-frame receptive field, so the stream equals a whole-utterance decode).  As in the reference's
 * streaming path the EOS frame's codes are decoded too (magpie.cpp:4733-4742).  n_frames_out [B] optional. */
typedef int (*mgb_stream_callback)(int utterance, const float * pcm, int n_samples, int frames_done, int is_last, void * user);
MGB_API int mgb_stream_generate(mgb_session * s, mgb_codec * codec, int max_steps, float temperature, int top_k, uint64_t seed,
                                int ignore_eos, int frames_per_chunk, int codec_context_frames, mgb_stream_callback cb, void * user,
                                int32_t * n_frames_out);

/* ---- in-process multi-GPU pool ---------------------------------------------------------------------
 * The reference is single-device (src/magpie.cpp:14-67 picks ONE backend; magpie_synthesize_codes_graph_reuse,
 * magpie.cpp:4063-4432, synthesises one utterance).  Independent utterances shard with no collective: a pool
 * owns one model replica per device and, per call, one session + one submission thread (one stream) per device; utterance i
 * runs on device mgb_shard_device(i, n_devices).  devices == NULL / n_devices <= 0 = all visible devices.  No NCCL.
 *   tokens [n_utt][max_text] (padded), n_tokens [n_utt], speakers [n_utt] or NULL (= 0)
 *   mgb_pool_generate:       codes_out [n_utt][max_steps][8], n_frames_out [n_utt]   (as mgb_generate, per utterance)
 *   mgb_pool_teacher_forced: codes_in [n_utt][T][8] -> greedy_out [n_utt][T][8] (optional)
 *   device_ms_out (optional) [n_devices]: device time of each device's loop (CUDA events on its stream) */
typedef struct mgb_pool mgb_pool;
MGB_API mgb_pool *  mgb_pool_new(const char * gguf_path, const int * devices, int n_devices, int precision);
MGB_API void        mgb_pool_free(mgb_pool * p);
MGB_API int         mgb_pool_n_devices(const mgb_pool * p);
MGB_API mgb_model * mgb_pool_model(mgb_pool * p, int i);          /* replica i (borrowed) */
MGB_API int mgb_pool_generate(mgb_pool * p, int n_utt, const int32_t * tokens, const int32_t * n_tokens, int max_text,
                              const int32_t * speakers, int max_steps, float temperature, int top_k, uint64_t seed, int ignore_eos,
                              int32_t * codes_out, int32_t * n_frames_out, float * device_ms_out);
MGB_API int mgb_pool_teacher_forced(mgb_pool * p, int n_utt, const int32_t * tokens, const int32_t * n_tokens, int max_text,
                                    const int32_t * speakers, const int32_t * codes_in, int T, int32_t * greedy_out,
                                    float * device_ms_out);

/* ---- nano-codec ----------------------------------------------------------------------------
 * magpie_codec_init / _free (src/nano-codec.cpp:339-374, 205-333) and magpie_codec_decode
 * (nano-codec.cpp:758-845) incl. fsq_dequantize_cpu (nano-codec.cpp:721-752). */
MGB_API mgb_codec * mgb_codec_load(const char * gguf_path, int device);
MGB_API void        mgb_codec_free(mgb_codec * c);
MGB_API int         mgb_codec_get_hparams(const mgb_codec * c, mgb_codec_hparams * out);
/* codes host [B][8][T] codebook-major int32 -> pcm host [B][T*hop] float32 */
MGB_API int mgb_codec_decode(mgb_codec * c, const int32_t * codes, int batch, int n_frames, float * pcm_out);
/* FSQ only: codes [B][8][T] -> latent [B][32][T] (bit-exact vs nano-codec.cpp:721-752) */
MGB_API int mgb_codec_fsq_dequantize(mgb_codec * c, const int32_t * codes, int batch, int n_frames, float * latent_out);
MGB_API float   mgb_codec_last_ms(const mgb_codec * c);
MGB_API int64_t mgb_codec_last_launches(const mgb_codec * c);

#ifdef __cplusplus
}
#endif
#endif /* MAGPIE_B200_H */
