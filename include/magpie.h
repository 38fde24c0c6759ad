// magpie.h -- drop-in replacement for the reference's src/magpie.h (m1el/magpie-tts.cpp) on the
// synthesis hot path, backed by the B200-native library (include/magpie_b200.h) instead of ggml.
//
// Kept from the reference header (same names, argument meaning, return/err behaviour):
//   enums/structs   magpie_backend_type (magpie.h:24-29), magpie_hparams (:35-80), magpie_tokenizer (:86-104),
//                   magpie_context public fields n_threads/temperature/top_k/speaker_id/codec (:290-308),
//                   magpie_sample_result (:310-313), magpie_stream_params + callbacks (:604-628),
//                   magpie_codec_hparams (:655-677)
//   functions       magpie_init / _init_with_backend / _free / _get_backend_name (:316-329), magpie_tokenize (:110),
//                   magpie_encode_text (:555), magpie_synthesize_codes{,_cached,_optimized,_graph_reuse} (:571-595),
//                   magpie_local_transformer_sample_all (:372-377), magpie_is_eos (:823),
//                   magpie_split_sentences / magpie_synthesize_streaming / _sentence_streaming (:631-648),
//                   magpie_codec_init / _init_with_backend / _free / magpie_codec_decode (:745-759)
// Not kept: every magpie_build_* / magpie_codec_build_* graph builder (they take and return ggml_tensor*;
// ggml-graph internals are not part of the drop-in surface) and the weight structs' ggml_tensor* members,
// which are replaced by an opaque device handle.  Errors: nullptr / empty vector / false / -1 plus a line on
// stderr, exactly as the reference; there is no CPU fallback.
#ifndef MAGPIE_H
#define MAGPIE_H

#include <cstdint>
#include <map>
#include <string>
#include <vector>

#if defined(__GNUC__)
#define MAGPIE_API __attribute__((visibility("default")))
#else
#define MAGPIE_API
#endif

enum magpie_backend_type {
    MAGPIE_BACKEND_CPU   = 0,   // not available in this build (fails loudly)
    MAGPIE_BACKEND_CUDA  = 1,
    MAGPIE_BACKEND_METAL = 2,   // not available in this build
    MAGPIE_BACKEND_AUTO  = 3,   // = CUDA
};

struct magpie_hparams {
    int32_t d_model = 768, d_ffn = 3072, d_head = 64;
    int32_t enc_layers = 6, enc_heads = 12, enc_kernel = 3;
    int32_t dec_layers = 12, dec_sa_heads = 12, dec_xa_heads = 1, dec_xa_d_head = 128, dec_kernel = 1;
    int32_t lt_dim = 256, lt_ffn_dim = 1024, lt_layers = 1, lt_heads = 1;
    int32_t text_vocab_size = 2380, num_codebooks = 8, codebook_size = 2016, vocab_per_cb = 2024;
    int32_t num_speakers = 5, context_frames = 110;
    int32_t text_bos_id = 2378, text_eos_id = 2379, audio_bos_id = 2016, audio_eos_id = 2017;
    int32_t max_dec_steps = 500, sample_rate = 22050;
    float   eps = 1e-5f;
};

struct magpie_tokenizer {
    std::vector<std::string> vocab;
    std::map<std::string, int32_t> token_to_id;
    std::map<std::string, std::string> dict;
    int32_t pad_id = -1, oov_id = -1, space_id = -1, bos_id = -1, eos_id = -1;
    bool loaded = false;
};

MAGPIE_API std::vector<int32_t> magpie_tokenize(const magpie_tokenizer * tok, const std::string & text);

struct magpie_model_impl;     // device weights (opaque)
struct magpie_model {
    magpie_hparams   hparams;
    magpie_tokenizer tokenizer;
    magpie_backend_type backend_type = MAGPIE_BACKEND_CUDA;
    magpie_model_impl * impl = nullptr;
};

// reference src/magpie.h:237-259 without the ggml tensor members (the device cache is owned by the session behind
// magpie_model_impl); the public scalars and reset() are kept
struct magpie_kv_cache {
    int32_t seq_len = 0;      // current sequence length
    int32_t max_seq = 0;      // context_frames + max_dec_steps + 16 (reference magpie.cpp:4077)
    int32_t enc_seq_len = 0;  // encoder sequence length (cross-attention)
    void reset() { seq_len = 0; enc_seq_len = 0; }
};

// reference src/magpie.h:265-288 without the ggml allocator
struct magpie_state {
    magpie_kv_cache kv_cache;
    std::vector<int32_t> generated_codes;  // [n_frames][num_codebooks] of the last synthesis call
    int32_t n_generated_frames = 0;
    std::vector<float> encoder_output;     // [enc_seq][d_model], filled by magpie_encode_text
    int32_t enc_seq_len = 0;
    void reset() {
        kv_cache.reset();
        generated_codes.clear();
        n_generated_frames = 0;
        encoder_output.clear();
        enc_seq_len = 0;
    }
};

struct magpie_codec;

struct magpie_context {
    magpie_model model;
    magpie_state state;
    int   n_threads;          // kept for source compatibility; unused (as in the reference)
    float temperature;
    int   top_k;
    int   speaker_id;
    magpie_codec * codec;
    magpie_context() : n_threads(4), temperature(0.7f), top_k(80), speaker_id(0), codec(nullptr) {}
};

struct magpie_sample_result {
    std::vector<int32_t> sampled_codes;
    std::vector<int32_t> argmax_codes;
};

MAGPIE_API magpie_context * magpie_init(const char * model_path);
MAGPIE_API magpie_context * magpie_init_with_backend(const char * model_path, magpie_backend_type backend);
MAGPIE_API void             magpie_free(magpie_context * ctx);
MAGPIE_API const char *     magpie_get_backend_name(magpie_context * ctx);

MAGPIE_API bool magpie_encode_text(magpie_context * ctx, const int32_t * tokens, int n_tokens);

// All four variants run the same device-resident loop (the reference's CLI path is _graph_reuse).
// Returns codes frame-major [n_frames][8]; empty on error.
MAGPIE_API std::vector<int32_t> magpie_synthesize_codes(magpie_context * ctx, const int32_t * tokens, int n_tokens);
MAGPIE_API std::vector<int32_t> magpie_synthesize_codes_cached(magpie_context * ctx, const int32_t * tokens, int n_tokens);
MAGPIE_API std::vector<int32_t> magpie_synthesize_codes_optimized(magpie_context * ctx, const int32_t * tokens, int n_tokens);
MAGPIE_API std::vector<int32_t> magpie_synthesize_codes_graph_reuse(magpie_context * ctx, const int32_t * tokens, int n_tokens);

MAGPIE_API magpie_sample_result magpie_local_transformer_sample_all(magpie_context * ctx, const float * decoder_hidden,
                                                         float temperature, int top_k, bool forbid_eos = false);

// reference signature (src/magpie.h:823, magpie.cpp:3273) and a pointer overload for callers that hold a raw frame
MAGPIE_API bool magpie_is_eos(const std::vector<int32_t> & frame_codes, int32_t eos_id);
MAGPIE_API bool magpie_is_eos(const int32_t * codes, int n_codebooks, int eos_id);

// ---- streaming -------------------------------------------------------------------------------------
typedef bool (*magpie_audio_callback)(const float * samples, int n_samples, void * user_data);
typedef void (*magpie_progress_callback)(int frames_generated, int sentence_index, int total_sentences, void * user_data);

struct magpie_stream_params {
    float temperature = 0.7f;
    int   top_k = 80;
    int   speaker_id = 0;
    int   frames_per_chunk = 4;
    bool  sentence_chunking = true;
    magpie_audio_callback    on_audio = nullptr;
    magpie_progress_callback on_progress = nullptr;
    void * user_data = nullptr;
    // Extension (appended, default = reference behaviour): 0 decodes every chunk with zero causal history, as the reference does
    // (magpie.cpp:4482-4500, audible seams every chunk); N > 0 decodes each chunk together with the previous N frames of codes and
    // emits only the new samples.  The codec's receptive field is 24.8 frames, so N >= 25 makes the streamed
    // audio identical to a whole-utterance decode.
    int   codec_context_frames = 0;
};

MAGPIE_API std::vector<std::string> magpie_split_sentences(const char * text);
MAGPIE_API int magpie_synthesize_streaming(magpie_context * ctx, magpie_codec * codec, const char * text, const magpie_stream_params & params);
MAGPIE_API int magpie_synthesize_sentence_streaming(magpie_context * ctx, magpie_codec * codec, const int32_t * tokens, int n_tokens,
                                         const magpie_stream_params & params);

// ---- audio codec -----------------------------------------------------------------------------------
struct magpie_codec_hparams {
    int32_t sample_rate = 22050, num_codebooks = 8, codebook_size = 2016, hop_length = 1024, latent_dim = 32;
    int32_t fsq_levels[4] = {8, 7, 6, 6};
    int32_t pre_conv_kernel = 7, post_conv_kernel = 3, base_channels = 864;
    int32_t num_upsample_layers = 5;
    int32_t up_sample_rates[5] = {8, 8, 4, 2, 2};
    int32_t up_channels[5] = {432, 216, 108, 54, 27};
    int32_t resblock_kernel_sizes[3] = {3, 7, 11};
    int32_t resblock_dilations[3] = {1, 3, 5};
};

struct magpie_codec_impl;
struct magpie_codec {
    magpie_codec_hparams hparams;
    magpie_backend_type backend_type = MAGPIE_BACKEND_CUDA;
    magpie_codec_impl * impl = nullptr;
};

MAGPIE_API magpie_codec * magpie_codec_init(const char * codec_path);
MAGPIE_API magpie_codec * magpie_codec_init_with_backend(const char * codec_path, magpie_backend_type backend);
MAGPIE_API void           magpie_codec_free(magpie_codec * codec);
// codes [num_codebooks][n_frames] codebook-major -> n_frames * hop_length samples; empty on error
MAGPIE_API std::vector<float> magpie_codec_decode(magpie_codec * codec, const int32_t * codes, int n_frames);
// B200 extension: decode `batch` independent chunks [batch][8][n_frames] in one launch sequence
MAGPIE_API std::vector<float> magpie_codec_decode_batch(magpie_codec * codec, const int32_t * codes, int batch, int n_frames);

#endif  // MAGPIE_H
