// Host side of include/magpie.h: the reference's API surface (src/magpie.h) written over the C-ABI of
// libmagpie_b200.so.  Host-only logic lives here: text normalisation + tokenizer (reference
// src/magpie.cpp:127-495), sentence splitting and the streaming driver (magpie.cpp:4438-4863), and the
// generation-loop bookkeeping.  All arithmetic on model data happens in the sm_100a kernels.
#include "../../include/magpie.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>

#include "../../include/magpie_b200.h"

struct magpie_model_impl {
    mgb_model *   model = nullptr;
    mgb_session * lt_session = nullptr;    // scratch session for stand-alone magpie_local_transformer_sample_all calls
    mgb_session * enc_session = nullptr;   // session of the last magpie_encode_text
    uint64_t seed = 0;
    uint64_t draws = 0;
};
struct magpie_codec_impl { mgb_codec * codec = nullptr; };

namespace {

// ---- text normalisation (behaviour of normalize_text, magpie.cpp:260-349) ---------------------------
const char * const kSmall[] = {"zero", "one", "two", "three", "four", "five", "six", "seven", "eight", "nine", "ten",
                               "eleven", "twelve", "thirteen", "fourteen", "fifteen", "sixteen", "seventeen", "eighteen", "nineteen"};
const char * const kTens[] = {"", "", "twenty", "thirty", "forty", "fifty", "sixty", "seventy", "eighty", "ninety"};

std::string cardinal(int64_t n) {
    if (n < 0) return "minus " + cardinal(-n);
    if (n < 20) return kSmall[n];
    if (n < 100) return n % 10 ? std::string(kTens[n / 10]) + " " + kSmall[n % 10] : std::string(kTens[n / 10]);
    if (n < 1000) {
        std::string s = std::string(kSmall[n / 100]) + " hundred";
        if (n % 100) s += " and " + cardinal(n % 100);
        return s;
    }
    struct Scale { int64_t unit; const char * name; };
    static const Scale scales[] = {{1000000000LL, "billion"}, {1000000LL, "million"}, {1000LL, "thousand"}};
    if (n >= 1000000000000LL) return std::to_string(n);
    for (const Scale & sc : scales)
        if (n >= sc.unit) {
            std::string s = cardinal(n / sc.unit) + " " + sc.name;
            if (n % sc.unit) s += " " + cardinal(n % sc.unit);
            return s;
        }
    return std::to_string(n);
}

std::string year_words(int64_t n) {            // 2024 -> "twenty twenty four", 1900 -> "nineteen hundred"
    if (n < 1000 || n > 9999) return cardinal(n);
    const int hi = (int)(n / 100), lo = (int)(n % 100);
    if (lo == 0) return cardinal(hi) + " hundred";
    if (lo < 10) return cardinal(n);
    return cardinal(hi) + " " + cardinal(lo);
}

std::string ordinal(int64_t n) {
    static const char * const first12[] = {"", "first", "second", "third", "fourth", "fifth", "sixth", "seventh",
                                           "eighth", "ninth", "tenth", "eleventh", "twelfth"};
    if (n >= 1 && n <= 12) return first12[n];
    std::string c = cardinal(n);
    if (n >= 13 && n <= 19) return c + "th";
    if (n % 10 == 0 && n >= 20 && n < 100) return (!c.empty() && c.back() == 'y') ? c.substr(0, c.size() - 1) + "ieth" : c + "th";
    const int unit = (int)(n % 10);
    if (unit >= 1 && unit <= 3) {
        static const char * const repl[] = {"", "first", "second", "third"};
        return c.substr(0, c.rfind(' ') + 1) + repl[unit];      // rfind == npos -> +1 wraps to 0, as in the reference
    }
    return c + "th";
}

inline bool is_digit(char c) { return c >= '0' && c <= '9'; }
inline char lower(char c) { return (c >= 'A' && c <= 'Z') ? (char)(c - 'A' + 'a') : c; }

std::string normalize(const std::string & t) {
    std::string out;
    out.reserve(t.size() * 2);
    size_t i = 0;
    auto read_int = [&](int & digits) {
        int64_t v = 0; digits = 0;
        while (i < t.size() && is_digit(t[i])) { v = v * 10 + (t[i] - '0'); digits++; i++; }
        return v;
    };
    while (i < t.size()) {
        int nd = 0;
        if (t[i] == '$' && i + 1 < t.size() && is_digit(t[i + 1])) {            // $50 -> fifty dollars
            i++;
            const int64_t v = read_int(nd);
            out += cardinal(v) + (v == 1 ? " dollar" : " dollars");
            continue;
        }
        const bool neg = t[i] == '-' && i + 1 < t.size() && is_digit(t[i + 1]);
        if (is_digit(t[i]) || neg) {
            if (neg) i++;
            const int64_t v = read_int(nd);
            if (i < t.size() && t[i] == '%') {                                    // 50% -> fifty percent
                i++;
                out += (neg ? "minus " : "") + cardinal(v) + " percent";
                continue;
            }
            bool ord = false;
            if (i + 1 < t.size()) {
                const char a = lower(t[i]), b = lower(t[i + 1]);
                ord = (a == 's' && b == 't') || (a == 'n' && b == 'd') || (a == 'r' && b == 'd') || (a == 't' && b == 'h');
                if (ord) i += 2;
            }
            std::string w = ord ? ordinal(v) : ((nd == 4 && v >= 1000 && v <= 2099) ? year_words(v) : cardinal(v));
            if (neg && v != 0) w = "minus " + w;
            out += w;
            continue;
        }
        out += t[i++];
    }
    return out;
}

std::vector<std::string> split(const std::string & s, char delim) {
    std::vector<std::string> parts;
    size_t a = 0;
    for (size_t b = s.find(delim); b != std::string::npos; b = s.find(delim, a)) { parts.push_back(s.substr(a, b - a)); a = b + 1; }
    parts.push_back(s.substr(a));
    return parts;
}

bool tokenizer_from_model(magpie_tokenizer & tok, mgb_model * m, const magpie_hparams & hp) {
    const char * vocab = mgb_model_meta_str(m, "magpie.tokenizer.vocab");
    if (!vocab) { fprintf(stderr, "magpie_tokenizer: vocabulary not found in model\n"); return false; }
    tok.vocab = split(vocab, '\n');
    for (size_t i = 0; i < tok.vocab.size(); i++) tok.token_to_id[tok.vocab[i]] = (int32_t)i;
    if (const char * dict = mgb_model_meta_str(m, "magpie.tokenizer.dict"))
        for (const std::string & line : split(dict, '\n')) {
            const size_t tab = line.find('\t');
            if (tab != std::string::npos) tok.dict[line.substr(0, tab)] = line.substr(tab + 1);
        }
    tok.pad_id = mgb_model_meta_u32(m, "magpie.tokenizer.pad", 94);        // defaults: magpie.cpp:392-394
    tok.oov_id = mgb_model_meta_u32(m, "magpie.tokenizer.oov", 95);
    tok.space_id = mgb_model_meta_u32(m, "magpie.tokenizer.space", 93);
    tok.bos_id = hp.text_bos_id;
    tok.eos_id = hp.text_eos_id;
    tok.loaded = true;
    return true;
}

void copy_hparams(magpie_hparams & d, const mgb_hparams & s) {
#define F(x) d.x = s.x
    F(d_model); F(d_ffn); F(d_head); F(enc_layers); F(enc_heads); F(enc_kernel); F(dec_layers); F(dec_sa_heads);
    F(dec_xa_heads); F(dec_xa_d_head); F(dec_kernel); F(lt_dim); F(lt_ffn_dim); F(lt_layers); F(lt_heads);
    F(text_vocab_size); F(num_codebooks); F(codebook_size); F(vocab_per_cb); F(num_speakers); F(context_frames);
    F(text_bos_id); F(text_eos_id); F(audio_bos_id); F(audio_eos_id); F(max_dec_steps); F(sample_rate); F(eps);
#undef F
}

int env_precision() {
    // default bf16: the B200 production path (persistent frame-loop kernel at batch 1, tcgen05 GEMMs when batched);
    // MAGPIE_PRECISION=f32 selects the f32 parity mode (1e-4 against the reference's f32 CPU arithmetic)
    const char * e = getenv("MAGPIE_PRECISION");
    return (e && (!strcmp(e, "f32") || !strcmp(e, "F32") || !strcmp(e, "fp32"))) ? MGB_PREC_F32 : MGB_PREC_BF16;
}
int env_device() { const char * e = getenv("MAGPIE_DEVICE"); return e ? atoi(e) : 0; }

// encode + prefill a fresh one-utterance session sized for this request; nullptr on error
mgb_session * start_utterance(magpie_context * ctx, const int32_t * tokens, int n_tokens, int max_steps, bool want_enc) {
    const magpie_hparams & hp = ctx->model.hparams;
    if (!ctx->model.impl || !tokens || n_tokens <= 0) { fprintf(stderr, "magpie: invalid args\n"); return nullptr; }
    mgb_model_set_max_dec_steps(ctx->model.impl->model, hp.max_dec_steps);
    const int max_seq = hp.context_frames + max_steps + 16;
    mgb_session * s = mgb_session_new(ctx->model.impl->model, 1, n_tokens, max_seq);
    if (!s) { fprintf(stderr, "magpie: %s\n", mgb_last_error()); return nullptr; }
    if (want_enc) ctx->state.encoder_output.assign((size_t)n_tokens * hp.d_model, 0.0f);
    const int32_t n = n_tokens, spk = ctx->speaker_id;
    if (mgb_encode_text(s, tokens, &n, want_enc ? ctx->state.encoder_output.data() : nullptr) != MGB_OK ||
        mgb_prefill(s, &spk) != MGB_OK) {
        fprintf(stderr, "magpie: %s\n", mgb_last_error());
        mgb_session_free(s);
        return nullptr;
    }
    ctx->state.enc_seq_len = n_tokens;
    ctx->state.kv_cache.enc_seq_len = n_tokens;
    ctx->state.kv_cache.max_seq = max_seq;
    ctx->state.kv_cache.seq_len = hp.context_frames;
    return s;
}

std::vector<int32_t> synthesize(magpie_context * ctx, const int32_t * tokens, int n_tokens) {
    std::vector<int32_t> out;
    if (!ctx) { fprintf(stderr, "magpie_synthesize_codes: invalid args\n"); return out; }
    const magpie_hparams & hp = ctx->model.hparams;
    const int T = hp.max_dec_steps;
    mgb_session * s = start_utterance(ctx, tokens, n_tokens, T, true);
    if (!s) return out;
    std::vector<int32_t> codes((size_t)T * 8);
    int32_t n_frames = 0;
    magpie_model_impl * im = ctx->model.impl;
    const int rc = mgb_generate(s, T, ctx->temperature, ctx->top_k, nullptr, im->seed + (im->draws++), 0, codes.data(), &n_frames, nullptr);
    const float ms = mgb_session_last_loop_ms(s);
    mgb_session_free(s);
    if (rc != MGB_OK) { fprintf(stderr, "magpie: %s\n", mgb_last_error()); return out; }
    ctx->state.kv_cache.seq_len = hp.context_frames + 1 + n_frames;
    ctx->state.n_generated_frames = n_frames;
    ctx->state.generated_codes.assign(codes.begin(), codes.begin() + (size_t)n_frames * 8);
    if (ms > 0.0f) fprintf(stderr, "magpie: generated %d frames in %.1f ms (%.1f frames/s)\n", n_frames, ms, n_frames * 1e3f / ms);
    out.assign(codes.begin(), codes.begin() + (size_t)n_frames * 8);
    return out;
}

}  // namespace

// ---- tokenizer --------------------------------------------------------------------------------------
std::vector<int32_t> magpie_tokenize(const magpie_tokenizer * tok, const std::string & text) {
    std::vector<int32_t> ids;
    if (!tok || !tok->loaded) { fprintf(stderr, "magpie_tokenize: tokenizer not loaded\n"); return ids; }
    ids.push_back(tok->bos_id);
    std::string spaced;
    for (char c : normalize(text)) {
        c = lower(c);
        if (c == ',' || c == '.' || c == '!' || c == '?' || c == ':' || c == ';') { spaced += ' '; spaced += c; spaced += ' '; }
        else spaced += c;
    }
    auto lookup = [&](const std::string & s) { auto it = tok->token_to_id.find(s); return it == tok->token_to_id.end() ? -1 : it->second; };
    for (const std::string & word : split(spaced, ' ')) {
        if (word.empty()) continue;
        if (word.size() == 1) {                       // single-character token (punctuation): no space appended
            const int32_t id = lookup(word);
            if (id >= 0) { ids.push_back(id); continue; }
        }
        auto pron = tok->dict.find(word);
        if (pron != tok->dict.end()) {                // IPA string: greedy longest match over <= 4 bytes
            const std::string & p = pron->second;
            for (size_t i = 0; i < p.size();) {
                size_t adv = 1;
                for (size_t len = std::min<size_t>(4, p.size() - i); len > 0; len--) {
                    const int32_t id = lookup(p.substr(i, len));
                    if (id >= 0) { ids.push_back(id); adv = len; break; }
                }
                i += adv;
            }
        } else {                                      // OOV: per-character upper-case fallback
            for (char c : word) {
                const char up = (c >= 'a' && c <= 'z') ? (char)(c - 'a' + 'A') : c;
                const int32_t id = lookup(std::string(1, up));
                if (id >= 0) ids.push_back(id);
            }
        }
        if (tok->space_id >= 0) ids.push_back(tok->space_id);
    }
    if (!ids.empty() && ids.back() == tok->space_id) ids.pop_back();
    ids.push_back(tok->eos_id);
    return ids;
}

// ---- context ----------------------------------------------------------------------------------------
magpie_context * magpie_init(const char * model_path) { return magpie_init_with_backend(model_path, MAGPIE_BACKEND_AUTO); }

magpie_context * magpie_init_with_backend(const char * model_path, magpie_backend_type backend) {
    if (backend == MAGPIE_BACKEND_CPU || backend == MAGPIE_BACKEND_METAL) {
        fprintf(stderr, "magpie_init: this build is CUDA (sm_100a) only; the CPU/Metal backends are not available\n");
        return nullptr;
    }
    if (!model_path) { fprintf(stderr, "magpie_init: null model path\n"); return nullptr; }
    mgb_model * m = mgb_model_load(model_path, env_device(), env_precision());
    if (!m) { fprintf(stderr, "magpie_init: %s\n", mgb_last_error()); return nullptr; }
    // MAGPIE_GELU_TABLE=0: plain f32 tanh-GELU instead of ggml-CPU's f16 lookup table (default on: the reference's CPU semantics)
    if (const char * e = getenv("MAGPIE_GELU_TABLE")) mgb_model_set_gelu_f16(m, atoi(e) != 0);
    magpie_context * ctx = new magpie_context();
    mgb_hparams hp;
    mgb_model_get_hparams(m, &hp);
    copy_hparams(ctx->model.hparams, hp);
    ctx->model.impl = new magpie_model_impl();
    ctx->model.impl->model = m;
    if (const char * e = getenv("MAGPIE_SEED")) ctx->model.impl->seed = strtoull(e, nullptr, 10);
    else { std::random_device rd; ctx->model.impl->seed = ((uint64_t)rd() << 32) ^ rd(); }    // reference: unseeded mt19937 (magpie.cpp:1129)
    if (!tokenizer_from_model(ctx->model.tokenizer, m, ctx->model.hparams)) {
        fprintf(stderr, "magpie_init: warning: tokenizer not loaded, text input will not work\n");     // magpie.cpp:817-820
    }
    ctx->state.kv_cache.max_seq = hp.context_frames + hp.max_dec_steps + 16;
    return ctx;
}

void magpie_free(magpie_context * ctx) {
    if (!ctx) return;
    if (ctx->model.impl) {
        if (ctx->model.impl->lt_session) mgb_session_free(ctx->model.impl->lt_session);
        if (ctx->model.impl->enc_session) mgb_session_free(ctx->model.impl->enc_session);
        mgb_model_free(ctx->model.impl->model);
        delete ctx->model.impl;
    }
    delete ctx;
}

const char * magpie_get_backend_name(magpie_context * ctx) { return ctx ? "CUDA (sm_100a, B200-native)" : "none"; }

bool magpie_encode_text(magpie_context * ctx, const int32_t * tokens, int n_tokens) {
    if (!ctx || !ctx->model.impl || !tokens || n_tokens <= 0) { fprintf(stderr, "magpie_encode_text: invalid args\n"); return false; }
    magpie_model_impl * im = ctx->model.impl;
    if (im->enc_session) { mgb_session_free(im->enc_session); im->enc_session = nullptr; }
    im->enc_session = mgb_session_new(im->model, 1, n_tokens, 0);
    if (!im->enc_session) { fprintf(stderr, "magpie_encode_text: %s\n", mgb_last_error()); return false; }
    ctx->state.encoder_output.assign((size_t)n_tokens * ctx->model.hparams.d_model, 0.0f);
    const int32_t n = n_tokens;
    if (mgb_encode_text(im->enc_session, tokens, &n, ctx->state.encoder_output.data()) != MGB_OK) {
        fprintf(stderr, "magpie_encode_text: %s\n", mgb_last_error());
        return false;
    }
    ctx->state.enc_seq_len = n_tokens;
    return true;
}

std::vector<int32_t> magpie_synthesize_codes(magpie_context * c, const int32_t * t, int n) { return synthesize(c, t, n); }
std::vector<int32_t> magpie_synthesize_codes_cached(magpie_context * c, const int32_t * t, int n) { return synthesize(c, t, n); }
std::vector<int32_t> magpie_synthesize_codes_optimized(magpie_context * c, const int32_t * t, int n) { return synthesize(c, t, n); }
std::vector<int32_t> magpie_synthesize_codes_graph_reuse(magpie_context * c, const int32_t * t, int n) { return synthesize(c, t, n); }

magpie_sample_result magpie_local_transformer_sample_all(magpie_context * ctx, const float * hidden, float temperature, int top_k,
                                                         bool forbid_eos) {
    magpie_sample_result r;
    if (!ctx || !ctx->model.impl || !hidden) { fprintf(stderr, "magpie_local_transformer_sample_all: invalid args\n"); return r; }
    magpie_model_impl * im = ctx->model.impl;
    if (!im->lt_session) im->lt_session = mgb_session_new(im->model, 1, 8, ctx->model.hparams.context_frames + 2);
    if (!im->lt_session) { fprintf(stderr, "magpie_local_transformer_sample_all: %s\n", mgb_last_error()); return r; }
    if (temperature >= 0.01f && top_k < 1) top_k = 1;      // top_k = 0 reads scored[-1] in the reference (UB): clamp
    int32_t s[8], a[8];
    const uint8_t fe = forbid_eos ? 1 : 0;
    if (mgb_lt_sample(im->lt_session, hidden, temperature, top_k, &fe, nullptr, nullptr, im->seed + (im->draws++), s, a, nullptr) != MGB_OK) {
        fprintf(stderr, "magpie_local_transformer_sample_all: %s\n", mgb_last_error());
        return r;
    }
    r.sampled_codes.assign(s, s + 8);
    r.argmax_codes.assign(a, a + 8);
    return r;
}

bool magpie_is_eos(const std::vector<int32_t> & frame_codes, int32_t eos_id) {      // magpie.cpp:3273-3278
    for (int32_t code : frame_codes) if (code == eos_id) return true;
    return false;
}
bool magpie_is_eos(const int32_t * codes, int n_codebooks, int eos_id) {
    for (int i = 0; i < n_codebooks; i++) if (codes[i] == eos_id) return true;
    return false;
}

// ---- streaming --------------------------------------------------------------------------------------
std::vector<std::string> magpie_split_sentences(const char * text) {
    std::vector<std::string> out;
    if (!text || !*text) return out;
    auto flush = [&](std::string & cur) {
        const size_t a = cur.find_first_not_of(" \t\n\r");
        if (a != std::string::npos) out.push_back(cur.substr(a));
        cur.clear();
    };
    std::string cur;
    for (const char * p = text; *p; p++) {
        cur += *p;
        const char nx = p[1];
        if ((*p == '.' || *p == '!' || *p == '?') && (nx == '\0' || nx == ' ' || nx == '\n' || nx == '\t')) flush(cur);
    }
    if (!cur.empty()) flush(cur);
    return out;
}

namespace {
struct stream_bridge { const magpie_stream_params * params; int total = 0; bool stopped = false; };
// C-ABI chunk callback -> the reference's on_audio / on_progress callbacks (magpie.h:604-617)
int stream_chunk(int /*utterance*/, const float * pcm, int n_samples, int frames_done, int /*is_last*/, void * user) {
    stream_bridge * br = static_cast<stream_bridge *>(user);
    br->total += n_samples;
    if (br->params->on_audio && !br->params->on_audio(pcm, n_samples, br->params->user_data)) br->stopped = true;
    if (br->params->on_progress) br->params->on_progress(frames_done - 1, 0, 1, br->params->user_data);
    return br->stopped ? 1 : 0;
}
}  // namespace

// magpie.cpp:4502-4829.  The whole loop stays on the device: the batch-1 persistent kernel runs frames_per_chunk frames per
// launch, the codec decodes each chunk (mgb_stream_generate); nothing is synchronised per frame.
int magpie_synthesize_sentence_streaming(magpie_context * ctx, magpie_codec * codec, const int32_t * tokens, int n_tokens,
                                         const magpie_stream_params & params) {
    if (!ctx || !codec || !codec->impl || !tokens || n_tokens <= 0) return -1;
    const magpie_hparams & hp = ctx->model.hparams;
    ctx->temperature = params.temperature; ctx->top_k = params.top_k; ctx->speaker_id = params.speaker_id;
    mgb_session * s = start_utterance(ctx, tokens, n_tokens, hp.max_dec_steps, false);
    if (!s) return -1;
    magpie_model_impl * im = ctx->model.impl;
    int k = ctx->top_k;
    if (ctx->temperature >= 0.01f && k < 1) k = 1;
    stream_bridge br;
    br.params = &params;
    int32_t n_frames = 0;
    const int rc = mgb_stream_generate(s, codec->impl->codec, hp.max_dec_steps, ctx->temperature, k, im->seed + (im->draws++), 0,
                                       params.frames_per_chunk > 0 ? params.frames_per_chunk : 4, std::max(0, params.codec_context_frames),
                                       stream_chunk, &br, &n_frames);
    const bool ok = rc == MGB_OK || br.stopped;          // a callback that asks to stop ends the sentence, it is not an error
    if (!ok) fprintf(stderr, "magpie: [streaming] %s\n", mgb_last_error());
    mgb_session_free(s);
    ctx->state.n_generated_frames = n_frames;
    return ok ? br.total : -1;
}

int magpie_synthesize_streaming(magpie_context * ctx, magpie_codec * codec, const char * text, const magpie_stream_params & params) {
    if (!ctx || !codec || !text) return -1;
    std::vector<std::string> sentences;
    if (params.sentence_chunking) {
        sentences = magpie_split_sentences(text);
        if (sentences.empty()) sentences.push_back(text);
    } else sentences.push_back(text);
    int total = 0;
    for (size_t i = 0; i < sentences.size(); i++) {
        std::vector<int32_t> tokens = magpie_tokenize(&ctx->model.tokenizer, sentences[i]);
        if (tokens.empty()) { if (params.sentence_chunking) continue; return -1; }
        if (params.sentence_chunking && params.on_progress) params.on_progress(0, (int)i, (int)sentences.size(), params.user_data);
        const int n = magpie_synthesize_sentence_streaming(ctx, codec, tokens.data(), (int)tokens.size(), params);
        if (n < 0) return -1;
        total += n;
    }
    return total;
}

// ---- codec ------------------------------------------------------------------------------------------
magpie_codec * magpie_codec_init(const char * path) { return magpie_codec_init_with_backend(path, MAGPIE_BACKEND_AUTO); }

magpie_codec * magpie_codec_init_with_backend(const char * path, magpie_backend_type backend) {
    if (backend == MAGPIE_BACKEND_CPU || backend == MAGPIE_BACKEND_METAL) {
        fprintf(stderr, "magpie_codec_init: this build is CUDA (sm_100a) only\n");
        return nullptr;
    }
    if (!path) { fprintf(stderr, "magpie_codec_init: null path\n"); return nullptr; }
    mgb_codec * c = mgb_codec_load(path, env_device());
    if (!c) { fprintf(stderr, "magpie_codec_init: %s\n", mgb_last_error()); return nullptr; }
    magpie_codec * codec = new magpie_codec();
    mgb_codec_hparams hp;
    mgb_codec_get_hparams(c, &hp);
    codec->hparams.sample_rate = hp.sample_rate; codec->hparams.num_codebooks = hp.num_codebooks;
    codec->hparams.codebook_size = hp.codebook_size; codec->hparams.hop_length = hp.hop_length; codec->hparams.latent_dim = hp.latent_dim;
    codec->impl = new magpie_codec_impl();
    codec->impl->codec = c;
    return codec;
}

void magpie_codec_free(magpie_codec * codec) {
    if (!codec) return;
    if (codec->impl) { mgb_codec_free(codec->impl->codec); delete codec->impl; }
    delete codec;
}

std::vector<float> magpie_codec_decode_batch(magpie_codec * codec, const int32_t * codes, int batch, int n_frames) {
    std::vector<float> pcm;
    if (!codec || !codec->impl || !codes || n_frames <= 0 || batch <= 0) { fprintf(stderr, "magpie_codec_decode: invalid args\n"); return pcm; }
    pcm.resize((size_t)batch * n_frames * codec->hparams.hop_length);
    if (mgb_codec_decode(codec->impl->codec, codes, batch, n_frames, pcm.data()) != MGB_OK) {
        fprintf(stderr, "magpie_codec_decode: %s\n", mgb_last_error());
        pcm.clear();
    }
    return pcm;
}

std::vector<float> magpie_codec_decode(magpie_codec * codec, const int32_t * codes, int n_frames) {
    return magpie_codec_decode_batch(codec, codes, 1, n_frames);
}
