// Test driver for include/magpie.h (the reference-shaped C++ API).  Built by tests/test_cpp_api.py.
//   api_test tokenize <vocab.txt> <dict.tsv> <space_id> <bos> <eos> <text>      (CPU only)
//   api_test split <text>                                                      (CPU only)
//   api_test synth <model.gguf> <codec.gguf> <text> <max_steps>                (GPU) greedy codes + LT sample + encode
//   api_test stream <model.gguf> <codec.gguf> <text> <max_steps> <frames_per_chunk> [codec_context_frames [dump.f32]]   (GPU)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "../../include/magpie.h"

static std::vector<std::string> read_lines(const char * path) {
    std::vector<std::string> out;
    std::ifstream f(path);
    for (std::string l; std::getline(f, l);) out.push_back(l);
    return out;
}

struct StreamLog { std::vector<int> chunk_samples; std::vector<float> audio; int progress_calls = 0; int last_frames = -1; };
static bool on_audio(const float * s, int n, void * u) {
    StreamLog * l = (StreamLog *)u;
    l->chunk_samples.push_back(n);
    l->audio.insert(l->audio.end(), s, s + n);
    return true;
}
static void on_progress(int frames, int, int, void * u) { StreamLog * l = (StreamLog *)u; l->progress_calls++; l->last_frames = frames; }

int main(int argc, char ** argv) {
    if (argc < 2) return 2;
    const std::string cmd = argv[1];
    if (cmd == "tokenize" && argc == 8) {
        magpie_tokenizer tok;
        tok.vocab = read_lines(argv[2]);
        for (size_t i = 0; i < tok.vocab.size(); i++) tok.token_to_id[tok.vocab[i]] = (int32_t)i;
        for (const std::string & l : read_lines(argv[3])) {
            const size_t t = l.find('\t');
            if (t != std::string::npos) tok.dict[l.substr(0, t)] = l.substr(t + 1);
        }
        tok.space_id = atoi(argv[4]); tok.bos_id = atoi(argv[5]); tok.eos_id = atoi(argv[6]); tok.loaded = true;
        for (int32_t id : magpie_tokenize(&tok, argv[7])) printf("%d ", id);
        printf("\n");
        magpie_tokenizer unloaded;
        return magpie_tokenize(&unloaded, "x").empty() ? 0 : 1;
    }
    if (cmd == "split" && argc == 3) {
        for (const std::string & s : magpie_split_sentences(argv[2])) printf("[%s]\n", s.c_str());
        return 0;
    }
    if ((cmd == "synth" && argc == 6) || (cmd == "stream" && argc >= 7 && argc <= 9)) {
        magpie_context * ctx = magpie_init(argv[2]);
        if (!ctx) return 3;
        if (magpie_init_with_backend(argv[2], MAGPIE_BACKEND_CPU) != nullptr) return 4;      // no CPU fallback
        magpie_codec * codec = magpie_codec_init(argv[3]);
        if (!codec) return 5;
        ctx->model.hparams.max_dec_steps = atoi(argv[5]);
        ctx->temperature = 0.0f; ctx->top_k = 1; ctx->speaker_id = 1;
        std::vector<int32_t> tokens = magpie_tokenize(&ctx->model.tokenizer, argv[4]);
        printf("tokens:"); for (int32_t t : tokens) printf(" %d", t); printf("\n");
        if (cmd == "synth") {
            if (!magpie_encode_text(ctx, tokens.data(), (int)tokens.size())) return 6;
            printf("enc: %d %zu %.9g %.9g\n", ctx->state.enc_seq_len, ctx->state.encoder_output.size(),
                   ctx->state.encoder_output[0], ctx->state.encoder_output.back());
            std::vector<int32_t> codes = magpie_synthesize_codes_graph_reuse(ctx, tokens.data(), (int)tokens.size());
            std::vector<int32_t> codes2 = magpie_synthesize_codes(ctx, tokens.data(), (int)tokens.size());
            if (codes != codes2) return 7;
            printf("codes:"); for (int32_t c : codes) printf(" %d", c); printf("\n");
            std::vector<float> h(ctx->model.hparams.d_model);
            for (size_t i = 0; i < h.size(); i++) h[i] = 0.01f * (float)((int)(i % 17) - 8);
            magpie_sample_result r = magpie_local_transformer_sample_all(ctx, h.data(), 0.0f, 80, true);
            if (r.sampled_codes.size() != 8 || r.sampled_codes != r.argmax_codes) return 8;
            printf("lt:"); for (int32_t c : r.sampled_codes) printf(" %d", c); printf("\n");
            if (!codes.empty()) {
                const int T = (int)codes.size() / 8;
                std::vector<int32_t> cbm(codes.size());
                for (int t = 0; t < T; t++) for (int cb = 0; cb < 8; cb++) cbm[cb * T + t] = codes[t * 8 + cb];
                std::vector<float> pcm = magpie_codec_decode(codec, cbm.data(), T);
                printf("pcm: %zu\n", pcm.size());
                if ((int)pcm.size() != T * codec->hparams.hop_length) return 9;
            }
            if (!magpie_synthesize_codes(ctx, nullptr, 0).empty()) return 10;                 // error path: empty vector
            if (!magpie_codec_decode(codec, nullptr, 0).empty()) return 11;
        } else {
            StreamLog log;
            magpie_stream_params sp;
            sp.temperature = 0.0f; sp.top_k = 1; sp.speaker_id = 1; sp.frames_per_chunk = atoi(argv[6]);
            sp.on_audio = on_audio; sp.on_progress = on_progress; sp.user_data = &log;
            if (argc > 7) sp.codec_context_frames = atoi(argv[7]);
            if (argc > 8) sp.sentence_chunking = false;
            const int total = magpie_synthesize_streaming(ctx, codec, argv[4], sp);
            if (argc > 8) { FILE * f = fopen(argv[8], "wb"); if (f) { fwrite(log.audio.data(), 4, log.audio.size(), f); fclose(f); } }
            printf("total: %d\nchunks:", total);
            for (int n : log.chunk_samples) printf(" %d", n);
            printf("\nprogress: %d %d\n", log.progress_calls, log.last_frames);
            if (total != (int)log.audio.size()) return 12;
            if (magpie_synthesize_streaming(ctx, codec, nullptr, sp) != -1) return 13;
        }
        magpie_codec_free(codec);
        magpie_free(ctx);
        return 0;
    }
    return 2;
}
