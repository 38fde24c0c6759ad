// magpie-tts command line front end over include/magpie.h.
// Same flags, defaults, messages on error and output format as the reference CLI (src/magpie-tts.cpp:70-226):
// -m/--model -c/--codec -t/--text -o/--output -s/--speaker --temp --top-k -q/--quiet -h/--help; mono 16-bit PCM WAV
// at 22050 Hz; the codec is run on independent 32-frame chunks (magpie-tts.cpp:181-206) -- here all full chunks of
// an utterance go to the GPU as one batch.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/magpie.h"

namespace {

struct Options {
    std::string model = "weights/magpie-357m-f32.gguf";
    std::string codec = "weights/nano-codec-f32.gguf";
    std::string text;
    std::string output = "output.wav";
    int speaker = 0;
    float temp = 0.7f;
    int top_k = 80;
    bool quiet = false, have_text = false;
};

void usage(const char * prog) {
    fprintf(stderr,
            "Magpie TTS - Text-to-Speech (B200-native build)\n\n"
            "Usage: %s [options]\n\n"
            "Options:\n"
            "  -m, --model PATH     Path to model GGUF (default: weights/magpie-357m-f32.gguf)\n"
            "  -c, --codec PATH     Path to codec GGUF (default: weights/nano-codec-f32.gguf)\n"
            "  -t, --text TEXT      Text to synthesize (required)\n"
            "  -o, --output PATH    Output WAV file (default: output.wav)\n"
            "  -s, --speaker ID     Speaker ID (default: 0)\n"
            "  --temp FLOAT         Sampling temperature (default: 0.7, 0=deterministic)\n"
            "  --top-k INT          Top-k sampling (default: 80)\n"
            "  -q, --quiet          Minimal output\n"
            "  -h, --help           Show this help\n"
            "\nEnvironment: MAGPIE_PRECISION=f32|bf16 (default f32), MAGPIE_DEVICE=<cuda index>, MAGPIE_SEED=<n>\n",
            prog);
}

// returns 0 = run, 1 = exit with error, 2 = exit ok (help)
int parse(int argc, char ** argv, Options & o) {
    struct Flag { const char * s; const char * l; const char * what; int id; };
    static const Flag flags[] = {
        {"-m", "--model", "--model requires a path", 0}, {"-c", "--codec", "--codec requires a path", 1},
        {"-t", "--text", "--text requires text", 2},     {"-o", "--output", "--output requires a path", 3},
        {"-s", "--speaker", "--speaker requires an ID", 4}, {nullptr, "--temp", "--temp requires a value", 5},
        {nullptr, "--top-k", "--top-k requires a value", 6},
    };
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        if (a == "-h" || a == "--help") { usage(argv[0]); return 2; }
        if (a == "-q" || a == "--quiet") { o.quiet = true; continue; }
        const Flag * f = nullptr;
        for (const Flag & c : flags) if ((c.s && a == c.s) || a == c.l) f = &c;
        if (!f) { fprintf(stderr, "Unknown option: %s\n", argv[i]); usage(argv[0]); return 1; }
        if (++i >= argc) { fprintf(stderr, "Error: %s\n", f->what); return 1; }
        switch (f->id) {
            case 0: o.model = argv[i]; break;
            case 1: o.codec = argv[i]; break;
            case 2: o.text = argv[i]; o.have_text = true; break;
            case 3: o.output = argv[i]; break;
            case 4: o.speaker = atoi(argv[i]); break;
            case 5: o.temp = (float)atof(argv[i]); break;
            case 6: o.top_k = atoi(argv[i]); break;
        }
    }
    if (!o.have_text) { fprintf(stderr, "Error: --text is required\n\n"); usage(argv[0]); return 1; }
    return 0;
}

// 44-byte RIFF header + int16(clamp(x, -1, 1) * 32767) truncation, as the reference writer (magpie-tts.cpp:30-68)
bool write_wav16(const std::string & path, const std::vector<float> & pcm, int32_t rate) {
    FILE * f = fopen(path.c_str(), "wb");
    if (!f) return false;
    std::vector<int16_t> s(pcm.size());
    for (size_t i = 0; i < pcm.size(); i++) {
        float v = pcm[i] > 1.0f ? 1.0f : (pcm[i] < -1.0f ? -1.0f : pcm[i]);
        s[i] = static_cast<int16_t>(v * 32767.0f);
    }
    const int32_t data_bytes = (int32_t)(s.size() * sizeof(int16_t));
    unsigned char h[44];
    auto put32 = [&](int off, int32_t v) { memcpy(h + off, &v, 4); };
    auto put16 = [&](int off, int16_t v) { memcpy(h + off, &v, 2); };
    memcpy(h, "RIFF", 4); put32(4, 36 + data_bytes); memcpy(h + 8, "WAVEfmt ", 8);
    put32(16, 16); put16(20, 1); put16(22, 1); put32(24, rate); put32(28, rate * 2); put16(32, 2); put16(34, 16);
    memcpy(h + 36, "data", 4); put32(40, data_bytes);
    bool ok = fwrite(h, 1, 44, f) == 44 && (s.empty() || fwrite(s.data(), 2, s.size(), f) == s.size());
    ok = fclose(f) == 0 && ok;
    return ok;
}

}  // namespace

int main(int argc, char ** argv) {
    Options o;
    const int pr = parse(argc, argv, o);
    if (pr) return pr == 2 ? 0 : 1;
    if (!o.quiet)
        fprintf(stderr, "Magpie TTS\n  Model: %s\n  Codec: %s\n  Text: \"%s\"\n  Output: %s\n  Speaker: %d\n  Temperature: %.2f\n  Top-k: %d\n\n",
                o.model.c_str(), o.codec.c_str(), o.text.c_str(), o.output.c_str(), o.speaker, o.temp, o.top_k);

    if (!o.quiet) fprintf(stderr, "Loading model...\n");
    magpie_context * ctx = magpie_init(o.model.c_str());
    if (!ctx) { fprintf(stderr, "Error: Failed to load model from %s\n", o.model.c_str()); return 1; }
    if (!o.quiet) fprintf(stderr, "Backend: %s\n\n", magpie_get_backend_name(ctx));
    ctx->temperature = o.temp; ctx->top_k = o.top_k; ctx->speaker_id = o.speaker;

    if (!o.quiet) fprintf(stderr, "Tokenizing...\n");
    const std::vector<int32_t> tokens = magpie_tokenize(&ctx->model.tokenizer, o.text);
    if (tokens.empty()) { fprintf(stderr, "Error: Tokenization failed\n"); magpie_free(ctx); return 1; }
    if (!o.quiet) fprintf(stderr, "Tokens: %zu\n\nSynthesizing...\n", tokens.size());

    const std::vector<int32_t> codes = magpie_synthesize_codes_graph_reuse(ctx, tokens.data(), (int)tokens.size());
    if (codes.empty()) { fprintf(stderr, "Error: Synthesis failed\n"); magpie_free(ctx); return 1; }
    const int n_frames = (int)codes.size() / 8;
    if (!o.quiet) fprintf(stderr, "Generated %d frames\n\nLoading codec...\n", n_frames);

    magpie_codec * codec = magpie_codec_init(o.codec.c_str());
    if (!codec) { fprintf(stderr, "Error: Failed to load codec from %s\n", o.codec.c_str()); magpie_free(ctx); return 1; }

    if (!o.quiet) fprintf(stderr, "Decoding audio...\n");
    const int kChunk = 32;                       // independent 32-frame chunks, zero causal history each
    const int full = n_frames / kChunk, tail = n_frames % kChunk;
    std::vector<float> audio;
    auto gather = [&](int first_frame, int nchunks, int len) {          // [frames][8] -> [chunk][8][len]
        std::vector<int32_t> cbm((size_t)nchunks * 8 * len);
        for (int c = 0; c < nchunks; c++)
            for (int t = 0; t < len; t++)
                for (int cb = 0; cb < 8; cb++)
                    cbm[((size_t)c * 8 + cb) * len + t] = codes[((size_t)first_frame + (size_t)c * len + t) * 8 + cb];
        return cbm;
    };
    bool ok = true;
    if (full > 0) {
        const std::vector<float> a = magpie_codec_decode_batch(codec, gather(0, full, kChunk).data(), full, kChunk);
        ok = !a.empty();
        audio.insert(audio.end(), a.begin(), a.end());
    }
    if (ok && tail > 0) {
        const std::vector<float> a = magpie_codec_decode(codec, gather(full * kChunk, 1, tail).data(), tail);
        ok = !a.empty();
        audio.insert(audio.end(), a.begin(), a.end());
    }
    if (!ok) { fprintf(stderr, "Error: Codec decode failed\n"); magpie_codec_free(codec); magpie_free(ctx); return 1; }

    if (!o.quiet) fprintf(stderr, "Writing %s...\n", o.output.c_str());
    if (!write_wav16(o.output, audio, 22050)) {
        fprintf(stderr, "Error: Failed to write %s\n", o.output.c_str());
        magpie_codec_free(codec); magpie_free(ctx);
        return 1;
    }
    if (!o.quiet) fprintf(stderr, "\nDone! Generated %.2f seconds of audio.\n", (float)audio.size() / 22050.0f);
    else printf("%s\n", o.output.c_str());
    magpie_codec_free(codec);
    magpie_free(ctx);
    return 0;
}
