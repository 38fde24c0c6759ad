"""TEST INFRASTRUCTURE (checker side; nothing under magpie_tts_cpp_b200/ imports this).

Builds oracle/_ref/libref_pieces.so from the REFERENCE'S OWN SOURCES, where they lie under /root/reference.

The reference as a whole cannot be built here (it links ggml, which is neither vendored nor installed),
but a few functions on the hot path are self-contained C++ (std:: only).  This recipe extracts exactly those function
definitions from the reference files at build time (by name + brace matching; nothing is copied into the repository,
the generated translation unit lives in the git-ignored oracle/_ref/), adds extern "C" shims and compiles them with
g++.  oracle/make_golden.py runs the resulting library to produce the golden vectors committed under tests/golden/.

  src/nano-codec.cpp   fsq_dequantize_cpu                                   (:721-752)
  src/magpie.cpp       sample_top_k                                         (:1072-1109)
  src/magpie.cpp       split_string .. normalize_text, magpie_tokenize      (:128-349, 404-495)
  src/magpie.cpp       magpie_split_sentences                               (:4439-4480)
  src/magpie.h         struct magpie_tokenizer                              (:86-104)
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("MAGPIE_REFERENCE", "/root/reference")
OUT_DIR = os.path.join(HERE, "_ref")
LIB = os.path.join(OUT_DIR, "libref_pieces.so")


def _extract(src: str, signature_regex: str) -> str:
    """Text of the definition whose first line matches `signature_regex` (up to the matching closing brace)."""
    m = re.search(signature_regex, src, re.M)
    if not m:
        raise RuntimeError(f"reference function not found: {signature_regex}")
    i = src.index("{", m.start())
    depth, j = 0, i
    while True:
        c = src[j]
        if c == "{":
            depth += 1
        elif c == "}":
            depth -= 1
            if depth == 0:
                break
        j += 1
    end = j + 1
    if src[end:end + 1] == ";":
        end += 1
    return src[m.start():end] + "\n"


SHIMS = r'''
extern "C" {
void ref_fsq_dequantize(const int32_t * codes, float * latent, int num_cb, int n_frames) { fsq_dequantize_cpu(codes, latent, num_cb, n_frames); }

// draws u exactly as sample_top_k will (from a copy of the generator), then samples; returns the pick, *u_out = the draw
int ref_sample_top_k(const float * logits, int n, float temperature, int top_k, unsigned seed, float * u_out) {
    std::mt19937 rng(seed);
    std::mt19937 copy = rng;
    std::uniform_real_distribution<float> dist(0.0f, 1.0f);
    *u_out = dist(copy);
    std::vector<float> v(logits, logits + n);
    return sample_top_k(v, temperature, top_k, rng);
}

static magpie_tokenizer g_tok;
void ref_tok_reset(int space_id, int bos_id, int eos_id, int oov_id) {
    g_tok = magpie_tokenizer();
    g_tok.space_id = space_id; g_tok.bos_id = bos_id; g_tok.eos_id = eos_id; g_tok.oov_id = oov_id; g_tok.loaded = true;
}
void ref_tok_add_vocab(const char * s) { g_tok.token_to_id[s] = (int32_t)g_tok.vocab.size(); g_tok.vocab.push_back(s); }
void ref_tok_add_dict(const char * w, const char * ipa) { g_tok.dict[w] = ipa; }
int ref_tokenize(const char * text, int32_t * out, int cap) {
    std::vector<int32_t> t = magpie_tokenize(&g_tok, text);
    for (size_t i = 0; i < t.size() && (int)i < cap; i++) out[i] = t[i];
    return (int)t.size();
}
int ref_normalize(const char * text, char * out, int cap) {
    std::string s = normalize_text(text);
    snprintf(out, cap, "%s", s.c_str());
    return (int)s.size();
}
int ref_split_sentences(const char * text, char * out, int cap) {   // sentences joined by '\x1f'
    std::string j;
    for (const std::string & s : magpie_split_sentences(text)) { j += s; j += '\x1f'; }
    snprintf(out, cap, "%s", j.c_str());
    return (int)j.size();
}
}
'''


def build(force: bool = False) -> str:
    """Returns the library path; raises if the reference sources are not available (e.g. on the GPU box)."""
    if os.path.exists(LIB) and not force:
        return LIB
    cpp = open(os.path.join(REF, "src", "magpie.cpp"), encoding="utf-8").read()
    hdr = open(os.path.join(REF, "src", "magpie.h"), encoding="utf-8").read()
    codec = open(os.path.join(REF, "src", "nano-codec.cpp"), encoding="utf-8").read()
    parts = [
        "// GENERATED at build time from the reference sources by oracle/build_ref_pieces.py -- do not commit\n",
        "#include <algorithm>\n#include <cmath>\n#include <cstdint>\n#include <cstdio>\n#include <cstring>\n#include <map>\n"
        "#include <random>\n#include <string>\n#include <vector>\n",
        _extract(hdr, r"^struct magpie_tokenizer \{"),
        _extract(codec, r"^static void fsq_dequantize_cpu\("),
        _extract(cpp, r"^static int32_t sample_top_k\("),
        _extract(cpp, r"^static std::vector<std::string> split_string\("),
        _extract(cpp, r"^static std::string to_lower\("),
        _extract(cpp, r"^static std::string number_to_words\("),
        _extract(cpp, r"^static std::string year_to_words\("),
        _extract(cpp, r"^static std::string ordinal_to_words\("),
        _extract(cpp, r"^static std::string normalize_text\("),
        _extract(cpp, r"^std::vector<int32_t> magpie_tokenize\("),
        _extract(cpp, r"^std::vector<std::string> magpie_split_sentences\("),
        SHIMS,
    ]
    # number_to_words has a default argument in its definition and is used before... keep the reference order
    os.makedirs(OUT_DIR, exist_ok=True)
    gen = os.path.join(OUT_DIR, "ref_pieces.cpp")
    with open(gen, "w", encoding="utf-8") as f:
        f.write("".join(parts))
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-w", "-o", LIB, gen])
    return LIB


def available() -> bool:
    return os.path.exists(os.path.join(REF, "src", "magpie.cpp"))


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
