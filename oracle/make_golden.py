"""TEST INFRASTRUCTURE.  Generates tests/golden/ref_pieces.json by RUNNING THE REFERENCE'S OWN CODE (the self-contained
functions compiled by oracle/build_ref_pieces.py from /root/reference) on fixed inputs.  Run in the build container (the
GPU box has no /root/reference); the JSON is committed and is what the tests read.

    python oracle/make_golden.py
"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import build_ref_pieces  # noqa: E402
from magpie_tts_cpp_b200 import fixtures  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "ref_pieces.json")

TOKENIZE_TEXTS = [
    "Hello, world!", "$5", "$1", "50%", "-3%", "21st", "2nd", "13th", "20th", "104th", "2024", "1900", "2005", "3000", "115", "-7",
    "1000001", "0", "ab.cd", "a.b", "hi  there", "end.", "wait ", "The show", "chai", "It costs $1,234 in 1999.", "Dr. Smith's 3rd try: 42%!",
    "", "   ", "?!", "the quick brown fox", "12345678901", "1st 2nd 3rd 4th 11th 12th 101st", "x-ray 7-11", "ünïcödé",
]
SENTENCE_TEXTS = ["", "One.", "One. Two! Three?", "No boundary", "A.B. C.", "  lead. trail  ", "Dots... here. End", "Mr. X went. \nNew line! Tab?\tOk"]


def main():
    lib = C.CDLL(build_ref_pieces.build(force=True))
    g = {"source": "m1el/magpie-tts.cpp functions compiled from /root/reference by oracle/build_ref_pieces.py"}

    # ---- FSQ (nano-codec.cpp:721-752): every index of every codebook incl. the out-of-range ids 2016..2023 and negatives
    idx = np.concatenate([np.arange(2024, dtype=np.int32), np.array([-1, -7, -2016, 4095, 100000], np.int32)])
    T = len(idx)
    codes = np.stack([np.roll(idx, 13 * cb) for cb in range(8)]).astype(np.int32)
    lat = np.zeros((32, T), np.float32)
    lib.ref_fsq_dequantize(codes.ctypes.data_as(C.c_void_p), lat.ctypes.data_as(C.c_void_p), 8, T)
    g["fsq"] = {"indices": idx.tolist(), "roll_per_codebook": 13,
                "latent_bits_cb0": lat[:4].view(np.uint32).tolist(),                  # 4 dims of codebook 0: the whole function table
                "latent_bits_checksum": int(np.bitwise_xor.reduce(lat.view(np.uint32).ravel() * np.arange(1, lat.size + 1, dtype=np.uint32)))}

    # ---- sample_top_k (magpie.cpp:1072-1109) with the reference's own mt19937 draw
    rng = np.random.default_rng(2024)
    cases = []
    lib.ref_sample_top_k.restype = C.c_int
    for i, (n, k, temp) in enumerate([(2024, 80, 0.7), (2024, 80, 1.0), (2024, 1, 0.7), (2024, 2024, 0.5), (2024, 5000, 1.3), (64, 8, 0.2),
                                      (2024, 80, 0.05), (2024, 40, 0.7), (2024, 80, 0.7), (2024, 80, 0.7)]):
        logits = (rng.standard_normal(n) * (0.3 + 0.4 * (i % 4))).astype(np.float32)
        if i == 8:
            logits[[5, 900, 901]] = logits.max() + 1.0        # exact ties at the top
        if i == 9:
            logits[[2016, 2018, 2019, 2020, 2021, 2022, 2023]] = -np.inf      # the always-forbidden ids
        for seed in (1, 7, 12345):
            u = C.c_float()
            pick = lib.ref_sample_top_k(logits.ctypes.data_as(C.c_void_p), n, C.c_float(temp), k, seed, C.byref(u))
            cases.append({"case": i, "n": n, "top_k": k, "temperature": temp, "seed": seed, "u_bits": int(np.float32(u.value).view(np.uint32)),
                          "pick": int(pick)})
        g.setdefault("sampler_logits_bits", {})[str(i)] = logits.view(np.uint32).tolist()
    g["sampler"] = cases

    # ---- tokenizer + normaliser (magpie.cpp:128-495) on the synthetic vocabulary of the fixtures
    vocab, ids = fixtures.synthetic_vocab()
    lib.ref_tok_reset(ids["space"], 2378, 2379, ids.get("oov", -1))
    for v in vocab:
        lib.ref_tok_add_vocab(v.encode("utf-8"))
    for w, ipa in fixtures.SYNTHETIC_DICT.items():
        lib.ref_tok_add_dict(w.encode("utf-8"), ipa.encode("utf-8"))
    buf = (C.c_int32 * 4096)()
    sbuf = C.create_string_buffer(16384)
    toks, norms = {}, {}
    for t in TOKENIZE_TEXTS:
        n = lib.ref_tokenize(t.encode("utf-8"), buf, 4096)
        toks[t] = [int(buf[i]) for i in range(n)]
        lib.ref_normalize(t.encode("utf-8"), sbuf, 16384)
        norms[t] = sbuf.value.decode("utf-8", errors="replace")
    g["tokenize"] = toks
    g["normalize"] = norms
    g["tokenizer_ids"] = {"space": ids["space"], "bos": 2378, "eos": 2379, "oov": ids.get("oov", -1)}

    # ---- sentence splitter (magpie.cpp:4439-4480)
    sents = {}
    for t in SENTENCE_TEXTS:
        lib.ref_split_sentences(t.encode("utf-8"), sbuf, 16384)
        sents[t] = [x for x in sbuf.value.decode("utf-8").split("\x1f") if x != ""]
    g["split_sentences"] = sents

    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w", encoding="utf-8") as f:
        json.dump(g, f, ensure_ascii=False, separators=(",", ":"))
    print(OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
