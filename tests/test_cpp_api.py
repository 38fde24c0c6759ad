"""The reference-shaped C++ API (include/magpie.h) and the magpie-tts CLI.
CPU: tokenizer / text normaliser / sentence splitter (host logic, reference magpie.cpp:127-495, 4439-4479).
GPU: greedy synthesis through the C++ API and the CLI's WAV against the oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "magpie_tts_cpp_b200")
def snr_db(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return 10 * np.log10(np.sum(b ** 2) / max(np.sum((a - b) ** 2), 1e-300))


HELLO = [2378, 7, 4, 11, 11, 14, 32, 26, 22, 14, 17, 11, 3, 32, 28, 2379]


@pytest.fixture(scope="module")
def api_test(tmp_path_factory):
    from magpie_tts_cpp_b200 import binding
    if not os.path.exists(binding.LIB_PATH):
        binding.build()
    out = str(tmp_path_factory.mktemp("cpp") / "api_test")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", out, os.path.join(ROOT, "tests", "cpp", "api_test.cpp"),
                           "-L" + PKG, "-lmagpie_b200", "-Wl,-rpath," + PKG])
    return out


@pytest.fixture(scope="module")
def vocab_files(tmp_path_factory, fx):
    d = tmp_path_factory.mktemp("vocab")
    vocab, ids = fx.synthetic_vocab()
    (d / "vocab.txt").write_text("\n".join(vocab) + "\n")
    (d / "dict.tsv").write_text("".join(f"{k}\t{v}\n" for k, v in fx.SYNTHETIC_DICT.items()))
    return str(d / "vocab.txt"), str(d / "dict.tsv"), vocab, ids


def tokenize(api_test, vocab_files, text):
    v, dct, vocab, ids = vocab_files
    out = subprocess.check_output([api_test, "tokenize", v, dct, str(ids["space"]), "2378", "2379", text], text=True)
    return [int(x) for x in out.split()]


def spell(vocab, toks):
    return "".join(vocab[t] if t < len(vocab) else {2378: "<", 2379: ">"}[t] for t in toks)


def test_tokenize_hello_world(api_test, vocab_files):
    assert tokenize(api_test, vocab_files, "Hello, world!") == HELLO


def test_tokenize_normalisation_rules(api_test, vocab_files):
    vocab = vocab_files[2]
    cases = {
        "$5": "<FIVE DOLLARS>", "$1": "<ONE DOLLAR>", "50%": "<FIFTY PERCENT>", "-3%": "<MINUS THREE PERCENT>",
        "21st": "<TWENTY FIRST>", "2nd": "<SECOND>", "13th": "<THIRTEENTH>", "20th": "<TWENTIETH>", "104th": "<ONE HUNDRED AND FOURTH>",
        "2024": "<TWENTY TWENTY FOUR>", "1900": "<NINETEEN HUNDRED>", "2005": "<TWO THOUSAND FIVE>", "3000": "<THREE THOUSAND>",
        "115": "<ONE HUNDRED AND FIFTEEN>", "-7": "<MINUS SEVEN>", "1000001": "<ONE MILLION ONE>", "0": "<ZERO>",
        "ab.cd": "<AB .CD>", "a.b": "<a.b>",      # single characters found in the vocab are emitted as-is, without a space
         "hi  there": "<HI THERE>", "end.": "<END .>", "wait ": "<WAIT>",
    }
    for text, want in cases.items():
        assert spell(vocab, tokenize(api_test, vocab_files, text)) == want, text


def test_tokenize_dictionary_longest_match(api_test, vocab_files):
    vocab = vocab_files[2]
    # dictionary words become IPA tokens (greedy longest match over <= 4 bytes: "oʊ", "tʃ", "aɪ" are single tokens)
    toks = tokenize(api_test, vocab_files, "The show")
    assert spell(vocab, toks) == "<ðə ʃoʊ>"
    assert len(toks) == 2 + 2 + 1 + 2
    assert spell(vocab, tokenize(api_test, vocab_files, "chai")) == "<tʃaɪ>"
    assert len(tokenize(api_test, vocab_files, "chai")) == 4


def test_split_sentences(api_test):
    out = subprocess.check_output([api_test, "split", "Hi there. How are you?  Fine!No split.here ok"], text=True)
    assert out.splitlines() == ["[Hi there.]", "[How are you?]", "[Fine!No split.here ok]"]
    assert subprocess.check_output([api_test, "split", ""], text=True) == ""


def test_cli_usage_errors():
    cli = os.path.join(PKG, "bin", "magpie-tts")
    if not os.path.exists(cli):
        subprocess.check_call(["make", "-C", os.path.join(PKG, "csrc"), "-s", "cli"])
    r = subprocess.run([cli], capture_output=True, text=True)
    assert r.returncode == 1 and "--text is required" in r.stderr
    r = subprocess.run([cli, "--nope"], capture_output=True, text=True)
    assert r.returncode == 1 and "Unknown option: --nope" in r.stderr
    r = subprocess.run([cli, "-t"], capture_output=True, text=True)
    assert r.returncode == 1 and "--text requires text" in r.stderr
    r = subprocess.run([cli, "-h"], capture_output=True, text=True)
    assert r.returncode == 0 and "--top-k" in r.stderr


def read_wav(path):
    b = open(path, "rb").read()
    assert b[:4] == b"RIFF" and b[8:16] == b"WAVEfmt " and b[36:40] == b"data"
    fmt = struct.unpack("<IHHIIHH", b[16:36])
    assert fmt == (16, 1, 1, 22050, 44100, 2, 16)
    n = struct.unpack("<I", b[40:44])[0]
    assert struct.unpack("<I", b[4:8])[0] == 36 + n and len(b) == 44 + n
    return np.frombuffer(b[44:], dtype="<i2")


@pytest.mark.gpu
def test_cpp_api_synthesis_matches_oracle(api_test, oracle_mod, tiny_model_path, codec_path):
    out = subprocess.check_output([api_test, "synth", tiny_model_path, codec_path, "Hello, world!", "20"], text=True,
                                  env=dict(os.environ, MAGPIE_PRECISION="f32"))
    lines = dict(l.split(":", 1) for l in out.splitlines())
    assert [int(x) for x in lines["tokens"].split()] == HELLO
    o = oracle_mod.OracleModel(tiny_model_path)
    ref = o.synthesize(HELLO, speaker=1, temperature=0.0, max_steps=20)
    codes = np.array([int(x) for x in lines["codes"].split()], np.int32).reshape(-1, 8)
    assert codes.shape == ref.shape and np.mean(np.all(codes == ref, axis=1)) >= 0.99
    h = (0.01 * ((np.arange(o.hp["d_model"]) % 17) - 8)).astype(np.float32)
    s, a, _ = o.lt_sample(h, 0.0, 80, forbid_eos=True)
    assert [int(x) for x in lines["lt"].split()] == list(a)
    enc = o.encode_text(HELLO)
    e = lines["enc"].split()
    assert int(e[0]) == 16 and int(e[1]) == 16 * o.hp["d_model"]
    assert abs(float(e[2]) - enc[0, 0]) < 2e-3 and abs(float(e[3]) - enc[-1, -1]) < 2e-3
    assert int(lines["pcm"]) == len(ref) * 1024


@pytest.mark.gpu
def test_streaming_api(api_test, oracle_mod, tiny_model_path, codec_path):
    out = subprocess.check_output([api_test, "stream", tiny_model_path, codec_path, "Hello. World!", "12", "4"], text=True,
                                  env=dict(os.environ, MAGPIE_PRECISION="f32"))
    lines = dict(l.split(":", 1) for l in out.splitlines() if ":" in l)
    chunks = [int(x) for x in lines["chunks"].split()]
    assert int(lines["total"]) == sum(chunks) and all(c % 1024 == 0 and 0 < c <= 4 * 1024 for c in chunks)
    # two sentences, each at least min 4 frames; chunks are flushed per sentence
    assert sum(chunks) >= 2 * 4 * 1024


@pytest.mark.gpu
def test_streaming_with_codec_context_is_seamless(api_test, oracle_mod, tiny_model_path, codec_path, tmp_path):
    """codec_context_frames = 25 (>= the codec's 24.8-frame receptive field): the chunks streamed every 4 frames concatenate to
    the whole-utterance decode of the same codes; with 0 (the reference's zero-history chunks) they do not."""
    env = dict(os.environ, MAGPIE_PRECISION="f32")
    o = oracle_mod.OracleModel(tiny_model_path)
    codes = o.synthesize(HELLO, speaker=1, temperature=0.0, max_steps=12)     # greedy: the streamed codes are the same
    audio = {}
    for ctx_frames in (0, 25):
        f = str(tmp_path / f"stream{ctx_frames}.f32")
        subprocess.check_output([api_test, "stream", tiny_model_path, codec_path, "Hello, world!", "12", "4", str(ctx_frames), f], text=True, env=env)
        audio[ctx_frames] = np.fromfile(f, np.float32)
    n = len(codes) * 1024
    assert len(audio[0]) == len(audio[25]) >= n
    whole = oracle_mod.OracleCodec(codec_path).decode(np.ascontiguousarray(codes.T))
    assert snr_db(audio[25][:n], whole) >= 40.0
    assert snr_db(audio[0][4 * 1024:n], whole[4 * 1024:]) < 20.0      # zero-history chunks differ from the continuous decode


@pytest.mark.gpu
def test_cli_wav_matches_oracle(tmp_path, oracle_mod, tiny_model_path, codec_path):
    cli = os.path.join(PKG, "bin", "magpie-tts")
    if not os.path.exists(cli):
        subprocess.check_call(["make", "-C", os.path.join(PKG, "csrc"), "-s", "cli"])
    wav = str(tmp_path / "hello.wav")
    r = subprocess.run([cli, "-m", tiny_model_path, "-c", codec_path, "-t", "Hello, world!", "--temp", "0", "-s", "1", "-o", wav, "-q"],
                       capture_output=True, text=True, env=dict(os.environ, MAGPIE_PRECISION="f32"))
    assert r.returncode == 0, r.stderr
    assert r.stdout.strip() == wav                      # -q prints only the output path on stdout
    pcm16 = read_wav(wav)
    o = oracle_mod.OracleModel(tiny_model_path)
    codes = o.synthesize(HELLO, speaker=1, temperature=0.0)     # max_dec_steps of the tiny fixture = 24
    oc = oracle_mod.OracleCodec(codec_path)
    ref = np.concatenate([oc.decode(np.ascontiguousarray(codes[i:i + 32].T)) for i in range(0, len(codes), 32)])
    ref16 = (np.clip(ref, -1, 1) * np.float32(32767.0)).astype(np.int16)      # C truncation toward zero
    assert pcm16.shape == ref16.shape
    err = pcm16.astype(np.int32) - ref16.astype(np.int32)
    snr = 10 * np.log10(np.sum(ref16.astype(np.float64) ** 2) / max(np.sum(err.astype(np.float64) ** 2), 1e-9))
    assert snr >= 40.0 and np.abs(err).max() <= 40
