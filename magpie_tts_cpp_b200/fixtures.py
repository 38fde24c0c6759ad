"""Random-init GGUF fixtures with the reference's tensor names, layout and KV keys.

The real Magpie-357M / nano-codec checkpoints are not available offline, so parity and
benchmarks run on seeded random-init files that follow the L0 contract exactly:

* tensor names / shapes: reference loader `src/magpie.cpp:607-667`, `src/nano-codec.cpp:84-199`
* layout: ggml ``ne`` = reversed PyTorch shape, row-major data
  (`scripts/convert_magpie_to_gguf.py:301-304`, `scripts/convert_codec_to_gguf.py`)
* KV keys: the spellings the C++ reader honours (`src/magpie.cpp:85-120`), typed UINT32 / FLOAT32
* tokenizer keys: `magpie.tokenizer.{vocab,dict,pad,oov,space}` (`src/magpie.cpp:353-398`)
* Q8_0 only on the converter's pattern set with inner dim >= 32
  (`scripts/convert_magpie_to_gguf.py:156-178, 308-320`)

Distributions: Linear/conv U(+-1/sqrt(fan_in)), LayerNorm 1+0.1 N(0,1),
embeddings and baked context N(0,1), position embeddings 0.02 N(0,1), snake alpha U(0.5,1.5).
"""
from __future__ import annotations

import os
import re
from dataclasses import dataclass, field, asdict

import numpy as np

import gguf
from gguf import GGUFWriter, GGMLQuantizationType


@dataclass
class MagpieConfig:
    d_model: int = 768
    d_ffn: int = 3072
    d_head: int = 64
    enc_layers: int = 6
    enc_heads: int = 12
    enc_kernel: int = 3
    dec_layers: int = 12
    dec_sa_heads: int = 12
    dec_xa_heads: int = 1
    dec_xa_d_head: int = 128
    dec_kernel: int = 1
    lt_dim: int = 256
    lt_ffn_dim: int = 1024
    lt_layers: int = 1
    lt_heads: int = 1
    text_vocab_size: int = 2380
    num_codebooks: int = 8
    codebook_size: int = 2016
    vocab_per_cb: int = 2024
    num_speakers: int = 5
    context_frames: int = 110
    text_bos_id: int = 2378
    text_eos_id: int = 2379
    audio_bos_id: int = 2016
    audio_eos_id: int = 2017
    max_dec_steps: int = 500
    sample_rate: int = 22050
    eps: float = 1e-5
    # not GGUF keys: table lengths of the fixture
    enc_pos_rows: int = 4096
    dec_pos_rows: int = 2816
    lt_pos_rows: int = 10


def full_config(**kw) -> MagpieConfig:
    return MagpieConfig(**kw)


def tiny_config(**kw) -> MagpieConfig:
    """Small architecture with the same structure, for fast CPU cross-checks."""
    base = dict(d_model=128, d_ffn=256, d_head=64, enc_layers=2, enc_heads=2, dec_layers=2,
                dec_sa_heads=2, dec_xa_heads=1, dec_xa_d_head=128, lt_dim=64, lt_ffn_dim=128,
                num_speakers=2, context_frames=10, max_dec_steps=24, enc_pos_rows=128,
                dec_pos_rows=128)
    base.update(kw)
    return MagpieConfig(**base)


_HP_KEYS = ["d_model", "d_ffn", "d_head", "enc_layers", "enc_heads", "enc_kernel", "dec_layers",
            "dec_sa_heads", "dec_xa_heads", "dec_xa_d_head", "dec_kernel", "lt_dim", "lt_ffn_dim",
            "lt_layers", "lt_heads", "text_vocab_size", "num_codebooks", "codebook_size",
            "vocab_per_cb", "num_speakers", "context_frames", "text_bos_id", "text_eos_id",
            "audio_bos_id", "audio_eos_id", "max_dec_steps", "sample_rate"]

# Synthetic tokenizer: A-Z, punctuation, space, pad, oov, then IPA-ish symbols.
_PUNCT = [",", ".", "!", "?", ":", ";"]
_IPA = list("abcdefghijklmnopqrstuvwxyz") + ["ˈ", "ə", "ɪ", "ŋ", "ʃ", "θ", "ð", "oʊ", "aɪ", "tʃ"]


def synthetic_vocab():
    vocab = [chr(ord("A") + i) for i in range(26)] + _PUNCT + [" ", "<pad>", "<oov>"] + _IPA
    ids = {"space": 32, "pad": 33, "oov": 34}
    assert vocab[ids["space"]] == " "
    return vocab, ids


SYNTHETIC_DICT = {
    "the": "ðə",
    "quick": "kwɪk",
    "fox": "fɑks".replace("ɑ", "a"),
    "thing": "θɪŋ",
    "show": "ʃoʊ",
    "chai": "tʃaɪ",
}

_Q_PATTERNS = [
    r"\.layers\.\d+\.self_attention\.(qkv_net|o_net)\.weight$",
    r"\.layers\.\d+\.cross_attention\.(q_net|kv_net|o_net)\.weight$",
    r"\.layers\.\d+\.pos_ff\.(proj|o_net)\.conv\.weight$",
    r"^final_proj\.weight$",
    r"^local_transformer_out_projections\.\d+\.weight$",
    r"^local_transformer_in_projection\.weight$",
]


def _should_quantize(name: str) -> bool:
    return any(re.search(p, name) for p in _Q_PATTERNS)


class _Gen:
    def __init__(self, seed):
        self.rng = np.random.default_rng(seed)

    def uniform(self, shape, fan_in):
        b = 1.0 / np.sqrt(fan_in)
        a = self.rng.random(shape, dtype=np.float32)
        a *= np.float32(2 * b)
        a -= np.float32(b)
        return a

    def normal(self, shape, scale=1.0, loc=0.0):
        a = self.rng.standard_normal(shape, dtype=np.float32)
        if scale != 1.0:
            a *= np.float32(scale)
        if loc != 0.0:
            a += np.float32(loc)
        return a

    def ln(self, n):
        return self.normal((n,), 0.1, 1.0)


def magpie_tensors(cfg: MagpieConfig, seed: int = 1234) -> "dict[str, np.ndarray]":
    """All model tensors in PyTorch shapes (float32), keyed by GGUF name, sorted order."""
    g = _Gen(seed)
    d, f = cfg.d_model, cfg.d_ffn
    t: dict[str, np.ndarray] = {}
    t["text_embedding.weight"] = g.normal((cfg.text_vocab_size, d))
    for cb in range(cfg.num_codebooks):
        t[f"audio_embeddings.{cb}.weight"] = g.normal((cfg.vocab_per_cb, d))
    t["baked_context_embedding.weight"] = g.normal((cfg.num_speakers, cfg.context_frames * d))
    t["encoder.position_embeddings.weight"] = g.normal((cfg.enc_pos_rows, d), 0.02)
    ke = cfg.enc_kernel
    for l in range(cfg.enc_layers):
        p = f"encoder.layers.{l}."
        t[p + "norm_self.weight"] = g.ln(d)
        t[p + "self_attention.qkv_net.weight"] = g.uniform((3 * d, d), d)
        t[p + "self_attention.o_net.weight"] = g.uniform((d, d), d)
        t[p + "norm_pos_ff.weight"] = g.ln(d)
        t[p + "pos_ff.proj.conv.weight"] = g.uniform((f, d, ke), d * ke)
        t[p + "pos_ff.o_net.conv.weight"] = g.uniform((d, f, ke), f * ke)
    t["encoder.norm_out.weight"] = g.ln(d)
    t["decoder.position_embeddings.weight"] = g.normal((cfg.dec_pos_rows, d), 0.02)
    dxa = cfg.dec_xa_heads * cfg.dec_xa_d_head
    kd = cfg.dec_kernel
    for l in range(cfg.dec_layers):
        p = f"decoder.layers.{l}."
        t[p + "norm_self.weight"] = g.ln(d)
        t[p + "self_attention.qkv_net.weight"] = g.uniform((3 * d, d), d)
        t[p + "self_attention.o_net.weight"] = g.uniform((d, d), d)
        t[p + "norm_xattn_query.weight"] = g.ln(d)
        t[p + "cross_attention.q_net.weight"] = g.uniform((dxa, d), d)
        t[p + "cross_attention.kv_net.weight"] = g.uniform((2 * dxa, d), d)
        t[p + "cross_attention.o_net.weight"] = g.uniform((d, dxa), dxa)
        t[p + "norm_xattn_memory.weight"] = g.ln(d)
        t[p + "norm_pos_ff.weight"] = g.ln(d)
        t[p + "pos_ff.proj.conv.weight"] = g.uniform((f, d, kd), d * kd)
        t[p + "pos_ff.o_net.conv.weight"] = g.uniform((d, f, kd), f * kd)
    t["decoder.norm_out.weight"] = g.ln(d)
    nv = cfg.num_codebooks * cfg.vocab_per_cb
    t["final_proj.weight"] = g.uniform((nv, d), d)
    t["final_proj.bias"] = g.uniform((nv,), d)
    L, lf = cfg.lt_dim, cfg.lt_ffn_dim
    t["local_transformer_in_projection.weight"] = g.uniform((L, d), d)
    t["local_transformer_in_projection.bias"] = g.uniform((L,), d)
    t["local_transformer.position_embeddings.weight"] = g.normal((cfg.lt_pos_rows, L), 0.02)
    p = "local_transformer.layers.0."
    t[p + "norm_self.weight"] = g.ln(L)
    t[p + "self_attention.qkv_net.weight"] = g.uniform((3 * L, L), L)
    t[p + "self_attention.o_net.weight"] = g.uniform((L, L), L)
    t[p + "norm_pos_ff.weight"] = g.ln(L)
    t[p + "pos_ff.proj.conv.weight"] = g.uniform((lf, L, 1), L)
    t[p + "pos_ff.o_net.conv.weight"] = g.uniform((L, lf, 1), lf)
    for cb in range(cfg.num_codebooks):
        t[f"local_transformer_out_projections.{cb}.weight"] = g.uniform((cfg.vocab_per_cb, L), L)
        t[f"local_transformer_out_projections.{cb}.bias"] = g.uniform((cfg.vocab_per_cb,), L)
    return dict(sorted(t.items()))


def write_magpie_gguf(path: str, cfg: MagpieConfig | None = None, seed: int = 1234,
                      quant: str = "f32", with_tokenizer: bool = True) -> MagpieConfig:
    """Write a random-init model GGUF. quant in {"f32","f16","q8_0"} (converter semantics)."""
    cfg = cfg or full_config()
    tensors = magpie_tensors(cfg, seed)
    w = GGUFWriter(path, "magpie")
    w.add_string("general.name", "magpie-tts-random-init")
    for k in _HP_KEYS:
        w.add_uint32("magpie." + k, int(getattr(cfg, k)))
    w.add_float32("magpie.eps", float(cfg.eps))
    if with_tokenizer:
        vocab, ids = synthetic_vocab()
        w.add_string("magpie.tokenizer.vocab", "\n".join(vocab))
        w.add_string("magpie.tokenizer.dict",
                     "".join(f"{k}\t{v}\n" for k, v in SYNTHETIC_DICT.items()))
        for k, v in ids.items():
            w.add_uint32("magpie.tokenizer." + k, v)
    quant = quant.lower()
    for name, a in tensors.items():
        qt = None
        if quant != "f32" and _should_quantize(name) and a.size >= 256 and a.ndim >= 2:
            if quant == "f16":
                qt = GGMLQuantizationType.F16
            elif quant in ("q8_0", "q8") and a.shape[-1] >= 32:
                qt = GGMLQuantizationType.Q8_0
        if qt is None:
            w.add_tensor(name, np.ascontiguousarray(a, dtype=np.float32))
        elif qt == GGMLQuantizationType.F16:
            w.add_tensor(name, a.astype(np.float16))
        else:
            q = gguf.quants.quantize(a, qt)
            w.add_tensor(name, q, raw_shape=q.shape, raw_dtype=qt)
    w.write_header_to_file()
    w.write_kv_data_to_file()
    w.write_tensors_to_file()
    w.close()
    return cfg


# ---------------------------------------------------------------------------------------------
# nano-codec
# ---------------------------------------------------------------------------------------------

@dataclass
class CodecConfig:
    sample_rate: int = 22050
    num_codebooks: int = 8
    codebook_size: int = 2016
    hop_length: int = 1024
    latent_dim: int = 32
    base_channels: int = 864
    up_rates: tuple = (8, 8, 4, 2, 2)
    up_channels: tuple = (432, 216, 108, 54, 27)
    res_kernels: tuple = (3, 7, 11)
    res_dilations: tuple = (1, 3, 5)
    pre_kernel: int = 7
    post_kernel: int = 3


def codec_tensors(cfg: CodecConfig | None = None, seed: int = 4321) -> "dict[str, np.ndarray]":
    """Codec tensors in PyTorch shapes; names as shortened by `convert_codec_to_gguf.py:110-132`."""
    cfg = cfg or CodecConfig()
    g = _Gen(seed)
    t: dict[str, np.ndarray] = {}
    t["dec.pre.weight"] = g.uniform((cfg.base_channels, cfg.latent_dim, cfg.pre_kernel),
                                    cfg.latent_dim * cfg.pre_kernel)
    t["dec.pre.bias"] = g.uniform((cfg.base_channels,), cfg.latent_dim * cfg.pre_kernel)
    cin = cfg.base_channels
    for i, (s, c) in enumerate(zip(cfg.up_rates, cfg.up_channels)):
        t[f"dec.act.{i}.activation.snake_act.alpha"] = \
            (0.5 + g.rng.random((1, cin // 2, 1), dtype=np.float32)).astype(np.float32)
        k = 2 * s
        # ConvTranspose1d(cin, c, k, stride=s, groups=c): weight (cin, 1, k); fan_in = 2*k/s taps
        t[f"dec.up.{i}.c.weight"] = g.uniform((cin, 1, k), 2 * k // s)
        t[f"dec.up.{i}.c.bias"] = g.uniform((c,), 2 * k // s)
        for j, kk in enumerate(cfg.res_kernels):
            for d_i in range(len(cfg.res_dilations)):
                p = f"dec.rl.{i}.rb.{j}.rb.{d_i}."
                t[p + "in_act.alpha"] = (0.5 + g.rng.random((1, c // 2, 1), dtype=np.float32))
                t[p + "in_conv.weight"] = g.uniform((c, c, kk), c * kk)
                t[p + "in_conv.bias"] = g.uniform((c,), c * kk)
                t[p + "sk_act.alpha"] = (0.5 + g.rng.random((1, c // 2, 1), dtype=np.float32))
                t[p + "sk_conv.weight"] = g.uniform((c, c, kk), c * kk)
                t[p + "sk_conv.bias"] = g.uniform((c,), c * kk)
        cin = c
    t["dec.post_act.alpha"] = (0.5 + g.rng.random((1, cin // 2, 1), dtype=np.float32))
    t["dec.post.weight"] = g.uniform((1, cin, cfg.post_kernel), cin * cfg.post_kernel)
    t["dec.post.bias"] = g.uniform((1,), cin * cfg.post_kernel)
    for i in range(cfg.num_codebooks):
        t[f"vq.fsqs.{i}.dim_base_index"] = np.array([1, 8, 56, 336], np.float32).reshape(1, 4, 1)
        t[f"vq.fsqs.{i}.num_levels"] = np.array([8, 7, 6, 6], np.float32).reshape(1, 4, 1)
    return dict(sorted(t.items()))


def write_codec_gguf(path: str, cfg: CodecConfig | None = None, seed: int = 4321,
                     f16: bool = False) -> CodecConfig:
    cfg = cfg or CodecConfig()
    tensors = codec_tensors(cfg, seed)
    w = GGUFWriter(path, "nano-codec")
    w.add_string("general.name", "nano-codec-random-init")
    for k in ("sample_rate", "num_codebooks", "codebook_size", "hop_length", "latent_dim"):
        w.add_uint32("codec." + k, int(getattr(cfg, k)))
    for name, a in tensors.items():
        a = np.ascontiguousarray(a, dtype=np.float32)
        if f16 and name.endswith(".weight"):
            w.add_tensor(name, a.astype(np.float16))
        else:
            w.add_tensor(name, a)
    w.write_header_to_file()
    w.write_kv_data_to_file()
    w.write_tensors_to_file()
    w.close()
    return cfg


def fixture_dir() -> str:
    d = os.environ.get("MAGPIE_FIXTURE_DIR", "/tmp/magpie_b200_fixtures")
    os.makedirs(d, exist_ok=True)
    return d


def ensure_fixture(kind: str) -> str:
    """Create (once per box) and return the path of a named fixture.

    kinds: model-f32, model-q8, model-f16, model-tiny, codec-f32, codec-f16, model-long-q8
    """
    path = os.path.join(fixture_dir(), kind + ".gguf")
    if os.path.exists(path) and os.path.getsize(path) > 0:
        return path
    tmp = path + f".tmp{os.getpid()}"
    if kind == "model-f32":
        write_magpie_gguf(tmp, full_config())
    elif kind == "model-f16":
        write_magpie_gguf(tmp, full_config(), quant="f16")
    elif kind == "model-q8":
        write_magpie_gguf(tmp, full_config(), quant="q8_0")
    elif kind == "model-long-q8":
        write_magpie_gguf(tmp, full_config(max_dec_steps=2600), quant="q8_0")
    elif kind == "model-tiny":
        write_magpie_gguf(tmp, tiny_config())
    elif kind == "codec-f32":
        write_codec_gguf(tmp)
    elif kind == "codec-f16":
        write_codec_gguf(tmp, f16=True)
    else:
        raise ValueError(kind)
    os.replace(tmp, path)
    return path


if __name__ == "__main__":
    import sys
    import time
    for k in sys.argv[1:] or ["model-tiny", "codec-f32", "model-f32"]:
        t0 = time.time()
        p = ensure_fixture(k)
        print(f"{k}: {p} {os.path.getsize(p) / 1e6:.1f} MB in {time.time() - t0:.1f}s")
