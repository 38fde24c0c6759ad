"""ctypes front-end for oracle/libmagpie_oracle.so (CPU restatement of the reference hot path).

TEST INFRASTRUCTURE ONLY -- see the header of magpie_oracle.c.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module.

GGUF files are read with the independent `gguf` python package (not with the product's C++
reader), so this path also cross-validates the product's GGUF parser.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmagpie_oracle.so")

HP_FIELDS = ["d_model", "d_ffn", "d_head", "enc_layers", "enc_heads", "enc_kernel", "dec_layers",
             "dec_sa_heads", "dec_xa_heads", "dec_xa_d_head", "dec_kernel", "lt_dim", "lt_ffn_dim",
             "lt_layers", "lt_heads", "text_vocab_size", "num_codebooks", "codebook_size",
             "vocab_per_cb", "num_speakers", "context_frames", "text_bos_id", "text_eos_id",
             "audio_bos_id", "audio_eos_id", "max_dec_steps", "sample_rate"]
HP_DEFAULTS = dict(d_model=768, d_ffn=3072, d_head=64, enc_layers=6, enc_heads=12, enc_kernel=3,
                   dec_layers=12, dec_sa_heads=12, dec_xa_heads=1, dec_xa_d_head=128, dec_kernel=1,
                   lt_dim=256, lt_ffn_dim=1024, lt_layers=1, lt_heads=1, text_vocab_size=2380,
                   num_codebooks=8, codebook_size=2016, vocab_per_cb=2024, num_speakers=5,
                   context_frames=110, text_bos_id=2378, text_eos_id=2379, audio_bos_id=2016,
                   audio_eos_id=2017, max_dec_steps=500, sample_rate=22050, eps=1e-5)


class HParams(C.Structure):
    _fields_ = [(k, C.c_int32) for k in HP_FIELDS] + [("eps", C.c_float)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "magpie_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or \
            (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_LIB_PATH)):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        build()
    L = C.CDLL(_LIB_PATH)
    vp, i32p, f32p, i64p = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_float), C.POINTER(C.c_int64)
    L.orc_model_new.restype = vp
    L.orc_model_new.argtypes = [C.POINTER(HParams)]
    L.orc_model_free.argtypes = [vp]
    L.orc_model_set_gelu_table.argtypes = [vp, C.c_int]
    L.orc_model_set_tensor.restype = C.c_int
    L.orc_model_set_tensor.argtypes = [vp, C.c_char_p, vp, C.c_int, C.c_int, i64p]
    L.orc_encode_text.restype = C.c_int
    L.orc_encode_text.argtypes = [vp, vp, C.c_int, vp]
    L.orc_state_new.restype = vp
    L.orc_state_new.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int]
    L.orc_state_free.argtypes = [vp]
    L.orc_state_pos.restype = C.c_int
    L.orc_state_pos.argtypes = [vp]
    L.orc_state_max_seq.restype = C.c_int
    L.orc_state_max_seq.argtypes = [vp]
    for n in ("orc_state_k", "orc_state_v", "orc_state_xk", "orc_state_xv"):
        getattr(L, n).restype = vp
        getattr(L, n).argtypes = [vp]
    L.orc_audio_embedding.argtypes = [vp, vp, vp]
    L.orc_decoder_step.argtypes = [vp, vp, vp]
    L.orc_final_proj.argtypes = [vp, vp, vp]
    L.orc_sample_top_k.restype = C.c_int32
    L.orc_sample_top_k.argtypes = [vp, C.c_int, C.c_float, C.c_int, C.c_float]
    L.orc_lt_sample.argtypes = [vp, vp, C.c_float, C.c_int, C.c_int, vp, vp, vp, vp, vp]
    L.orc_synthesize.restype = C.c_int
    L.orc_synthesize.argtypes = [vp, vp, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, vp, vp, vp]
    L.orc_codec_new.restype = vp
    L.orc_codec_free.argtypes = [vp]
    L.orc_codec_set_conv_f16.argtypes = [vp, C.c_int]
    L.orc_codec_set_tensor.restype = C.c_int
    L.orc_codec_set_tensor.argtypes = [vp, C.c_char_p, vp, C.c_int, C.c_int, i64p]
    L.orc_fsq_dequantize.argtypes = [vp, C.c_int, C.c_int, vp]
    L.orc_codec_decode.restype = C.c_int
    L.orc_codec_decode.argtypes = [vp, vp, C.c_int, vp]
    L.orc_codec_half_snake.argtypes = [vp, vp, C.c_int, C.c_int, vp, C.c_int]
    L.orc_codec_causal_conv1d.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, C.c_int]
    L.orc_codec_conv_transpose1d.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int]
    L.orc_layer_norm.argtypes = [vp, vp, vp, C.c_int, C.c_float]
    L.orc_gelu.restype = C.c_float
    L.orc_gelu.argtypes = [C.c_float, C.c_int]
    L.orc_num_threads.restype = C.c_int
    L.orc_set_activation_rounding.argtypes = [C.c_int]
    L.orc_set_num_threads.argtypes = [C.c_int]
    _lib = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def read_gguf(path):
    """-> (kv dict of python scalars/strings, list of (name, ggml_type_int, ne list, raw ndarray))."""
    from gguf import GGUFReader, GGUFValueType
    r = GGUFReader(path)
    kv = {}
    for k, f in r.fields.items():
        if not f.types:
            continue
        t = f.types[0]
        if t == GGUFValueType.STRING:
            kv[k] = bytes(f.parts[-1]).decode("utf-8")
        elif t != GGUFValueType.ARRAY:
            kv[k] = f.parts[-1][0].item()
    tensors = []
    for t in r.tensors:
        tensors.append((t.name, int(t.tensor_type), [int(x) for x in t.shape], t.data))
    return kv, tensors, r


class OracleModel:
    """The reference's model object restated: owns dequantised f32 weights on the host."""

    def __init__(self, gguf_path: str, overrides: dict | None = None):
        L = lib()
        kv, tensors, self._reader = read_gguf(gguf_path)
        self.kv = kv
        hp = HParams()
        self.hp = {}
        for k in HP_FIELDS:
            v = int(kv.get("magpie." + k, HP_DEFAULTS[k]))
            if overrides and k in overrides:
                v = int(overrides[k])
            setattr(hp, k, v)
            self.hp[k] = v
        hp.eps = float(kv.get("magpie.eps", HP_DEFAULTS["eps"]))
        self.hp["eps"] = hp.eps
        self._h = L.orc_model_new(C.byref(hp))
        self.n_mapped = 0
        for name, ty, ne, data in tensors:
            arr = np.ascontiguousarray(data)
            nea = (C.c_int64 * 4)(*(ne + [1] * (4 - len(ne))))
            self.n_mapped += L.orc_model_set_tensor(self._h, name.encode(), _p(arr), ty, len(ne), nea)

    def __del__(self):
        try:
            lib().orc_model_free(self._h)
        except Exception:
            pass

    def set_gelu_table(self, on: bool):
        lib().orc_model_set_gelu_table(self._h, int(on))

    def encode_text(self, tokens) -> np.ndarray:
        tok = _i32(tokens)
        out = np.empty((len(tok), self.hp["d_model"]), np.float32)
        ok = lib().orc_encode_text(self._h, _p(tok), len(tok), _p(out))
        if not ok:
            raise ValueError("encode_text: invalid args")
        return out

    def new_state(self, enc_out, speaker=0, max_seq=0) -> "OracleState":
        return OracleState(self, enc_out, speaker, max_seq)

    def audio_embedding(self, codes) -> np.ndarray:
        c = _i32(codes)
        out = np.empty(self.hp["d_model"], np.float32)
        lib().orc_audio_embedding(self._h, _p(c), _p(out))
        return out

    def final_proj(self, hidden) -> np.ndarray:
        h = _f32(hidden)
        out = np.empty(self.hp["num_codebooks"] * self.hp["vocab_per_cb"], np.float32)
        lib().orc_final_proj(self._h, _p(h), _p(out))
        return out

    def lt_sample(self, hidden, temperature=0.0, top_k=80, forbid_eos=False, forced_codes=None,
                  uniforms=None, want_logits=True):
        h = _f32(hidden)
        V = self.hp["vocab_per_cb"]
        sampled = np.zeros(8, np.int32)
        argmax = np.zeros(8, np.int32)
        logits = np.empty((8, V), np.float32) if want_logits else None
        fc = _i32(forced_codes) if forced_codes is not None else None
        u = _f32(uniforms) if uniforms is not None else None
        lib().orc_lt_sample(self._h, _p(h), float(temperature), int(top_k), int(forbid_eos), _p(fc), _p(u),
                            _p(sampled), _p(argmax), _p(logits))
        return sampled, argmax, logits

    def synthesize(self, tokens, speaker=0, temperature=0.0, top_k=80, max_steps=0, uniforms=None,
                   want_hidden=False):
        tok = _i32(tokens)
        ms = max_steps if max_steps > 0 else self.hp["max_dec_steps"]
        codes = np.zeros((ms, 8), np.int32)
        hid = np.zeros((ms, self.hp["d_model"]), np.float32) if want_hidden else None
        u = _f32(uniforms) if uniforms is not None else None
        n = lib().orc_synthesize(self._h, _p(tok), len(tok), int(speaker), float(temperature), int(top_k),
                                 int(ms), _p(u), _p(codes), _p(hid))
        if n < 0:
            raise RuntimeError("synthesize failed")
        return (codes[:n], hid) if want_hidden else codes[:n]


class OracleState:
    def __init__(self, model: OracleModel, enc_out, speaker=0, max_seq=0):
        self.model = model
        e = _f32(enc_out)
        self.E = e.shape[0]
        self._h = lib().orc_state_new(model._h, _p(e), self.E, int(speaker), int(max_seq))

    def __del__(self):
        try:
            lib().orc_state_free(self._h)
        except Exception:
            pass

    @property
    def pos(self):
        return lib().orc_state_pos(self._h)

    def step(self, codes) -> np.ndarray:
        c = _i32(codes)
        out = np.empty(self.model.hp["d_model"], np.float32)
        lib().orc_decoder_step(self._h, _p(c), _p(out))
        return out

    def _view(self, fn, shape):
        ptr = getattr(lib(), fn)(self._h)
        n = int(np.prod(shape))
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float)), shape=(n,)).reshape(shape).copy()

    def kv(self):
        hp = self.model.hp
        ms = lib().orc_state_max_seq(self._h)
        shp = (hp["dec_layers"], ms, hp["d_model"])
        return self._view("orc_state_k", shp), self._view("orc_state_v", shp)

    def xkv(self):
        hp = self.model.hp
        shp = (hp["dec_layers"], self.E, hp["dec_xa_heads"] * hp["dec_xa_d_head"])
        return self._view("orc_state_xk", shp), self._view("orc_state_xv", shp)


class OracleCodec:
    def __init__(self, gguf_path: str, conv_f16: bool = True):
        L = lib()
        kv, tensors, self._reader = read_gguf(gguf_path)
        self.kv = kv
        self.hop = int(kv.get("codec.hop_length", 1024))
        self.sample_rate = int(kv.get("codec.sample_rate", 22050))
        self._h = L.orc_codec_new()
        L.orc_codec_set_conv_f16(self._h, int(conv_f16))
        self.n_mapped = 0
        for name, ty, ne, data in tensors:
            arr = np.ascontiguousarray(data)
            nea = (C.c_int64 * 4)(*(ne + [1] * (4 - len(ne))))
            self.n_mapped += L.orc_codec_set_tensor(self._h, name.encode(), _p(arr), ty, len(ne), nea)

    def __del__(self):
        try:
            lib().orc_codec_free(self._h)
        except Exception:
            pass

    def decode(self, codes) -> np.ndarray:
        """codes: [8][T] codebook-major int32 -> pcm float32 [T*hop]."""
        c = _i32(codes)
        assert c.ndim == 2 and c.shape[0] == 8
        T = c.shape[1]
        out = np.empty(T * self.hop, np.float32)
        lib().orc_codec_decode(self._h, _p(c), T, _p(out))
        return out


def fsq_dequantize(codes) -> np.ndarray:
    c = _i32(codes)
    ncb, T = c.shape
    out = np.empty((ncb * 4, T), np.float32)
    lib().orc_fsq_dequantize(_p(c), ncb, T, _p(out))
    return out


def sample_top_k(logits, temperature, top_k, u) -> int:
    l = _f32(logits)
    return int(lib().orc_sample_top_k(_p(l), len(l), float(temperature), int(top_k), float(u)))


def half_snake(x, alpha):
    x = _f32(x); a = _f32(alpha).reshape(-1)
    y = np.empty_like(x)
    lib().orc_codec_half_snake(_p(x), _p(y), x.shape[0], x.shape[1], _p(a), a.size)
    return y


def causal_conv1d(x, w, b, dil=1, f16=True):
    x = _f32(x); w = _f32(w); bb = _f32(b) if b is not None else None
    co, ci, k = w.shape
    y = np.empty((co, x.shape[1]), np.float32)
    lib().orc_codec_causal_conv1d(_p(x), _p(y), ci, co, x.shape[1], k, _p(w), _p(bb), int(dil), int(f16))
    return y


def conv_transpose1d(x, w, b, stride):
    x = _f32(x); w = _f32(w); bb = _f32(b) if b is not None else None
    ci, _, k = w.shape
    y = np.empty((ci // 2, x.shape[1] * stride), np.float32)
    lib().orc_codec_conv_transpose1d(_p(x), _p(y), ci, x.shape[1], k, _p(w), _p(bb), int(stride))
    return y


def layer_norm(x, w, eps=1e-5):
    x = _f32(x); w = _f32(w)
    y = np.empty_like(x)
    lib().orc_layer_norm(_p(x), _p(w), _p(y), x.size, float(eps))
    return y


def gelu(x: float, table=True) -> float:
    return float(lib().orc_gelu(float(x), int(table)))


def set_activation_rounding(on: bool):
    """ggml-CPU mul_mat rounds the activation row to the weight's storage type (f16 / Q8_0 blocks); off = weights only."""
    lib().orc_set_activation_rounding(int(on))


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int):
    lib().orc_set_num_threads(int(n))
