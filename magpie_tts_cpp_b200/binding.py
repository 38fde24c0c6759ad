"""ctypes binding of libmagpie_b200.so (the C-ABI in include/magpie_b200.h).

This is the host-side mirror used by tests/ and bench.py; the reference-shaped C++ API lives in
include/magpie.h.  There is NO fallback: if the CUDA library is missing or no device is visible,
every call raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libmagpie_b200.so")

PREC_F32, PREC_BF16 = 0, 1

HP_FIELDS = ["d_model", "d_ffn", "d_head", "enc_layers", "enc_heads", "enc_kernel", "dec_layers",
             "dec_sa_heads", "dec_xa_heads", "dec_xa_d_head", "dec_kernel", "lt_dim", "lt_ffn_dim",
             "lt_layers", "lt_heads", "text_vocab_size", "num_codebooks", "codebook_size",
             "vocab_per_cb", "num_speakers", "context_frames", "text_bos_id", "text_eos_id",
             "audio_bos_id", "audio_eos_id", "max_dec_steps", "sample_rate"]


class HParams(C.Structure):
    _fields_ = [(k, C.c_int32) for k in HP_FIELDS] + [("eps", C.c_float)]


STREAM_CB = C.CFUNCTYPE(C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int, C.c_void_p)


class CodecHParams(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("sample_rate", "num_codebooks", "codebook_size", "hop_length", "latent_dim")]


# every symbol include/magpie_b200.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "mgb_last_error", "mgb_device_count", "mgb_version", "mgb_shard_device",
    "mgb_model_load", "mgb_model_free", "mgb_model_get_hparams", "mgb_model_set_max_dec_steps",
    "mgb_model_set_gelu_f16", "mgb_model_precision", "mgb_model_device", "mgb_model_step_weight_bytes",
    "mgb_model_meta_str", "mgb_model_meta_u32",
    "mgb_session_new", "mgb_session_new_paged", "mgb_session_kv_pages", "mgb_generate_queue", "mgb_stream_generate", "mgb_session_free", "mgb_session_batch", "mgb_session_max_seq", "mgb_session_positions",
    "mgb_encode_text", "mgb_prefill", "mgb_decoder_step", "mgb_final_proj", "mgb_lt_sample",
    "mgb_generate", "mgb_teacher_forced", "mgb_session_last_loop_ms", "mgb_session_last_loop_launches",
    "mgb_session_debug_stamps",
    "mgb_pool_new", "mgb_pool_free", "mgb_pool_n_devices", "mgb_pool_model", "mgb_pool_generate", "mgb_pool_teacher_forced",
    "mgb_codec_load", "mgb_codec_free", "mgb_codec_get_hparams", "mgb_codec_decode",
    "mgb_codec_fsq_dequantize", "mgb_codec_last_ms", "mgb_codec_last_launches",
]


def build(force: bool = False) -> str:
    """Compile the CUDA extension in-tree (nvcc, sm_100a)."""
    src = os.path.join(_PKG, "csrc")
    args = ["make", "-C", src, "-s", "-j8"]
    if force:
        subprocess.check_call(["make", "-C", src, "-s", "clean"])
    subprocess.check_call(args)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, cp = C.c_void_p, C.c_char_p
    L.mgb_last_error.restype = cp
    L.mgb_version.restype = cp
    L.mgb_device_count.restype = C.c_int
    L.mgb_shard_device.argtypes = [C.c_int64, C.c_int]
    L.mgb_model_load.restype = vp
    L.mgb_model_load.argtypes = [cp, C.c_int, C.c_int]
    L.mgb_model_free.argtypes = [vp]
    L.mgb_model_get_hparams.argtypes = [vp, C.POINTER(HParams)]
    L.mgb_model_set_max_dec_steps.argtypes = [vp, C.c_int32]
    L.mgb_model_set_gelu_f16.argtypes = [vp, C.c_int]
    L.mgb_model_precision.argtypes = [vp]
    L.mgb_model_device.argtypes = [vp]
    L.mgb_model_step_weight_bytes.restype = C.c_int64
    L.mgb_model_step_weight_bytes.argtypes = [vp]
    L.mgb_model_meta_str.restype = cp
    L.mgb_model_meta_str.argtypes = [vp, cp]
    L.mgb_model_meta_u32.restype = C.c_int32
    L.mgb_model_meta_u32.argtypes = [vp, cp, C.c_int32]
    L.mgb_session_new.restype = vp
    L.mgb_session_new.argtypes = [vp, C.c_int, C.c_int, C.c_int]
    L.mgb_session_free.argtypes = [vp]
    L.mgb_session_new_paged.restype = vp
    L.mgb_session_new_paged.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int]
    L.mgb_session_kv_pages.argtypes = [vp, vp, vp]
    L.mgb_stream_generate.argtypes = [vp, vp, C.c_int, C.c_float, C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_int, STREAM_CB, vp, vp]
    L.mgb_generate_queue.argtypes = [vp, C.c_int, vp, vp, C.c_int, vp, vp, C.c_int, C.c_float, C.c_int, C.c_uint64, vp, vp, vp]
    L.mgb_session_batch.argtypes = [vp]
    L.mgb_session_max_seq.argtypes = [vp]
    L.mgb_session_positions.argtypes = [vp, vp]
    L.mgb_encode_text.argtypes = [vp, vp, vp, vp]
    L.mgb_prefill.argtypes = [vp, vp]
    L.mgb_decoder_step.argtypes = [vp, vp, vp]
    L.mgb_final_proj.argtypes = [vp, vp, vp]
    L.mgb_lt_sample.argtypes = [vp, vp, C.c_float, C.c_int, vp, vp, vp, C.c_uint64, vp, vp, vp]
    L.mgb_generate.argtypes = [vp, C.c_int, C.c_float, C.c_int, vp, C.c_uint64, C.c_int, vp, vp, vp]
    L.mgb_teacher_forced.argtypes = [vp, vp, C.c_int, vp, vp, vp]
    L.mgb_session_last_loop_ms.restype = C.c_float
    L.mgb_session_last_loop_ms.argtypes = [vp]
    L.mgb_session_last_loop_launches.restype = C.c_int64
    L.mgb_session_last_loop_launches.argtypes = [vp]
    L.mgb_session_debug_stamps.argtypes = [vp, vp, C.c_int]
    L.mgb_pool_new.restype = vp
    L.mgb_pool_new.argtypes = [cp, vp, C.c_int, C.c_int]
    L.mgb_pool_free.argtypes = [vp]
    L.mgb_pool_n_devices.argtypes = [vp]
    L.mgb_pool_model.restype = vp
    L.mgb_pool_model.argtypes = [vp, C.c_int]
    L.mgb_pool_generate.argtypes = [vp, C.c_int, vp, vp, C.c_int, vp, C.c_int, C.c_float, C.c_int, C.c_uint64, C.c_int, vp, vp, vp]
    L.mgb_pool_teacher_forced.argtypes = [vp, C.c_int, vp, vp, C.c_int, vp, vp, C.c_int, vp, vp]
    L.mgb_codec_load.restype = vp
    L.mgb_codec_load.argtypes = [cp, C.c_int]
    L.mgb_codec_free.argtypes = [vp]
    L.mgb_codec_get_hparams.argtypes = [vp, C.POINTER(CodecHParams)]
    L.mgb_codec_decode.argtypes = [vp, vp, C.c_int, C.c_int, vp]
    L.mgb_codec_fsq_dequantize.argtypes = [vp, vp, C.c_int, C.c_int, vp]
    L.mgb_codec_last_ms.restype = C.c_float
    L.mgb_codec_last_ms.argtypes = [vp]
    L.mgb_codec_last_launches.restype = C.c_int64
    L.mgb_codec_last_launches.argtypes = [vp]
    _lib = L
    return L


class MagpieError(RuntimeError):
    pass


def _err(what):
    return MagpieError(f"{what}: {lib().mgb_last_error().decode('utf-8', 'replace')}")


def _chk(rc, what):
    if rc != 0:
        raise _err(what)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def device_count() -> int:
    return int(lib().mgb_device_count())


class Model:
    """magpie_init / magpie_free (reference src/magpie.cpp:777-915) on device `device`."""

    def __init__(self, gguf_path: str, device: int = 0, precision: int = PREC_F32):
        self._h = lib().mgb_model_load(os.fsencode(gguf_path), int(device), int(precision))
        if not self._h:
            raise _err("magpie_init")
        hp = HParams()
        _chk(lib().mgb_model_get_hparams(self._h, C.byref(hp)), "hparams")
        self.hp = {k: getattr(hp, k) for k in HP_FIELDS}
        self.hp["eps"] = hp.eps
        self.precision = precision
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            lib().mgb_model_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_gelu_f16(self, on: bool):
        _chk(lib().mgb_model_set_gelu_f16(self._h, int(on)), "set_gelu_f16")

    def set_max_dec_steps(self, n: int):
        _chk(lib().mgb_model_set_max_dec_steps(self._h, int(n)), "set_max_dec_steps")
        self.hp["max_dec_steps"] = int(n)

    @property
    def step_weight_bytes(self) -> int:
        return int(lib().mgb_model_step_weight_bytes(self._h))

    def meta_str(self, key: str):
        r = lib().mgb_model_meta_str(self._h, key.encode())
        return r.decode("utf-8") if r is not None else None

    def meta_u32(self, key: str, default: int = -1) -> int:
        return int(lib().mgb_model_meta_u32(self._h, key.encode(), default))

    def session(self, batch: int = 1, max_text: int = 128, max_seq: int = 0, kv_pages: int = 0) -> "Session":
        return Session(self, batch, max_text, max_seq, kv_pages)


class Session:
    """Device state of `batch` independent utterances (KV caches, positions)."""

    def __init__(self, model: Model, batch: int, max_text: int, max_seq: int = 0, kv_pages: int = 0):
        self.model = model
        self._h = lib().mgb_session_new_paged(model._h, int(batch), int(max_text), int(max_seq), int(kv_pages))
        if not self._h:
            raise _err("mgb_session_new")
        self.B = batch
        self.max_text = max_text
        self.max_seq = int(lib().mgb_session_max_seq(self._h))

    def close(self):
        if getattr(self, "_h", None):
            lib().mgb_session_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def pos(self) -> int:
        p = np.zeros(self.B, np.int32)
        _chk(lib().mgb_session_positions(self._h, _p(p)), "positions")
        return int(p[0])

    def encode_text(self, token_lists, want_output: bool = True):
        """token_lists: list of B token-id lists. Returns list of [E_b][d] arrays (or None)."""
        assert len(token_lists) == self.B
        tok = np.zeros((self.B, self.max_text), np.int32)
        n = np.zeros(self.B, np.int32)
        for b, t in enumerate(token_lists):
            t = list(t)
            if len(t) > self.max_text:
                raise MagpieError("encode_text: too many tokens for this session")
            tok[b, :len(t)] = t
            n[b] = len(t)
        d = self.model.hp["d_model"]
        out = np.empty((self.B, self.max_text, d), np.float32) if want_output else None
        _chk(lib().mgb_encode_text(self._h, _p(tok), _p(n), _p(out)), "magpie_encode_text")
        if out is None:
            return None
        return [out[b, :n[b]].copy() for b in range(self.B)]

    def prefill(self, speakers=None):
        spk = _i32(speakers if speakers is not None else [0] * self.B)
        assert spk.shape == (self.B,)
        _chk(lib().mgb_prefill(self._h, _p(spk)), "mgb_prefill")

    def decoder_step(self, codes=None, want_hidden: bool = True):
        c = _i32(codes).reshape(self.B, 8) if codes is not None else None
        out = np.empty((self.B, self.model.hp["d_model"]), np.float32) if want_hidden else None
        _chk(lib().mgb_decoder_step(self._h, _p(c), _p(out)), "mgb_decoder_step")
        return out

    def final_proj(self, hidden=None):
        hp = self.model.hp
        h = _f32(hidden).reshape(self.B, hp["d_model"]) if hidden is not None else None
        out = np.empty((self.B, hp["num_codebooks"] * hp["vocab_per_cb"]), np.float32)
        _chk(lib().mgb_final_proj(self._h, _p(h), _p(out)), "mgb_final_proj")
        return out

    def lt_sample(self, hidden=None, temperature=0.0, top_k=80, forbid_eos=None, forced_codes=None,
                  uniforms=None, seed=0, want_logits=True):
        hp = self.model.hp
        h = _f32(hidden).reshape(self.B, hp["d_model"]) if hidden is not None else None
        fe = np.ascontiguousarray(np.broadcast_to(np.asarray(forbid_eos, np.uint8), (self.B,))) if forbid_eos is not None else None
        fc = _i32(forced_codes).reshape(self.B, 8) if forced_codes is not None else None
        u = _f32(uniforms).reshape(self.B, 8) if uniforms is not None else None
        sampled = np.zeros((self.B, 8), np.int32)
        argmax = np.zeros((self.B, 8), np.int32)
        logits = np.empty((self.B, 8, hp["vocab_per_cb"]), np.float32) if want_logits else None
        _chk(lib().mgb_lt_sample(self._h, _p(h), float(temperature), int(top_k), _p(fe), _p(fc), _p(u),
                                 C.c_uint64(seed), _p(sampled), _p(argmax), _p(logits)),
             "magpie_local_transformer_sample_all")
        return sampled, argmax, logits

    def generate(self, max_steps=0, temperature=0.0, top_k=80, uniforms=None, seed=0, ignore_eos=False,
                 want_hidden=False):
        hp = self.model.hp
        T = max_steps if max_steps > 0 else hp["max_dec_steps"]
        u = _f32(uniforms).reshape(self.B, T, 8) if uniforms is not None else None
        codes = np.zeros((self.B, T, 8), np.int32)
        n = np.zeros(self.B, np.int32)
        hid = np.empty((self.B, T, hp["d_model"]), np.float32) if want_hidden else None
        _chk(lib().mgb_generate(self._h, int(T), float(temperature), int(top_k), _p(u), C.c_uint64(seed),
                                int(ignore_eos), _p(codes), _p(n), _p(hid)), "magpie_synthesize_codes")
        out = [codes[b, :n[b]].copy() for b in range(self.B)]
        return (out, hid) if want_hidden else out

    @property
    def kv_pages(self):
        """(pages in the pool, pages currently assigned to utterances) of the paged self-attention cache."""
        t, u = C.c_int32(0), C.c_int32(0)
        _chk(lib().mgb_session_kv_pages(self._h, C.byref(t), C.byref(u)), "kv_pages")
        return int(t.value), int(u.value)

    def generate_queue(self, token_lists, speakers=None, max_steps=0, max_steps_per_utt=None, temperature=0.0, top_k=80, seed=0):
        """Continuous batching: len(token_lists) >= B utterances streamed through the B slots.  Returns (codes list, steps run)."""
        hp = self.model.hp
        n = len(token_lists)
        T = max_steps if max_steps > 0 else hp["max_dec_steps"]
        mt = max(len(t) for t in token_lists)
        tok = np.zeros((n, mt), np.int32)
        nt = np.zeros(n, np.int32)
        for i, t in enumerate(token_lists):
            tok[i, :len(t)] = t
            nt[i] = len(t)
        spk = _i32(speakers) if speakers is not None else None
        lim = _i32(max_steps_per_utt) if max_steps_per_utt is not None else None
        codes = np.zeros((n, T, 8), np.int32)
        nf = np.zeros(n, np.int32)
        steps = C.c_int64(0)
        _chk(lib().mgb_generate_queue(self._h, n, _p(tok), _p(nt), mt, _p(spk), _p(lim), int(T), float(temperature), int(top_k),
                                      C.c_uint64(seed), _p(codes), _p(nf), C.byref(steps)), "mgb_generate_queue")
        return [codes[i, :nf[i]].copy() for i in range(n)], int(steps.value)

    def stream_generate(self, codec: "Codec", on_audio, max_steps=0, temperature=0.0, top_k=80, seed=0, ignore_eos=False,
                        frames_per_chunk=4, codec_context_frames=0):
        """Batched streaming synthesis: on_audio(utterance, pcm ndarray, frames_done, is_last) per utterance and chunk
        (return True to stop).  Returns frames per utterance."""
        def tramp(u, ptr, n, fd, last, _user):
            try:
                return 1 if on_audio(int(u), np.ctypeslib.as_array(ptr, shape=(n,)).copy(), int(fd), bool(last)) else 0
            except Exception:      # noqa: BLE001 -- an exception must not unwind through the C frames
                import traceback
                traceback.print_exc()
                return 1
        cbf = STREAM_CB(tramp)
        nf = np.zeros(self.B, np.int32)
        _chk(lib().mgb_stream_generate(self._h, codec._h, int(max_steps), float(temperature), int(top_k), C.c_uint64(seed), int(ignore_eos),
                                       int(frames_per_chunk), int(codec_context_frames), cbf, None, _p(nf)), "mgb_stream_generate")
        return nf

    def teacher_forced(self, codes_in, want_hidden=True, want_logits=True, want_greedy=True):
        hp = self.model.hp
        c = _i32(codes_in)
        assert c.ndim == 3 and c.shape[0] == self.B and c.shape[2] == 8
        T = c.shape[1]
        hid = np.empty((self.B, T, hp["d_model"]), np.float32) if want_hidden else None
        lg = np.empty((self.B, T, 8, hp["vocab_per_cb"]), np.float32) if want_logits else None
        gr = np.zeros((self.B, T, 8), np.int32) if want_greedy else None
        _chk(lib().mgb_teacher_forced(self._h, _p(c), int(T), _p(hid), _p(lg), _p(gr)), "mgb_teacher_forced")
        return hid, lg, gr

    def debug_stamps(self, n: int) -> np.ndarray:
        out = np.zeros(n, np.uint64)
        _chk(lib().mgb_session_debug_stamps(self._h, _p(out), n), "debug_stamps")
        return out

    @property
    def last_loop_ms(self) -> float:
        return float(lib().mgb_session_last_loop_ms(self._h))

    @property
    def last_loop_launches(self) -> int:
        return int(lib().mgb_session_last_loop_launches(self._h))


class Pool:
    """In-process multi-GPU synthesis (mgb_pool_*): one replica + one submission thread per device, utterance i on device
    i mod G, no collective.  The C++ side owns the threads; this is only the ctypes mirror."""

    def __init__(self, gguf_path: str, devices=None, precision: int = PREC_BF16):
        dv = _i32(devices) if devices is not None else None
        self._h = lib().mgb_pool_new(os.fsencode(gguf_path), _p(dv), 0 if dv is None else len(dv), int(precision))
        if not self._h:
            raise _err("mgb_pool_new")
        self.n_devices = int(lib().mgb_pool_n_devices(self._h))
        hp = HParams()
        _chk(lib().mgb_model_get_hparams(lib().mgb_pool_model(self._h, 0), C.byref(hp)), "hparams")
        self.hp = {k: getattr(hp, k) for k in HP_FIELDS}
        self.last_device_ms = np.zeros(self.n_devices, np.float32)

    def close(self):
        if getattr(self, "_h", None):
            lib().mgb_pool_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _pad(token_lists):
        mt = max(len(t) for t in token_lists)
        tok = np.zeros((len(token_lists), mt), np.int32)
        n = np.zeros(len(token_lists), np.int32)
        for i, t in enumerate(token_lists):
            tok[i, :len(t)] = t
            n[i] = len(t)
        return tok, n, mt

    def generate(self, token_lists, speakers=None, max_steps=0, temperature=0.0, top_k=80, seed=0, ignore_eos=False):
        tok, n, mt = self._pad(token_lists)
        T = max_steps if max_steps > 0 else self.hp["max_dec_steps"]
        spk = _i32(speakers) if speakers is not None else None
        codes = np.zeros((len(token_lists), T, 8), np.int32)
        nf = np.zeros(len(token_lists), np.int32)
        _chk(lib().mgb_pool_generate(self._h, len(token_lists), _p(tok), _p(n), mt, _p(spk), int(T), float(temperature), int(top_k),
                                     C.c_uint64(seed), int(ignore_eos), _p(codes), _p(nf), _p(self.last_device_ms)), "mgb_pool_generate")
        return [codes[i, :nf[i]].copy() for i in range(len(token_lists))]

    def teacher_forced(self, token_lists, codes_in, speakers=None, want_greedy=True):
        tok, n, mt = self._pad(token_lists)
        c = _i32(codes_in)
        assert c.ndim == 3 and c.shape[0] == len(token_lists) and c.shape[2] == 8
        spk = _i32(speakers) if speakers is not None else None
        gr = np.zeros_like(c) if want_greedy else None
        _chk(lib().mgb_pool_teacher_forced(self._h, len(token_lists), _p(tok), _p(n), mt, _p(spk), _p(c), int(c.shape[1]), _p(gr),
                                           _p(self.last_device_ms)), "mgb_pool_teacher_forced")
        return gr


class Codec:
    """magpie_codec_init / magpie_codec_decode (reference src/nano-codec.cpp:339-374, 758-845)."""

    def __init__(self, gguf_path: str, device: int = 0):
        self._h = lib().mgb_codec_load(os.fsencode(gguf_path), int(device))
        if not self._h:
            raise _err("magpie_codec_init")
        hp = CodecHParams()
        _chk(lib().mgb_codec_get_hparams(self._h, C.byref(hp)), "codec hparams")
        self.hp = {k: getattr(hp, k) for k, _ in CodecHParams._fields_}
        self.hop = self.hp["hop_length"]

    def close(self):
        if getattr(self, "_h", None):
            lib().mgb_codec_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def decode(self, codes) -> np.ndarray:
        """codes [8][T] or [B][8][T] codebook-major -> pcm [T*hop] or [B][T*hop]."""
        c = _i32(codes)
        single = c.ndim == 2
        if single:
            c = c[None]
        assert c.ndim == 3 and c.shape[1] == 8
        B, _, T = c.shape
        out = np.empty((B, T * self.hop), np.float32)
        _chk(lib().mgb_codec_decode(self._h, _p(c), B, T, _p(out)), "magpie_codec_decode")
        return out[0] if single else out

    def fsq_dequantize(self, codes) -> np.ndarray:
        c = _i32(codes)
        single = c.ndim == 2
        if single:
            c = c[None]
        B, _, T = c.shape
        out = np.empty((B, 32, T), np.float32)
        _chk(lib().mgb_codec_fsq_dequantize(self._h, _p(c), B, T, _p(out)), "fsq_dequantize")
        return out[0] if single else out

    @property
    def last_ms(self) -> float:
        return float(lib().mgb_codec_last_ms(self._h))

    @property
    def last_launches(self) -> int:
        return int(lib().mgb_codec_last_launches(self._h))
