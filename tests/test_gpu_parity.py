"""GPU parity tests: the sm_100a path (through the C-ABI) against the CPU oracle on identical
random-init GGUF weights and inputs.

Tolerances (BASELINE.json north_star): f32 logits/hidden within 1e-4, bf16 within 2e-2, measured as
|a-b| <= tol*|b| + tol*rms(b) (pure rtol is undefined at zero crossings); greedy codes
identical for >= 99 % of frames; FSQ bit-exact; codec waveform >= 40 dB SNR.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HELLO = [2378, 7, 4, 11, 11, 14, 32, 26, 22, 14, 17, 11, 3, 32, 28, 2379]  # "Hello, world!" synthetic vocab


def close(a, b, tol):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    fin = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), fin)
    assert np.array_equal(a[~fin], b[~fin])
    rms = np.sqrt(np.mean(b[fin] ** 2))
    err = np.abs(a[fin] - b[fin]) - tol * np.abs(b[fin]) - tol * rms
    worst = float(err.max())
    assert worst <= 0, f"tolerance {tol} exceeded by {worst:.3e} (rms {rms:.3e}, max abs diff {np.abs(a[fin]-b[fin]).max():.3e})"


def snr_db(x, ref):
    x = np.asarray(x, np.float64); ref = np.asarray(ref, np.float64)
    return 10 * np.log10(np.sum(ref ** 2) / max(np.sum((x - ref) ** 2), 1e-300))


@pytest.fixture(scope="module")
def B():
    from magpie_tts_cpp_b200 import binding
    return binding


# The reference's CPU path evaluates GELU through ggml's f16 lookup table: a
# discontinuous function.  A 1e-6 relative perturbation of its input flips table entries and moves the
# 12-layer decoder output by ~5e-4 (tests/test_oracle_cpu.py::test_gelu_f16_table_amplifies_rounding_noise),
# so the f32 bar of 1e-4 is checked with the table OFF on both sides (pure f32 arithmetic, where the oracle
# agrees with float64 to ~1e-6), and the table-ON mode (bit-for-bit the reference's semantics, the product
# default) is checked at TABLE_TOL, the reference path's own noise floor.
TABLE_TOL = 3e-3


def set_table(m, o, on):
    m.set_gelu_f16(on)
    o.set_gelu_table(on)


@pytest.fixture(scope="module")
def tiny_pair(B, oracle_mod, tiny_model_path):
    m, o = B.Model(tiny_model_path, 0, B.PREC_F32), oracle_mod.OracleModel(tiny_model_path)
    set_table(m, o, False)
    return m, o


# ---- tiny architecture, f32: every stage against the oracle ------------------------------------------

def test_encoder_f32(tiny_pair):
    m, o = tiny_pair
    s = m.session(batch=2, max_text=32)
    toks = [HELLO, HELLO[:9] + [2379]]
    enc = s.encode_text(toks)
    for b in range(2):
        close(enc[b], o.encode_text(toks[b]), 1e-4)


def test_prefill_and_decoder_steps_f32(tiny_pair):
    m, o = tiny_pair
    hp = o.hp
    s = m.session(batch=2, max_text=32)
    toks = [HELLO, HELLO[:12] + [2379]]
    enc = s.encode_text(toks)
    s.prefill([1, 0])
    assert s.pos == hp["context_frames"]
    states = [o.new_state(enc[b], [1, 0][b]) for b in range(2)]
    rng = np.random.default_rng(42)
    frames = [np.full((2, 8), hp["audio_bos_id"], np.int32)] + [rng.integers(0, 2016, (2, 8)).astype(np.int32) for _ in range(6)]
    for f in frames:
        h = s.decoder_step(f)
        for b in range(2):
            close(h[b], states[b].step(f[b]), 1e-4)
    assert s.pos == hp["context_frames"] + len(frames)
    lg = s.final_proj()
    close(lg[0], o.final_proj(h[0]), 1e-4)


def test_lt_logits_and_greedy_f32(tiny_pair):
    m, o = tiny_pair
    s = m.session(batch=3, max_text=32)
    rng = np.random.default_rng(7)
    h = rng.standard_normal((3, o.hp["d_model"])).astype(np.float32)
    forced = rng.integers(0, 2016, (3, 8)).astype(np.int32)
    smp, am, lg = s.lt_sample(h, 0.0, 80, forbid_eos=[1, 0, 1], forced_codes=forced)
    for b in range(3):
        so, ao, lo = o.lt_sample(h[b], 0.0, 80, forbid_eos=bool([1, 0, 1][b]), forced_codes=forced[b])
        close(lg[b], lo, 1e-4)
        np.testing.assert_array_equal(am[b], ao)
        np.testing.assert_array_equal(smp[b], so)
    # free running: sampled codes are fed back
    smp2, am2, lg2 = s.lt_sample(h, 0.0, 80)
    for b in range(3):
        so, ao, lo = o.lt_sample(h[b], 0.0, 80)
        np.testing.assert_array_equal(smp2[b], so)
        close(lg2[b], lo, 1e-4)


def test_top_k_sampler_matches_oracle_given_uniforms(tiny_pair):
    m, o = tiny_pair
    s = m.session(batch=4, max_text=32)
    rng = np.random.default_rng(11)
    h = rng.standard_normal((4, o.hp["d_model"])).astype(np.float32)
    agree = total = 0
    for T, k in [(0.7, 80), (1.0, 5), (0.3, 1), (1.5, 2024)]:
        u = rng.random((4, 8)).astype(np.float32)
        forced = rng.integers(0, 2016, (4, 8)).astype(np.int32)   # same feedback on both sides
        smp, am, _ = s.lt_sample(h, T, k, forced_codes=forced, uniforms=u)
        for b in range(4):
            so, ao, _ = o.lt_sample(h[b], T, k, forced_codes=forced[b], uniforms=u[b])
            np.testing.assert_array_equal(am[b], ao)
            agree += int(np.sum(smp[b] == so)); total += 8
            if k == 1:
                np.testing.assert_array_equal(smp[b], ao)
    assert agree >= 0.97 * total       # expf ulp differences may flip a draw that lands on a CDF edge


@pytest.mark.parametrize("table", [False, True])
def test_generate_greedy_f32_matches_oracle(tiny_pair, table):
    m, o = tiny_pair
    set_table(m, o, table)
    s = m.session(batch=2, max_text=32, max_seq=10 + 24 + 16)
    toks = [HELLO, HELLO[:7] + [2379]]
    s.encode_text(toks, want_output=False)
    s.prefill([0, 1])
    out, hid = s.generate(max_steps=24, temperature=0.0, want_hidden=True)
    for b in range(2):
        ref, rh = o.synthesize(toks[b], speaker=[0, 1][b], temperature=0.0, max_steps=24, want_hidden=True)
        assert len(out[b]) == len(ref)
        assert np.mean(np.all(out[b] == ref, axis=1)) >= 0.99
        close(hid[b, :len(ref)], rh[:len(ref)], TABLE_TOL if table else 1e-4)
    set_table(m, o, False)


def test_teacher_forced_equals_stepwise(tiny_pair):
    m, o = tiny_pair
    hp = o.hp
    rng = np.random.default_rng(5)
    codes = rng.integers(0, 2016, (1, 10, 8)).astype(np.int32)
    s = m.session(batch=1, max_text=32)
    enc = s.encode_text([HELLO])
    s.prefill([0])
    hid, lg, gr = s.teacher_forced(codes)
    st = o.new_state(enc[0], 0)
    prev = np.full(8, hp["audio_bos_id"], np.int32)
    for t in range(10):
        h = st.step(prev)
        close(hid[0, t], h, 1e-4)
        so, ao, lo = o.lt_sample(h, 0.0, 80, forced_codes=codes[0, t])
        close(lg[0, t], lo, 1e-4)
        np.testing.assert_array_equal(gr[0, t], ao)
        prev = codes[0, t]


def test_error_paths(tiny_pair, B):
    m, _ = tiny_pair
    with pytest.raises(B.MagpieError):
        B.Model("/nonexistent.gguf")
    s = m.session(batch=1, max_text=8)
    with pytest.raises(B.MagpieError):
        s.encode_text([HELLO])                       # 16 tokens > max_text
    with pytest.raises(B.MagpieError):
        s.decoder_step(np.zeros((1, 8), np.int32))   # not prefilled
    s.encode_text([HELLO[:7] + [2379]])
    with pytest.raises(B.MagpieError):
        s.prefill([99])                              # speaker out of range


# ---- full Magpie-357M architecture --------------------------------------------------------------------

def _full_oracle(oracle_mod, full_model_path, table):
    o = oracle_mod.OracleModel(full_model_path)
    o.set_gelu_table(table)
    enc = o.encode_text(HELLO)
    rng = np.random.default_rng(42)
    codes = rng.integers(0, 2016, (12, 8)).astype(np.int32)     # config 2 stream, first frames
    st = o.new_state(enc, 0)
    prev = np.full(8, o.hp["audio_bos_id"], np.int32)
    hid, lgs, grs = [], [], []
    for t in range(12):
        h = st.step(prev)
        _, a, lg = o.lt_sample(h, 0.0, 80, forced_codes=codes[t])
        hid.append(h); lgs.append(lg); grs.append(a)
        prev = codes[t]
    return dict(o=o, enc=enc, codes=codes, hid=np.stack(hid), lg=np.stack(lgs), gr=np.stack(grs))


@pytest.fixture(scope="module")
def full_oracle(oracle_mod, full_model_path):
    return _full_oracle(oracle_mod, full_model_path, True)


@pytest.fixture(scope="module")
def full_oracle_notable(oracle_mod, full_model_path):
    return _full_oracle(oracle_mod, full_model_path, False)


@pytest.mark.parametrize("batch", [1, 3])       # batch 1 = megakernel path, batch 3 = per-op kernels
def test_full_model_teacher_forced_f32(B, full_model_path, full_oracle_notable, batch):
    fo = full_oracle_notable
    m = B.Model(full_model_path, 0, B.PREC_F32)
    m.set_gelu_f16(False)
    s = m.session(batch=batch, max_text=32)
    enc = s.encode_text([HELLO] * batch)
    close(enc[0], fo["enc"], 1e-4)
    s.prefill([0] * batch)
    hid, lg, gr = s.teacher_forced(np.repeat(fo["codes"][None], batch, axis=0))
    for b in range(batch):
        close(hid[b], fo["hid"], 1e-4)
        close(lg[b], fo["lg"], 1e-4)
        assert np.mean(np.all(gr[b] == fo["gr"], axis=1)) >= 0.99
    fp = s.final_proj()
    close(fp[0], fo["o"].final_proj(fo["hid"][-1]), 1e-4)


def test_full_model_teacher_forced_f32_reference_gelu_table(B, full_model_path, full_oracle):
    m = B.Model(full_model_path, 0, B.PREC_F32)        # product default: ggml-CPU f16 GELU table semantics
    s = m.session(batch=1, max_text=32)
    enc = s.encode_text([HELLO])
    close(enc[0], full_oracle["enc"], TABLE_TOL)
    s.prefill([0])
    hid, lg, gr = s.teacher_forced(full_oracle["codes"][None])
    close(hid[0], full_oracle["hid"], TABLE_TOL)
    close(lg[0], full_oracle["lg"], TABLE_TOL)
    assert np.mean(gr[0] == full_oracle["gr"]) >= 0.95


def test_full_model_teacher_forced_bf16(B, full_model_path, full_oracle):
    m = B.Model(full_model_path, 0, B.PREC_BF16)
    s = m.session(batch=2, max_text=32)
    enc = s.encode_text([HELLO, HELLO])
    close(enc[0], full_oracle["enc"], 2e-2)
    s.prefill([0, 0])
    codes = np.stack([full_oracle["codes"], full_oracle["codes"]])
    hid, lg, gr = s.teacher_forced(codes)
    for b in range(2):
        close(hid[b], full_oracle["hid"], 2e-2)
        close(lg[b], full_oracle["lg"], 2e-2)
    np.testing.assert_array_equal(gr[0], gr[1])          # batch rows are independent and deterministic
    # greedy agreement is a property of logit margins; with random-init weights the top-2 margin is
    # often below bf16 resolution, so the bf16 bar here is on logits (above), and codes are reported
    agree = np.mean(gr[0] == full_oracle["gr"])
    print(f"bf16 greedy code agreement vs f32 oracle: {agree:.3f}")


@pytest.mark.parametrize("kind,ggml_tol", [("model-f16", 1e-2), ("model-q8", 6e-2)])
def test_quantised_gguf_weights_match_oracle(B, fx, oracle_mod, kind, ggml_tol):
    """F16 and Q8_0 GGUF files (BASELINE configs[4]).  Two comparisons:
    (1) WEIGHT FORMAT: the oracle with the same exactly-dequantised weights and f32 activations (activation rounding off) --
        the product computes precisely that, so it is held to the plain bars: 1e-4 in f32 (GELU table off), 2e-2 in bf16, the
        latter with the Q8_0 matrices consumed as int8 + f16 scale inside the kernels where the path supports it;
    (2) ggml-CPU SEMANTICS: ggml's mul_mat additionally rounds the ACTIVATION row to f16 / Q8_0 blocks before each dot
        (SURVEY.md 8c item 4).  That is noise of the reference's CPU kernels, not of the file format; the distance to it is
        bounded at the noise level itself (f16: 2^-11 relative per element; Q8_0: 1/254 of the block maximum, ~5 % of the logit
        rms after the LT) and reported."""
    path = fx.ensure_fixture(kind)
    oracle_mod.set_activation_rounding(False)
    try:
        o = oracle_mod.OracleModel(path)
        o.set_gelu_table(False)
        enc_w = o.encode_text(HELLO)
        rng = np.random.default_rng(42)
        codes = rng.integers(0, 2016, (12, 8)).astype(np.int32)
        st = o.new_state(enc_w, 0)
        prev = np.full(8, o.hp["audio_bos_id"], np.int32)
        hid_w, lg_w = [], []
        for t in range(12):
            h = st.step(prev)
            _, _, lg = o.lt_sample(h, 0.0, 80, forced_codes=codes[t])
            hid_w.append(h); lg_w.append(lg); prev = codes[t]
        hid_w, lg_w = np.stack(hid_w), np.stack(lg_w)
    finally:
        oracle_mod.set_activation_rounding(True)
    fo = _full_oracle(oracle_mod, path, True)                 # full ggml-CPU semantics (activation rounding + GELU table)
    m = B.Model(path, 0, B.PREC_F32)
    m.set_gelu_f16(False)
    s = m.session(batch=2, max_text=32)
    enc = s.encode_text([HELLO, HELLO])
    close(enc[0], enc_w, 1e-4)
    s.prefill([0, 0])
    hid, lg, gr = s.teacher_forced(np.stack([codes, codes]))
    close(hid[0], hid_w, 1e-4)
    close(lg[0], lg_w, 1e-4)
    np.testing.assert_array_equal(gr[0], gr[1])
    m.set_gelu_f16(True)
    s.encode_text([HELLO, HELLO], want_output=False)
    s.prefill([0, 0])
    hid_t, lg_t, _ = s.teacher_forced(np.stack([fo["codes"], fo["codes"]]))
    close(hid_t[0], fo["hid"], ggml_tol)
    close(lg_t[0], fo["lg"], ggml_tol)
    d = np.abs(lg_t[0][np.isfinite(fo["lg"])] - fo["lg"][np.isfinite(fo["lg"])]).max() / np.sqrt(np.mean(fo["lg"][np.isfinite(fo["lg"])] ** 2))
    print(f"{kind}: max logit distance to the ggml-CPU semantics (activation rounding on) = {d:.3e} of the logit rms")
    s.close(); m.close()
    # the same file in bf16 compute, batch 1 (frame loop) and batched
    mb = B.Model(path, 0, B.PREC_BF16)
    for nb in (1, 3):
        sb = mb.session(batch=nb, max_text=32)
        sb.encode_text([HELLO] * nb, want_output=False)
        sb.prefill([0] * nb)
        hid_b, lg_b, _ = sb.teacher_forced(np.repeat(codes[None], nb, axis=0))
        close(hid_b[nb - 1], hid_w, 2e-2)
        close(lg_b[nb - 1], lg_w, 2e-2)
        sb.close()
    mb.close()


def _bf16_b1_run(m, codes, monkeypatch, env):
    for k in ("MGB_NO_LOOPK", "MGB_NO_MEGA", "MGB_LT_STREAM"):
        monkeypatch.delenv(k, raising=False)
    for k in env:
        monkeypatch.setenv(k, "1")
    s = m.session(batch=1, max_text=32)
    s.encode_text([HELLO], want_output=False)
    s.prefill([0])
    out = s.teacher_forced(codes)
    launches = s.last_loop_launches
    s.close()
    return out, launches


def test_bf16_fast_paths_match_per_op_kernels(B, full_model_path, full_oracle, monkeypatch):
    """batch-1 bf16, three implementations of the same frame loop: the persistent frame-loop kernel (benchmarked),
    the per-step megakernel + smem-resident LT, and the per-op kernels + streaming LT."""
    m = B.Model(full_model_path, 0, B.PREC_BF16)
    codes = full_oracle["codes"][None]
    (hid_l, lg_l, gr_l), n_l = _bf16_b1_run(m, codes, monkeypatch, [])
    (hid_f, lg_f, gr_f), n_f = _bf16_b1_run(m, codes, monkeypatch, ["MGB_NO_LOOPK"])
    (hid_s, lg_s, gr_s), n_s = _bf16_b1_run(m, codes, monkeypatch, ["MGB_NO_LOOPK", "MGB_NO_MEGA", "MGB_LT_STREAM"])
    assert n_l == 1 and n_f > n_l and n_s > n_f          # the three paths really are different kernels
    for hid, lg, gr in ((hid_l, lg_l, gr_l), (hid_f, lg_f, gr_f)):
        close(hid[0], hid_s[0], 2e-3)
        close(lg[0], lg_s[0], 2e-3)
        assert np.mean(gr == gr_s) >= 0.97
        close(hid[0], full_oracle["hid"], 2e-2)
        close(lg[0], full_oracle["lg"], 2e-2)
    # same bf16 weights, same f32 accumulation: the two fast paths differ only by summation order
    close(hid_l[0], hid_f[0], 2e-3)
    close(lg_l[0], lg_f[0], 2e-3)


def _bf16_b1_generate(m, monkeypatch, env, **kw):
    for k in ("MGB_NO_LOOPK", "MGB_NO_MEGA", "MGB_LT_STREAM"):
        monkeypatch.delenv(k, raising=False)
    for k in env:
        monkeypatch.setenv(k, "1")
    s = m.session(batch=1, max_text=32)
    s.encode_text([HELLO], want_output=False)
    s.prefill([0])
    out, hid = s.generate(want_hidden=True, **kw)
    pos = s.pos
    # the per-op step API must still work after the loop (state written back by the persistent kernel)
    h_next = s.decoder_step(None)
    s.close()
    return out[0], hid[0], pos, h_next[0]


def test_loop_kernel_generate_matches_stepwise_kernels(B, full_model_path, monkeypatch):
    m = B.Model(full_model_path, 0, B.PREC_BF16)
    # greedy, EOS honoured
    a, ha, pa, na = _bf16_b1_generate(m, monkeypatch, [], max_steps=40, temperature=0.0)
    b, hb, pb, nb = _bf16_b1_generate(m, monkeypatch, ["MGB_NO_LOOPK"], max_steps=40, temperature=0.0)
    n = min(len(a), len(b))
    assert n >= 4 and np.mean(np.all(a[:n] == b[:n], axis=1)) >= 0.9
    if np.array_equal(a[:n], b[:n]):
        assert len(a) == len(b)
        close(ha[:n], hb[:n], 2e-3)
    # greedy, fixed length; the follow-up decoder step sees the same state
    a, ha, pa, na = _bf16_b1_generate(m, monkeypatch, [], max_steps=16, temperature=0.0, ignore_eos=True)
    b, hb, pb, nb = _bf16_b1_generate(m, monkeypatch, ["MGB_NO_LOOPK"], max_steps=16, temperature=0.0, ignore_eos=True)
    assert len(a) == len(b) == 16 and pa == pb
    agree = np.mean(np.all(a == b, axis=1))
    assert agree >= 0.9
    if agree == 1.0:
        close(na, nb, 2e-3)
    # top-k sampling from the same uniforms.  The two paths' logits differ by ~1e-3 relative (bf16 weights, different
    # summation order), which is enough to move a draw across a CDF edge of the 80 near-equal candidates now and then, and
    # every later pick then follows a different trajectory.  So only the first pick of each run is compared (it depends
    # on the decoder hidden state alone), over six independent sets of uniforms; the sampler itself is pinned against the
    # oracle in test_top_k_sampler_matches_oracle_given_uniforms.
    same_first = 0
    for seed in range(3, 9):
        u = np.random.default_rng(seed).random((1, 4, 8)).astype(np.float32)
        a, _, _, _ = _bf16_b1_generate(m, monkeypatch, [], max_steps=4, temperature=0.7, top_k=80, uniforms=u, ignore_eos=True)
        b, _, _, _ = _bf16_b1_generate(m, monkeypatch, ["MGB_NO_LOOPK"], max_steps=4, temperature=0.7, top_k=80, uniforms=u, ignore_eos=True)
        assert len(a) == len(b) == 4
        same_first += int(a[0][0] == b[0][0])
    assert same_first >= 4


def test_loop_kernel_is_deterministic_and_chunkable(B, full_model_path, full_oracle, monkeypatch):
    """Two launches of T/2 frames == one launch of T frames (the state carried between launches is complete)."""
    for k in ("MGB_NO_LOOPK", "MGB_NO_MEGA", "MGB_LT_STREAM"):
        monkeypatch.delenv(k, raising=False)
    m = B.Model(full_model_path, 0, B.PREC_BF16)
    codes = full_oracle["codes"][None]
    s = m.session(batch=1, max_text=32)
    s.encode_text([HELLO], want_output=False); s.prefill([0])
    hid, lg, gr = s.teacher_forced(codes)
    s.encode_text([HELLO], want_output=False); s.prefill([0])
    hid2, lg2, gr2 = s.teacher_forced(codes)
    np.testing.assert_array_equal(hid, hid2)
    np.testing.assert_array_equal(lg, lg2)
    s.close()


def test_tensor_core_linear_matches_cuda_core_kernels(B, full_model_path, full_oracle, monkeypatch):
    """bf16, 24 utterances: the tcgen05 GEMM path (encoder, 110-frame prefill, batched decoder steps) against the
    CUDA-core linear kernels (MGB_NO_TC=1 at model load) and against the oracle."""
    codes = np.repeat(full_oracle["codes"][None, :6], 24, axis=0)

    def run(no_tc):
        if no_tc:
            monkeypatch.setenv("MGB_NO_TC", "1")
        else:
            monkeypatch.delenv("MGB_NO_TC", raising=False)
        m = B.Model(full_model_path, 0, B.PREC_BF16)
        monkeypatch.delenv("MGB_NO_TC", raising=False)
        s = m.session(batch=24, max_text=32)
        enc = s.encode_text([HELLO] * 24)
        s.prefill([0] * 24)
        hid, lg, gr = s.teacher_forced(codes)
        n = s.last_loop_launches
        s.close(); m.close()
        return enc, hid, lg, gr, n

    enc_t, hid_t, lg_t, gr_t, n_t = run(False)
    enc_c, hid_c, lg_c, gr_c, n_c = run(True)
    assert n_t > n_c                       # the tensor-core path adds one packing kernel per linear layer
    close(enc_t[0], enc_c[0], 2e-3)
    for b in (0, 7, 23):
        close(hid_t[b], hid_c[b], 4e-3)
        close(lg_t[b], lg_c[b], 4e-3)
        close(hid_t[b], full_oracle["hid"][:6], 2e-2)
        close(lg_t[b], full_oracle["lg"][:6], 2e-2)
    np.testing.assert_array_equal(gr_t[0], gr_t[23])     # batch rows are independent and deterministic


def test_batched_step_chain_modes_agree(B, full_model_path, full_oracle, monkeypatch):
    """bf16, 24 utterances x 6 teacher-forced steps: the default chain of the batched decoder step (ONE f16 activation image per GEMM
    operand against f16 twins of the bf16 weight images; LayerNorms folded through the QKV and FF1 GEMMs: FF2 / the folded cross-attention
    emit `x .* w` + row statistics)
    against (a) bf16 hi + lo image pairs (MGB_ACT_F16=0 at model load: f32-accurate activations), (b) the LayerNorm + pack launch in
    front of every QKV GEMM (MGB_NO_LNFOLD=1 at session creation: LayerNorm + pack launches, statistics exchange in the cross-attention kernel), and against the oracle.  f16 activations round at 2^-12 relative,
    an eighth of the bf16 weights' own step: the modes agree to a few 1e-3 of the rms, every mode keeps the 2e-2 bar."""
    codes = np.repeat(full_oracle["codes"][None, :6], 24, axis=0)

    def run(env):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        m = B.Model(full_model_path, 0, B.PREC_BF16)
        s = m.session(batch=24, max_text=32)
        for k in env:
            monkeypatch.delenv(k, raising=False)
        s.encode_text([HELLO] * 24, want_output=False)
        s.prefill([0] * 24)
        hid, lg, gr = s.teacher_forced(codes)
        n = s.last_loop_launches
        s.close(); m.close()
        return hid, lg, n

    hid_d, lg_d, n_d = run({})
    hid_a, lg_a, n_a = run({"MGB_ACT_F16": "0"})
    hid_l, lg_l, n_l = run({"MGB_NO_LNFOLD": "1"})
    assert n_l == n_d + 11 * 6 and n_a == n_d            # the fold removes the packing launch of layers 1..11, every step
    for b in (0, 11, 23):
        close(hid_d[b], hid_a[b], 3e-3)
        close(lg_d[b], lg_a[b], 3e-3)
        close(hid_d[b], hid_l[b], 3e-3)
        close(lg_d[b], lg_l[b], 3e-3)
        for hid, lg in ((hid_d, lg_d), (hid_a, lg_a), (hid_l, lg_l)):
            close(hid[b], full_oracle["hid"][:6], 2e-2)
            close(lg[b], full_oracle["lg"][:6], 2e-2)


def test_folded_cross_attention_matches_unfolded_kernels(B, full_model_path, full_oracle, monkeypatch):
    """bf16, 5 utterances with different texts: batched decoder steps with the folded cross-attention tables (one launch per
    layer) against the q_net GEMM + attention + o_net GEMM kernels (MGB_NO_XFOLD=1 at session creation) and the oracle."""
    texts = [HELLO, HELLO[:9] + [2379], HELLO, [2378, 5, 6, 2379], HELLO]
    codes = np.repeat(full_oracle["codes"][None, :6], 5, axis=0)
    m = B.Model(full_model_path, 0, B.PREC_BF16)

    def run(no_fold):
        if no_fold:
            monkeypatch.setenv("MGB_NO_XFOLD", "1")
        s = m.session(batch=5, max_text=32)
        monkeypatch.delenv("MGB_NO_XFOLD", raising=False)
        s.encode_text(texts, want_output=False)
        s.prefill([0, 1, 0, 2, 0])
        hid, lg, gr = s.teacher_forced(codes)
        n = s.last_loop_launches
        s.close()
        return hid, lg, gr, n

    hid_f, lg_f, gr_f, n_f = run(False)
    hid_u, lg_u, gr_u, n_u = run(True)
    assert n_f < n_u                                   # 4 launches fewer per layer and step
    for b in range(5):
        close(hid_f[b], hid_u[b], 2e-3)
        close(lg_f[b], lg_u[b], 2e-3)
    for b in (0, 2, 4):                                # the oracle's text and speaker
        close(hid_f[b], full_oracle["hid"][:6], 2e-2)
    m.close()


def test_long_kv_attention_key_split_matches_single_cta_and_oracle(B, oracle_mod, full_model_path, monkeypatch):
    """Decoder steps that reach >= 768 cached keys with few utterances divide each (head, utterance)'s keys over a CTA cluster
    (DSMEM merge in rank order) and scan them with a software-pipelined loop.  The long-KV kernels (cluster of 1, 2 and 3 CTAs,
    both occupancy builds, with and without the L2 prefetch, and the planner's own choice) must reproduce the one-CTA kernel
    and the oracle's hidden states at the long positions."""
    nb, T = 16, 700                                   # KV length 111 .. 810
    codes1 = np.random.default_rng(5).integers(0, 2016, (T, 8)).astype(np.int32)
    codes = np.repeat(codes1[None], nb, axis=0)

    def run(env):
        for k in ("MGB_NO_ATTN_SPLIT", "MGB_ATTN_SPLIT", "MGB_ATTN_OCC3", "MGB_ATTN_PF_MB"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        m = B.Model(full_model_path, 0, B.PREC_BF16)
        s = m.session(batch=nb, max_text=32, max_seq=110 + T + 16)
        s.encode_text([HELLO] * nb, want_output=False)
        s.prefill([0] * nb)
        hid, _, _ = s.teacher_forced(codes, want_logits=False, want_greedy=False)
        s.close(); m.close()
        return hid

    base = run({"MGB_NO_ATTN_SPLIT": "1"})
    for env in ({"MGB_ATTN_SPLIT": "1"}, {"MGB_ATTN_SPLIT": "2"}, {"MGB_ATTN_SPLIT": "3", "MGB_ATTN_PF_MB": "0"},
                {"MGB_ATTN_SPLIT": "3", "MGB_ATTN_OCC3": "1"}, {}):
        hid = run(env)
        close(hid[0, -200:], base[0, -200:], 2e-3)
        for b in (1, nb - 1):
            np.testing.assert_array_equal(hid[b], hid[0])        # equal rows, equal results (fixed merge order)
    o = oracle_mod.OracleModel(full_model_path)
    st = o.new_state(o.encode_text(HELLO), 0, 110 + T + 16)
    prev = np.full(8, o.hp["audio_bos_id"], np.int32)
    ref = []
    for t in range(T):
        ref.append(st.step(prev)); prev = codes1[t]
    close(hid[0, -8:], np.stack(ref[-8:]), 2e-2)
    close(base[0, -8:], np.stack(ref[-8:]), 2e-2)


@pytest.mark.parametrize("nb", [70, 130])
def test_large_batches_take_the_multi_tile_paths(B, full_model_path, full_oracle, nb):
    """More than 64 utterances: two 64-token tiles (70) / 128-token GEMM tiles and > 1 owned utterance per CTA in the batched
    local transformer (130).  Every row must reproduce the oracle (bf16 bar) and equal rows must give equal results."""
    m = B.Model(full_model_path, 0, B.PREC_BF16)
    s = m.session(batch=nb, max_text=32)
    s.encode_text([HELLO] * nb, want_output=False)
    s.prefill([0] * nb)
    codes = np.repeat(full_oracle["codes"][None, :4], nb, axis=0)
    hid, lg, gr = s.teacher_forced(codes)
    for b in (0, 63, 64, nb - 1):
        close(hid[b], full_oracle["hid"][:4], 2e-2)
        close(lg[b], full_oracle["lg"][:4], 2e-2)
    np.testing.assert_array_equal(gr[0], gr[nb - 1])
    s.close(); m.close()


def test_batched_generation_loop_with_both_local_transformer_kernels(B, full_model_path, monkeypatch):
    """Free-running generation of 18 utterances (different texts / speakers) in the device loop: the weight-stationary LT kernel
    (loop mode: device step counter, EOS bookkeeping, min-frames rule, Philox draws) against the per-utterance cluster kernel."""
    nb = 18
    rng = np.random.default_rng(5)
    texts = [[2378] + rng.integers(0, 90, int(rng.integers(3, 20))).tolist() + [2379] for _ in range(nb)]
    m = B.Model(full_model_path, 0, B.PREC_BF16)

    def run(no_batch, **kw):
        if no_batch:
            monkeypatch.setenv("MGB_NO_LT_BATCH", "1")
        s = m.session(batch=nb, max_text=32)
        s.encode_text(texts, want_output=False)
        s.prefill([b % 5 for b in range(nb)])
        out = s.generate(**kw)
        monkeypatch.delenv("MGB_NO_LT_BATCH", raising=False)
        s.close()
        return out

    a = run(False, max_steps=10, temperature=0.0, ignore_eos=True)
    b = run(True, max_steps=10, temperature=0.0, ignore_eos=True)
    assert all(len(x) == 10 for x in a) and all(len(x) == 10 for x in b)
    same = np.mean([np.array_equal(x[:4], y[:4]) for x, y in zip(a, b)])       # later frames follow the first differing pick
    assert same >= 0.8, same
    a = run(False, max_steps=12, temperature=0.0)                              # EOS honoured: lengths may be ragged
    b = run(True, max_steps=12, temperature=0.0)
    assert np.mean([len(x) == len(y) for x, y in zip(a, b)]) >= 0.8
    assert all(len(x) >= 4 for x in a)                                          # EOS is forbidden for the first 4 frames
    a = run(False, max_steps=6, temperature=0.7, top_k=80, seed=1234, ignore_eos=True)
    b = run(True, max_steps=6, temperature=0.7, top_k=80, seed=1234, ignore_eos=True)
    assert np.mean([x[0][0] == y[0][0] for x, y in zip(a, b)]) >= 0.8           # same Philox draw, same first pick
    m.close()


def test_batched_local_transformer_matches_per_utterance_kernel(B, full_model_path, full_oracle, monkeypatch):
    """bf16, 20 utterances: the cluster-resident batched LT (lt_cluster.cu, default from 4 utterances) against the
    one-cluster-per-utterance kernel (MGB_NO_LT_BATCH=1) and the oracle.  Same bf16 weights; the batched kernel multiplies on the
    tensor cores with the f32 activations split into bf16 hi + lo halves (2^-17 relative) and the f16 GELU table rounding of ggml
    can flip on such a difference, so the two kernels agree to a few 1e-4 of the logits' rms -- two orders below the bf16
    precision's own 2e-2 band against the oracle (measured: max abs diff 7.6e-5 at rms 0.38)."""
    nb = 20
    m = B.Model(full_model_path, 0, B.PREC_BF16)
    s = m.session(batch=nb, max_text=32)
    rng = np.random.default_rng(11)
    hid = np.stack([full_oracle["hid"][b % len(full_oracle["hid"])] for b in range(nb)]).astype(np.float32)
    hid[nb // 2:] += 0.05 * rng.standard_normal(hid[nb // 2:].shape).astype(np.float32)
    forced = rng.integers(0, 2016, (nb, 8)).astype(np.int32)
    u = rng.random((nb, 8)).astype(np.float32)

    def run(no_batch, **kw):
        if no_batch:
            monkeypatch.setenv("MGB_NO_LT_BATCH", "1")
        else:
            monkeypatch.delenv("MGB_NO_LT_BATCH", raising=False)
        out = s.lt_sample(hid, **kw)
        monkeypatch.delenv("MGB_NO_LT_BATCH", raising=False)
        return out

    # teacher-forced: logits of all 8 codebooks
    sa, aa, la = run(False, temperature=0.0, forced_codes=forced)
    sb, ab, lb = run(True, temperature=0.0, forced_codes=forced)
    close(la, lb, 3e-4)
    assert np.mean(aa == ab) >= 0.99
    # against the oracle for the rows that carry the oracle's own hidden states and forced codes
    o = full_oracle["o"]
    for b in (0, 3):
        _, a_ref, lg_ref = o.lt_sample(hid[b], 0.0, 80, forced_codes=forced[b])
        close(la[b], lg_ref, 2e-2)
    # free-running greedy and top-k sampling from given uniforms (forbid EOS on half of the rows)
    fe = (np.arange(nb) % 2).astype(np.uint8)
    sa, aa, _ = run(False, temperature=0.0, forbid_eos=fe, want_logits=False)
    sb, ab, _ = run(True, temperature=0.0, forbid_eos=fe, want_logits=False)
    assert np.mean(np.all(sa == sb, axis=1)) >= 0.9 and np.array_equal(sa, aa)
    sa, aa, _ = run(False, temperature=0.7, top_k=80, uniforms=u, want_logits=False)
    sb, ab, _ = run(True, temperature=0.7, top_k=80, uniforms=u, want_logits=False)
    assert np.mean(sa[:, 0] == sb[:, 0]) >= 0.9          # later picks follow the first differing draw
    s.close(); m.close()


# ---- nano-codec ---------------------------------------------------------------------------------------

def test_fsq_bit_exact(B, oracle_mod, codec_path):
    c = B.Codec(codec_path)
    idx = np.arange(2024, dtype=np.int32)
    codes = np.stack([np.roll(idx, 37 * cb) for cb in range(8)])
    lat = c.fsq_dequantize(codes)
    ref = oracle_mod.fsq_dequantize(codes)
    assert lat.dtype == np.float32
    np.testing.assert_array_equal(lat.view(np.uint32), ref.view(np.uint32))
    rng = np.random.default_rng(3)
    codes3 = rng.integers(0, 2016, (3, 8, 50)).astype(np.int32)
    lat3 = c.fsq_dequantize(codes3)
    for b in range(3):
        np.testing.assert_array_equal(lat3[b].view(np.uint32), oracle_mod.fsq_dequantize(codes3[b]).view(np.uint32))


def test_fsq_matches_reference_golden(B, codec_path):
    """GPU FSQ against tests/golden/ref_pieces.json, produced by running the reference's own fsq_dequantize_cpu
    (nano-codec.cpp:721-752) incl. out-of-range and negative indices: bit-exact."""
    import json
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_pieces.json"), encoding="utf-8"))
    idx = np.array(g["fsq"]["indices"], np.int32)
    codes = np.stack([np.roll(idx, g["fsq"]["roll_per_codebook"] * cb) for cb in range(8)]).astype(np.int32)
    bits = B.Codec(codec_path).fsq_dequantize(codes).view(np.uint32)
    np.testing.assert_array_equal(bits[:4], np.array(g["fsq"]["latent_bits_cb0"], np.uint32))
    assert int(np.bitwise_xor.reduce(bits.ravel() * np.arange(1, bits.size + 1, dtype=np.uint32))) == g["fsq"]["latent_bits_checksum"]


def test_codec_decode_snr(B, oracle_mod, codec_path):
    c = B.Codec(codec_path)
    o = oracle_mod.OracleCodec(codec_path, conv_f16=True)
    rng = np.random.default_rng(42)
    codes = rng.integers(0, 2016, (2, 8, 6)).astype(np.int32)
    pcm = c.decode(codes)
    assert pcm.shape == (2, 6 * 1024)
    for b in range(2):
        ref = o.decode(codes[b])
        s = snr_db(pcm[b], ref)
        assert s >= 40.0, f"codec SNR {s:.1f} dB"
        assert np.abs(pcm[b] - ref).max() < 1e-3
    # single-utterance API shape + causality (prefix of a longer decode)
    p1 = c.decode(codes[0][:, :3])
    np.testing.assert_allclose(p1, pcm[0][:3 * 1024], rtol=0, atol=1e-5)


def test_codec_tensor_core_path_matches_cuda_core_path_and_oracle(B, oracle_mod, codec_path, monkeypatch):
    """The tcgen05 implicit-GEMM pipeline (default) against the independent all-CUDA-core pipeline (MGB_CODEC_NO_TC=1)
    and the oracle: ragged length crossing several 128-step tiles in every stage, batch rows independent."""
    rng = np.random.default_rng(7)
    T = 21                                       # 168 / 1344 / 5376 / 10752 / 21504 steps per stage: partial tiles everywhere
    codes = rng.integers(0, 2016, (3, 8, T)).astype(np.int32)
    monkeypatch.delenv("MGB_CODEC_NO_TC", raising=False)
    c_tc = B.Codec(codec_path)
    pcm_tc = c_tc.decode(codes)
    n_tc = c_tc.last_launches
    monkeypatch.setenv("MGB_CODEC_NO_TC", "1")
    c_cc = B.Codec(codec_path)
    pcm_cc = c_cc.decode(codes)
    n_cc = c_cc.last_launches
    monkeypatch.delenv("MGB_CODEC_NO_TC", raising=False)
    assert n_tc != n_cc                          # really two different pipelines
    for b in range(3):
        assert snr_db(pcm_tc[b], pcm_cc[b]) >= 55.0
    ref = oracle_mod.OracleCodec(codec_path, conv_f16=True).decode(codes[1])
    assert snr_db(pcm_tc[1], ref) >= 40.0 and snr_db(pcm_cc[1], ref) >= 40.0
    assert np.abs(pcm_tc[1] - ref).max() < 2e-3
    # batch rows are independent: row 1 decoded alone gives the same samples
    np.testing.assert_allclose(c_tc.decode(codes[1]), pcm_tc[1], rtol=0, atol=1e-6)
    # a second call with a different geometry reuses the scratch images correctly (history rows re-zeroed)
    small = rng.integers(0, 2016, (2, 8, 3)).astype(np.int32)
    p_small = c_tc.decode(small)
    np.testing.assert_allclose(p_small[0], c_cc.decode(small[0]), rtol=0, atol=2e-3)
    assert snr_db(p_small[0], c_cc.decode(small[0])) >= 55.0


def test_codec_full_size_properties(B, oracle_mod, codec_path):
    """BASELINE config 3 at its full size (32 utterances x 1291 frames = 60 s each), checked through properties the causal
    codec offers: a prefix of the long decode equals the decode of the prefix (zero history, nano-codec.cpp:429-466), batch rows
    are independent, the head of the long decode matches the oracle, every sample is finite and inside tanh's range."""
    rng = np.random.default_rng(42)
    T = 1291
    codes = rng.integers(0, 2016, (32, 8, T)).astype(np.int32)
    c = B.Codec(codec_path)
    pcm = c.decode(codes)
    assert pcm.shape == (32, T * 1024)
    assert np.isfinite(pcm).all() and np.abs(pcm).max() <= 1.0
    for b, n in ((0, 40), (17, 129), (31, 1)):
        head = c.decode(np.ascontiguousarray(codes[b][:, :n]))
        np.testing.assert_allclose(head, pcm[b][:n * 1024], rtol=0, atol=1e-5)
    np.testing.assert_allclose(c.decode(codes[9]), pcm[9], rtol=0, atol=1e-6)
    ref = oracle_mod.OracleCodec(codec_path, conv_f16=True).decode(np.ascontiguousarray(codes[3][:, :5]))
    assert snr_db(pcm[3][:5 * 1024], ref) >= 40.0
    # rows differ (no aliasing of scratch images between utterances)
    assert np.abs(pcm[0] - pcm[1]).max() > 1e-3


def test_codec_ragged_lengths(B, oracle_mod, codec_path):
    c = B.Codec(codec_path)
    o = oracle_mod.OracleCodec(codec_path, conv_f16=True)
    rng = np.random.default_rng(1)
    for T in (1, 2, 33):
        codes = rng.integers(0, 2016, (8, T)).astype(np.int32)
        pcm = c.decode(codes)
        assert pcm.shape == (T * 1024,)
        if T <= 2:
            assert snr_db(pcm, o.decode(codes)) >= 40.0
