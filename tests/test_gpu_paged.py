"""Paged decoder self-attention cache and continuous batching (north_star (a): "GPU-resident paged KV cache"; the reference
allocates one contiguous cache per call, magpie.cpp:3315-3376, and synthesises one utterance at a time)."""
import numpy as np
import pytest

from test_gpu_parity import HELLO, close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def B():
    from magpie_tts_cpp_b200 import binding
    return binding


def test_paged_cache_with_interleaved_pages_matches_contiguous_layout(B, full_model_path):
    """bf16, 5 utterances x 300 teacher-forced frames (cache positions 0..410 = 4 pages each).  A pool of 20 pages handed out on
    demand gives every utterance NON-contiguous, interleaved pages; the attention kernels must read the same keys through the
    page table as the default session does from its up-front reservation: bit-identical hidden states."""
    nb, T = 5, 300
    rng = np.random.default_rng(9)
    texts = [[2378] + rng.integers(0, 90, int(rng.integers(5, 28))).tolist() + [2379] for _ in range(nb)]
    codes = rng.integers(0, 2016, (nb, T, 8)).astype(np.int32)
    m = B.Model(full_model_path, 0, B.PREC_BF16)

    def run(kv_pages):
        s = m.session(batch=nb, max_text=32, max_seq=110 + T + 16, kv_pages=kv_pages)
        s.encode_text(texts, want_output=False)
        s.prefill([b % 5 for b in range(nb)])
        total, used0 = s.kv_pages
        hid, _, gr = s.teacher_forced(codes, want_logits=False)
        total, used1 = s.kv_pages
        s.close()
        return hid, gr, total, used0, used1

    hid_a, gr_a, tot_a, u0_a, u1_a = run(0)
    hid_b, gr_b, tot_b, u0_b, u1_b = run(20)
    assert tot_a == nb * 4 and u0_a == u1_a == tot_a            # default: ceil(426 / 128) = 4 pages per utterance, all reserved
    assert tot_b == 20 and u0_b == nb and u1_b == nb * 4          # on demand: one page each after the prefill, four at the end
    np.testing.assert_array_equal(hid_a, hid_b)
    np.testing.assert_array_equal(gr_a, gr_b)
    # a pool that cannot hold the run fails loudly instead of overwriting another utterance's rows
    s = m.session(batch=nb, max_text=32, max_seq=110 + T + 16, kv_pages=12)
    s.encode_text(texts, want_output=False)
    s.prefill([0] * nb)
    with pytest.raises(B.MagpieError, match="page pool exhausted"):
        s.teacher_forced(codes, want_logits=False)
    s.close(); m.close()


def test_continuous_batching_matches_oracle_f32(B, oracle_mod, tiny_model_path):
    """f32, tiny architecture: 11 utterances with ragged frame limits through a 4-slot session.  Every utterance must equal the
    oracle's greedy synthesis (EOS rule included), and the loop must run about sum(frames) / 4 steps."""
    m = B.Model(tiny_model_path, 0, B.PREC_F32)
    o = oracle_mod.OracleModel(tiny_model_path)
    rng = np.random.default_rng(21)
    utts = [HELLO[:int(rng.integers(4, 15))] + [2379] for _ in range(11)]
    spk = [int(rng.integers(0, 2)) for _ in range(11)]
    limits = [6, 22, 9, 24, 5, 17, 24, 8, 12, 20, 7]
    s = m.session(batch=4, max_text=16, max_seq=10 + 24 + 16, kv_pages=4)
    out, steps = s.generate_queue(utts, speakers=spk, max_steps=24, max_steps_per_utt=limits)
    total = 0
    for i, u in enumerate(utts):
        ref = o.synthesize(u, speaker=spk[i], temperature=0.0, max_steps=limits[i])
        assert out[i].shape == ref.shape, (i, out[i].shape, ref.shape)
        assert np.mean(np.all(out[i] == ref, axis=1)) >= 0.99 if len(ref) else True
        total += len(ref)
    # a retired slot is noticed at the next poll (every 8 steps) and refilled
    assert steps <= total / 4 + 8 * (len(utts) / 4 + 2), (steps, total)
    assert steps < sum(sorted(limits)[-3:]) + 24 * 3            # far fewer than ceil(11 / 4) rounds of the longest utterance
    assert s.kv_pages[1] <= 4
    s.close(); m.close()


def test_continuous_batching_bf16_equals_plain_batches(B, full_model_path):
    """bf16, Magpie-357M: 40 utterances (texts of 6..60 tokens, ragged limits 5..60 frames) through 16 slots == the same
    utterances generated in plain 16-wide batches, with a page pool a quarter the size of the up-front reservation."""
    n, nb, T = 40, 16, 60
    rng = np.random.default_rng(33)
    utts = [[2378] + rng.integers(0, 90, int(rng.integers(4, 59))).tolist() + [2379] for _ in range(n)]
    spk = [int(rng.integers(0, 5)) for _ in range(n)]
    limits = [int(x) for x in rng.integers(5, T + 1, n)]
    m = B.Model(full_model_path, 0, B.PREC_BF16)
    sq = m.session(batch=nb, max_text=64, max_seq=110 + T + 16, kv_pages=nb * 2)
    out, steps = sq.generate_queue(utts, speakers=spk, max_steps=T, max_steps_per_utt=limits)
    sq.close()
    frames = sum(len(x) for x in out)
    ref = [None] * n
    plain_steps = 0
    sp = m.session(batch=nb, max_text=64, max_seq=110 + T + 16)
    for i0 in range(0, 48, nb):
        idx = [min(i, n - 1) for i in range(i0, i0 + nb)]           # the last batch is padded with copies of the last utterance
        sp.encode_text([utts[i] for i in idx], want_output=False)
        sp.prefill([spk[i] for i in idx])
        g = sp.generate(max_steps=T, temperature=0.0)
        for j, i in enumerate(idx):
            ref[i] = g[j][:limits[i]]
        plain_steps += max(len(ref[i]) for i in idx)               # a plain batch runs until its longest utterance is done
    sp.close()
    # Rows of a batch are independent, but a refill encodes / prefills 1-3 utterances at a time: fewer than 16 tokens take the
    # CUDA-core linear kernels instead of the tcgen05 GEMMs (another summation order), so a refilled utterance's cache can differ
    # from the plain batch's in the last bf16 bit and a greedy near-tie may flip now and then.  Hence: (almost) all identical.
    same = sum(int(len(out[i]) == len(ref[i]) and np.array_equal(out[i], ref[i])) for i in range(n))
    first = sum(int(len(out[i]) > 0 and len(ref[i]) > 0 and np.array_equal(out[i][0], ref[i][0])) for i in range(n))
    assert same >= 0.9 * n, f"{n - same} of {n} utterances differ between the queue and the plain batches"
    assert first >= 0.95 * n
    # ideal = sum(frames) / slots; on top of it: a finished slot is noticed at the next poll (every 8 steps), and the last
    # utterances drain the queue at different times (at most half the longest utterance on average)
    print(f"continuous batching: {steps} steps for {frames} frames in {nb} slots (ideal {frames / nb:.0f}; plain 16-wide batches {plain_steps})")
    assert steps <= frames / nb + 8 * (n / nb + 2) + T / 2, (steps, frames)
    assert steps < 0.8 * plain_steps, (steps, plain_steps)
    m.close()
