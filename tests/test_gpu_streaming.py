"""Batched streaming synthesis (mgb_stream_generate): BASELINE configs[4] asks for long-form STREAMING synthesis of many
utterances with a callback every 4 frames (reference: magpie_synthesize_sentence_streaming, magpie.cpp:4502-4829, one utterance,
one frame at a time, every chunk decoded with zero causal history, 4483-4500)."""
import numpy as np
import pytest

from test_gpu_parity import HELLO, snr_db

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def B():
    from magpie_tts_cpp_b200 import binding
    return binding


def collect(session, codec, **kw):
    audio, last = {}, {}

    def on_audio(u, pcm, frames_done, is_last):
        audio.setdefault(u, []).append(pcm)
        assert frames_done * 1024 == sum(len(x) for x in audio[u])
        last[u] = is_last
        return False

    nf = session.stream_generate(codec, on_audio, **kw)
    return {u: np.concatenate(v) for u, v in audio.items()}, nf, last


def test_streaming_three_utterances_f32_vs_oracle(B, oracle_mod, tiny_model_path, codec_path):
    """tiny architecture, f32: 3 utterances in lock step, chunks of 4 frames.  With 25 context frames the concatenated chunks of
    every utterance equal the whole-utterance decode of the oracle's greedy codes; with 0 (the reference's behaviour) every chunk
    equals the zero-history decode of its own 4 frames."""
    m = B.Model(tiny_model_path, 0, B.PREC_F32)
    c = B.Codec(codec_path)
    o = oracle_mod.OracleModel(tiny_model_path)
    oc = oracle_mod.OracleCodec(codec_path)
    utts = [HELLO, HELLO[:9] + [2379], HELLO[:5] + [2379]]
    spk = [0, 1, 0]
    refs = [o.synthesize(u, speaker=spk[i], temperature=0.0, max_steps=14) for i, u in enumerate(utts)]
    for ctx in (25, 0):
        s = m.session(batch=3, max_text=16)
        s.encode_text(utts, want_output=False)
        s.prefill(spk)
        audio, nf, last = collect(s, c, max_steps=14, frames_per_chunk=4, codec_context_frames=ctx, ignore_eos=True)
        s.close()
        assert list(nf) == [14, 14, 14] and all(last.values())
        for i in range(3):
            assert len(refs[i]) == 14                                   # (no EOS within 14 frames for these utterances)
            if ctx:
                whole = oc.decode(np.ascontiguousarray(refs[i].T))
                assert snr_db(audio[i], whole) >= 40.0
            else:
                parts = [oc.decode(np.ascontiguousarray(refs[i][t:t + 4].T)) for t in range(0, 14, 4)]
                assert snr_db(audio[i], np.concatenate(parts)) >= 40.0
    m.close()


def test_streaming_batch1_persistent_kernel_and_batch16(B, full_model_path, codec_path):
    """Magpie-357M bf16.  Batch 1 streams from the persistent frame-loop kernel (one launch per 4-frame chunk): same codes as
    mgb_generate, audio == one decode of all of them.  16 utterances stream from the graph-replayed batched step: frame counts
    follow mgb_generate (+1 for the EOS frame, which the streaming path decodes as the reference's does)."""
    m = B.Model(full_model_path, 0, B.PREC_BF16)
    c = B.Codec(codec_path)
    s = m.session(batch=1, max_text=32, max_seq=110 + 40 + 16)
    s.encode_text([HELLO], want_output=False); s.prefill([0])
    codes = s.generate(max_steps=40, temperature=0.0, ignore_eos=True)[0]
    s.encode_text([HELLO], want_output=False); s.prefill([0])
    audio, nf, _ = collect(s, c, max_steps=40, frames_per_chunk=4, codec_context_frames=25, ignore_eos=True)
    assert nf[0] == 40 and len(audio[0]) == 40 * 1024
    whole = c.decode(np.ascontiguousarray(codes.T))
    assert snr_db(audio[0], whole) >= 40.0
    s.close()
    nb = 16
    rng = np.random.default_rng(5)
    texts = [[2378] + rng.integers(0, 90, int(rng.integers(3, 20))).tolist() + [2379] for _ in range(nb)]
    s = m.session(batch=nb, max_text=32, max_seq=110 + 24 + 16)
    s.encode_text(texts, want_output=False); s.prefill([b % 5 for b in range(nb)])
    gen = s.generate(max_steps=24, temperature=0.0)
    s.encode_text(texts, want_output=False); s.prefill([b % 5 for b in range(nb)])
    audio, nf, last = collect(s, c, max_steps=24, frames_per_chunk=4, codec_context_frames=0)
    for b in range(nb):
        want = len(gen[b]) + (1 if len(gen[b]) < 24 else 0)
        assert nf[b] == want and len(audio[b]) == want * 1024 and last[b]
        assert np.isfinite(audio[b]).all() and np.abs(audio[b]).max() <= 1.0
    s.close(); m.close()
