import os, sys
import numpy as np
sys.path.insert(0, '/root/repo')
from magpie_tts_cpp_b200 import binding, fixtures
HELLO = [2378, 7, 4, 11, 11, 14, 32, 26, 22, 14, 17, 11, 3, 32, 28, 2379]
which = sys.argv[1]
if which == "codec":
    c = binding.Codec(fixtures.ensure_fixture("codec-f32"), 0)
    codes = np.random.default_rng(1).integers(0, 2016, (2, 8, 5)).astype(np.int32)
    pcm = c.decode(codes); print("codec ok", pcm.shape, float(np.abs(pcm).max()))
else:
    m = binding.Model(fixtures.ensure_fixture("model-f32"), 0, binding.PREC_BF16)
    B = 20
    rng = np.random.default_rng(7)
    texts = [[2378] + rng.integers(0, 90, int(rng.integers(3, 28))).tolist() + [2379] for _ in range(B)]
    s = m.session(batch=B, max_text=32, max_seq=110 + 8 + 16)
    s.encode_text(texts, want_output=False); s.prefill([b % 5 for b in range(B)])
    codes = np.repeat(rng.integers(0, 2016, (1, 3, 8)).astype(np.int32), B, axis=0)
    hid, lg, gr = s.teacher_forced(codes); print("batch step ok", hid.shape, bool(np.isfinite(hid).all()))
    s1 = m.session(batch=1, max_text=32, max_seq=110 + 8 + 16)
    s1.encode_text([HELLO], want_output=False); s1.prefill([0])
    out = s1.generate(max_steps=3, temperature=0.0, ignore_eos=True); print("loop ok", out[0].shape)
