import os, sys
import numpy as np
sys.path.insert(0, "/root/repo")
from magpie_tts_cpp_b200 import binding, fixtures
m = binding.Model(fixtures.ensure_fixture("model-f32"), 0, binding.PREC_BF16)
B = 64
rng = np.random.default_rng(7)
texts = [[2378] + rng.integers(0, 90, int(rng.integers(18, 79))).tolist() + [2379] for _ in range(B)]
s = m.session(batch=B, max_text=96, max_seq=110 + 40 + 16)
for temp in (0.0, 0.7):
    for _ in range(2):
        s.encode_text(texts, want_output=False); s.prefill([b % 5 for b in range(B)])
        out = s.generate(max_steps=40, temperature=temp, top_k=80, ignore_eos=True, seed=1)
    print("generate B=64 temp %.1f: %.1f us/step (%d frames)" % (temp, s.last_loop_ms * 1e3 / 40, len(out[0])))
s1 = m.session(batch=1, max_text=32, max_seq=110 + 200 + 16)
HELLO = [2378, 7, 4, 11, 11, 14, 32, 26, 22, 14, 17, 11, 3, 32, 28, 2379]
for temp in (0.0, 0.7):
    for _ in range(2):
        s1.encode_text([HELLO], want_output=False); s1.prefill([0])
        out = s1.generate(max_steps=200, temperature=temp, top_k=80, ignore_eos=True, seed=1)
    print("generate B=1 temp %.1f: %.1f us/frame (%d frames)" % (temp, s1.last_loop_ms * 1e3 / 200, len(out[0])))
