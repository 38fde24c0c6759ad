"""Batched decode (64 utterances) for profiling: python tools/b64_step.py [frames] [random_texts: 0|1]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from magpie_tts_cpp_b200 import binding, fixtures
HELLO = [2378, 7, 4, 11, 11, 14, 32, 26, 22, 14, 17, 11, 3, 32, 28, 2379]
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 215
m = binding.Model(fixtures.ensure_fixture("model-f32"), 0, binding.PREC_BF16)
B = int(os.environ.get("B64_B", "64"))
RAND = len(sys.argv) > 2 and sys.argv[2] == "1"
rng = np.random.default_rng(7)
texts = [[2378] + rng.integers(0, 90, int(rng.integers(18, 79))).tolist() + [2379] for _ in range(B)] if RAND else [HELLO] * B
s = m.session(batch=B, max_text=96 if RAND else 32, max_seq=110 + 215 + 16)
codes = np.repeat(np.random.default_rng(42).integers(0, 2016, (1, frames, 8)).astype(np.int32), B, axis=0)
for _ in range(2):
    s.encode_text(texts, want_output=False); s.prefill([b % 5 for b in range(B)])
    s.teacher_forced(codes, want_hidden=False, want_logits=False)
print("B=%d: %.0f frames/s, %.1f us/step, launches/step %.1f" % (B, B * frames / (s.last_loop_ms * 1e-3), s.last_loop_ms * 1e3 / frames, s.last_loop_launches / frames))
