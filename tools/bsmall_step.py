"""Batched decode at small batch sizes: python tools/bsmall_step.py B [frames]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from magpie_tts_cpp_b200 import binding, fixtures
import bench
B = int(sys.argv[1]); frames = int(sys.argv[2]) if len(sys.argv) > 2 else 100
m = binding.Model(fixtures.ensure_fixture("model-f32"), 0, binding.PREC_BF16)
texts, spk = bench.config4_texts(B)
s = m.session(batch=B, max_text=96, max_seq=110 + frames + 16)
codes = np.repeat(bench.forced_codes(frames), B, axis=0)
for _ in range(2):
    s.encode_text(texts, want_output=False); s.prefill(spk)
    s.teacher_forced(codes, want_hidden=False, want_logits=False)
print("B=%d: %.0f frames/s, %.1f us/step, launches/step %.1f" % (B, B * frames / (s.last_loop_ms * 1e-3), s.last_loop_ms * 1e3 / frames, s.last_loop_launches / frames))
