"""Profiling aid: clock64 stamps inside phase P5 (LN -> FFN1) of every layer, CTA 0, last frame.
   MGB_LOOP_FLAGS=4 python tools/loop_detail.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["MGB_LOOP_DBG"] = "1"
os.environ["MGB_LOOP_FLAGS"] = str(int(os.environ.get("MGB_LOOP_FLAGS", "0")) | 4)
from magpie_tts_cpp_b200 import binding, fixtures
HELLO = [2378, 7, 4, 11, 11, 14, 32, 26, 22, 14, 17, 11, 3, 32, 28, 2379]
T = 100
m = binding.Model(fixtures.ensure_fixture("model-f32"), 0, binding.PREC_BF16)
s = m.session(batch=1, max_text=32)
codes = np.random.default_rng(1).integers(0, 2016, (1, T, 8)).astype(np.int32)
for _ in range(2):
    s.encode_text([HELLO], want_output=False); s.prefill([0])
    s.teacher_forced(codes, want_hidden=False, want_logits=False)
print("flags=%s: %.1f us/frame" % (os.environ["MGB_LOOP_FLAGS"], s.last_loop_ms * 1e3 / T))
L = m.hp["dec_layers"]
st = s.debug_stamps(9 * L).astype(np.int64).reshape(L, 9)
d = np.diff(st, axis=1)
names = ["poll", "cbar", "layer_norm", "cbar", "ring_wait", "gemv_part", "release+cbar", "emit"]
for c, nm in enumerate(names):
    print(f"  {nm:14s} mean {d[:, c].mean():8.0f} cyc   min {d[:, c].min():6d} max {d[:, c].max():6d}")
print("  P5 total mean %.0f cyc; layer-to-layer (P5 start to next P5 start) mean %.0f cyc" % ((st[:, 8] - st[:, 0]).mean(), np.diff(st[:, 0]).mean()))
