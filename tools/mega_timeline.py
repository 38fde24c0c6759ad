"""Profiling aid: per-phase cycle breakdown of the batch-1 decoder megakernel (CTA 0), run on a GPU box:
   MGB_MEGA_DBG=1 python tools/mega_timeline.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["MGB_MEGA_DBG"] = "1"
from magpie_tts_cpp_b200 import binding, fixtures

HELLO = [2378, 7, 4, 11, 11, 14, 32, 26, 22, 14, 17, 11, 3, 32, 28, 2379]
m = binding.Model(fixtures.ensure_fixture("model-f32"), 0, binding.PREC_BF16)
s = m.session(batch=1, max_text=32)
s.encode_text([HELLO], want_output=False)
s.prefill([0])
codes = np.random.default_rng(0).integers(0, 2016, (1, 8)).astype(np.int32)
for _ in range(300):
    s.decoder_step(codes, want_hidden=False)
L = m.hp["dec_layers"]
st = s.debug_stamps(2 + 14 * L).astype(np.int64)
d = np.diff(st)
print("total cycles", st[-1] - st[0], "=", (st[-1] - st[0]) / 1.965e3, "us @1.965GHz")
print("start -> end of P1 (layer 0, incl. embedding):", d[0])
names = ["B1 wait", "P2 attn", "B2 wait", "P3 comb+o", "B3 wait", "P4 ln+xq", "B4 wait", "P5 xattn+xo", "B5 wait", "P6 ln+ff1", "B6 wait", "P7 ff2", "B7 wait", "P1 ln+qkv (next layer)"]
per = d[1:1 + 14 * L].reshape(L, 14)
for i, n in enumerate(names):
    print(f"{n:12s} mean {per[:, i].mean():8.0f} cyc  ({per[:, i].mean()/1965:6.2f} us)  min {per[:, i].min():6d} max {per[:, i].max():6d}")
print("per layer cycles:", per.sum(1))
import time
t0 = time.perf_counter()
for _ in range(200):
    s.decoder_step(None, want_hidden=False)
print("wall per decoder_step API call (sync each):", (time.perf_counter() - t0) / 200 * 1e6, "us")

# per-CTA arrival times (globaltimer, ns) at every barrier of the last step
G = 148
nb = 7 * L
arr = s.debug_stamps(1024 + nb * G)[1024:].astype(np.int64).reshape(nb, G)
print("\nper-barrier arrival spread (ns): barrier index within layer, mean over layers 1..L-1")
for j in range(7):
    rows = arr[[l * 7 + j for l in range(1, L)]]
    first = rows.min(1, keepdims=True)
    lat = rows - first
    worst = lat.argmax(1)
    print(f"B{j+1}: spread mean {lat.max(1).mean():8.0f} ns; median CTA lateness {np.median(lat,1).mean():7.0f}; late CTAs (top) {np.bincount(worst, minlength=G).argsort()[-5:][::-1]}")
    if j == 0:
        late = lat.mean(0)
        print("   mean lateness by CTA id (every 8th):", [int(x) for x in late[::8]])
prev = arr[6:-1:7]   # B7 of layer l-1
b1 = arr[7::7]
print("time from last CTA arriving at B7(l-1) to each CTA arriving at B1(l), mean over layers (ns), every 8th CTA:")
dt = (b1 - prev.max(1, keepdims=True)).mean(0)
print([int(x) for x in dt[::8]])

# whole-frame loop time (decoder megakernel + LT kernel + step counter), CUDA-graph replay
codes_tf = np.random.default_rng(1).integers(0, 2016, (1, 200, 8)).astype(np.int32)
s.encode_text([HELLO], want_output=False); s.prefill([0])
s.teacher_forced(codes_tf, want_hidden=False, want_logits=False)
s.encode_text([HELLO], want_output=False); s.prefill([0])
s.teacher_forced(codes_tf, want_hidden=False, want_logits=False)
print("frame loop: %.1f us/frame over 200 frames (graph replay)" % (s.last_loop_ms * 1e3 / 200))
