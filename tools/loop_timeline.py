"""Profiling aid: per-phase timeline (globaltimer, CTA 0, last frame) of the batch-1 persistent frame-loop kernel.
Every phase has two stamps: inputs arrived (poll done) and outputs emitted.
   [MGB_LOOP_FLAGS=n] python tools/loop_timeline.py [frames]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["MGB_LOOP_DBG"] = "1"
from magpie_tts_cpp_b200 import binding, fixtures

HELLO = [2378, 7, 4, 11, 11, 14, 32, 26, 22, 14, 17, 11, 3, 32, 28, 2379]
T = int(sys.argv[1]) if len(sys.argv) > 1 else 200
m = binding.Model(fixtures.ensure_fixture("model-f32"), 0, binding.PREC_BF16)
s = m.session(batch=1, max_text=32)
codes = np.random.default_rng(1).integers(0, 2016, (1, T, 8)).astype(np.int32)
for _ in range(2):
    s.encode_text([HELLO], want_output=False); s.prefill([0])
    s.teacher_forced(codes, want_hidden=False, want_logits=False)
print("flags=%s frame loop: %.1f us/frame over %d frames (%d launches)" % (os.environ.get("MGB_LOOP_FLAGS", "0"), s.last_loop_ms * 1e3 / T, T, s.last_loop_launches))
L = m.hp["dec_layers"]
n = 1 + 12 * L + 2 + 12 * 8
st = s.debug_stamps(n).astype(np.int64)
d = np.diff(st)
print("last frame total: %.1f us" % ((st[-1] - st[0]) / 1e3))
dec = d[:12 * L].reshape(L, 6, 2)
names = ["P1 ln+qkv", "P2 attn", "P3 comb+o", "P4 xattn", "P5 ln+ff1", "P6 ff2"]
print("  phase        wait(ns)  work(ns)   [mean over layers; wait = previous emit -> inputs arrived, work = -> own emit]")
for i, nm in enumerate(names):
    print(f"  {nm:10s} {dec[1:, i, 0].mean():9.0f} {dec[1:, i, 1].mean():9.0f}    max wait {dec[1:, i, 0].max():6d} max work {dec[1:, i, 1].max():6d}")
print("  per layer ns:", dec.sum((1, 2)).tolist(), " decoder total %.1f us (wait %.1f, work %.1f)" % (dec.sum() / 1e3, dec[:, :, 0].sum() / 1e3, dec[:, :, 1].sum() / 1e3))
print("  final LN + hidden: wait %d work %d ns" % (d[12 * L], d[12 * L + 1]))
lt = d[12 * L + 2:12 * L + 2 + 96].reshape(8, 6, 2)
for i, nm in enumerate(["A ln+qkv", "B attn+o", "C ln+ff1", "D ff2", "E out-proj", "F argmax"]):
    print(f"  LT {nm:10s} {lt[:, i, 0].mean():9.0f} {lt[:, i, 1].mean():9.0f}    max wait {lt[:, i, 0].max():6d} max work {lt[:, i, 1].max():6d}")
print("  LT total %.1f us (wait %.1f, work %.1f)" % (lt.sum() / 1e3, lt[:, :, 0].sum() / 1e3, lt[:, :, 1].sum() / 1e3))
