"""Batched decode at long KV lengths (BASELINE configs[4] per-GPU share: Q8_0 GGUF, 32 utterances) for profiling.

    python tools/long_step.py [warm_frames=2400] [probe_frames=2] [batch=32]

Runs `warm_frames` teacher-forced steps (KV length 110 -> 110 + warm_frames), prints the live step time of the first and the
last 200 of them, then brackets `probe_frames` more steps with cudaProfilerStart/Stop so that
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv python tools/long_step.py
lists exactly the launches of steps at the long KV length (tools/launch_summary.py X.csv summarises them).
"""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from magpie_tts_cpp_b200 import binding, fixtures

HELLO = [2378, 7, 4, 11, 11, 14, 32, 26, 22, 14, 17, 11, 3, 32, 28, 2379]
warm = int(sys.argv[1]) if len(sys.argv) > 1 else 2400
probe = int(sys.argv[2]) if len(sys.argv) > 2 else 2
B = int(sys.argv[3]) if len(sys.argv) > 3 else 32

m = binding.Model(fixtures.ensure_fixture("model-long-q8"), 0, binding.PREC_BF16)
s = m.session(batch=B, max_text=32, max_seq=m.hp["context_frames"] + warm + probe + 16)
rng = np.random.default_rng(42)
def codes(n):
    return np.repeat(rng.integers(0, 2016, (1, n, 8)).astype(np.int32), B, axis=0)

s.encode_text([HELLO] * B, want_output=False); s.prefill([b % 5 for b in range(B)])
kw = dict(want_hidden=False, want_logits=False)
s.teacher_forced(codes(8), **kw)                      # warm-up of the graph / allocations
s.teacher_forced(codes(200), **kw)
print("B=%d KV %4d..%4d: %.1f us/step" % (B, s.pos - 200, s.pos, s.last_loop_ms * 1e3 / 200))
rest = warm - 208 - 200
if rest > 0:
    s.teacher_forced(codes(rest), **kw)
s.teacher_forced(codes(200), **kw)
print("B=%d KV %4d..%4d: %.1f us/step, launches/step %.1f" % (B, s.pos - 200, s.pos, s.last_loop_ms * 1e3 / 200, s.last_loop_launches / 200))
kv_bytes = B * 36864.0 * s.pos
print("KV scan per step at this length: %.2f GB -> %.0f us at 6520.8 GB/s (+ 28 us of weights)" % (kv_bytes / 1e9, kv_bytes / 6520.8e3))

try:
    rt = ctypes.CDLL(os.environ.get("MGB_CUDART", "libcudart.so.12"))
except OSError:
    rt = ctypes.CDLL("/usr/local/cuda/lib64/libcudart.so.12")
rt.cudaProfilerStart()
s.teacher_forced(codes(probe), **kw)
rt.cudaProfilerStop()
print("probe: %.1f us/step over %d steps at KV %d" % (s.last_loop_ms * 1e3 / probe, probe, s.pos))
