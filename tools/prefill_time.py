"""Wall time of encode_text + prefill (untimed by bench.py): python tools/prefill_time.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from magpie_tts_cpp_b200 import binding, fixtures
HELLO = [2378, 7, 4, 11, 11, 14, 32, 26, 22, 14, 17, 11, 3, 32, 28, 2379]
m = binding.Model(fixtures.ensure_fixture("model-f32"), 0, binding.PREC_BF16)
for B in (1, 8, 64):
    s = m.session(batch=B, max_text=32, max_seq=110 + 64 + 16)
    for it in range(4):
        t0 = time.perf_counter(); s.encode_text([HELLO] * B, want_output=False); t1 = time.perf_counter()
        s.prefill([0] * B); t2 = time.perf_counter()
    print("B=%d: encode_text %.2f ms, prefill %.2f ms (wall, last of 4)" % (B, (t1 - t0) * 1e3, (t2 - t1) * 1e3))
    s.close()
