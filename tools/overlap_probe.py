"""Probe: do two half-batches on ONE GPU (two sessions, two streams, two host threads) overlap their latency-bound kernels?
python tools/overlap_probe.py [utterances] [frames]   -- prints frames/s for 1 x B, 2 x B/2 and 4 x B/4 sessions on device 0."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from magpie_tts_cpp_b200 import binding, fixtures  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 215
texts, speakers = bench.config4_texts(B)
codes = np.repeat(bench.forced_codes(frames), B, axis=0)
for slots in (1, 2, 4):
    pool = binding.Pool(fixtures.ensure_fixture("model-f32"), devices=[0] * slots, precision=binding.PREC_BF16)
    for _ in range(2):
        pool.teacher_forced(texts, codes, speakers, want_greedy=False)
    t0 = time.perf_counter()
    pool.teacher_forced(texts, codes, speakers, want_greedy=False)
    wall = time.perf_counter() - t0
    ms = pool.last_device_ms
    print(f"{slots} session(s) x {B // slots} utterances: loop ms per session {np.round(ms, 1).tolist()}, "
          f"{B * frames / (ms.max() * 1e-3):.0f} frames/s by the slowest loop, wall {wall * 1e3:.0f} ms (incl. encode + prefill)", flush=True)
    pool.close()
