"""Wall time of model / codec loading (GGUF parse, dequantise, upload, table building, tile packing): python tools/load_time.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from magpie_tts_cpp_b200 import binding, fixtures
import torch
torch.cuda.init(); torch.zeros(1, device="cuda")
for kind in ("model-f32", "model-q8"):
    path = fixtures.ensure_fixture(kind)
    open(path, "rb").read()          # page cache
    for prec, name in ((binding.PREC_BF16, "bf16"), (binding.PREC_F32, "f32")):
        t0 = time.perf_counter(); m = binding.Model(path, 0, prec); dt = time.perf_counter() - t0
        print(f"{kind} -> {name}: {dt * 1e3:.0f} ms ({os.path.getsize(path) / 1e6:.0f} MB file)", flush=True)
        m.close()
cp = fixtures.ensure_fixture("codec-f32"); open(cp, "rb").read()
t0 = time.perf_counter(); c = binding.Codec(cp, 0); dt = time.perf_counter() - t0
print(f"codec load: {dt * 1e3:.0f} ms")
import numpy as np
codes = np.random.default_rng(0).integers(0, 2016, (8, 32)).astype(np.int32)
t0 = time.perf_counter(); c.decode(codes); dt = time.perf_counter() - t0
print(f"first codec decode (weight repack, scratch allocation): {dt * 1e3:.0f} ms")
t0 = time.perf_counter(); c.decode(codes); dt = time.perf_counter() - t0
print(f"second codec decode: {dt * 1e3:.1f} ms")
