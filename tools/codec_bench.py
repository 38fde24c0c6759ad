"""Codec throughput: python tools/codec_bench.py [B] [T]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from magpie_tts_cpp_b200 import binding, fixtures
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
T = int(sys.argv[2]) if len(sys.argv) > 2 else 128
c = binding.Codec(fixtures.ensure_fixture("codec-f32"), 0)
codes = np.random.default_rng(42).integers(0, 2016, (B, 8, T)).astype(np.int32)
for _ in range(3):
    pcm = c.decode(codes)
print("B=%d T=%d: %.2f ms device, %.0f audio-s/s, %.1f TFLOP/s, %d launches" % (B, T, c.last_ms, B * T * 1024 / 22050.0 / (c.last_ms * 1e-3),
      B * T * 2.447e9 / (c.last_ms * 1e-3) / 1e12, c.last_launches))
if os.environ.get("MGB_CODEC_NO_TC") is None and B * T <= 64:
    os.environ["MGB_CODEC_NO_TC"] = "1"
    c2 = binding.Codec(fixtures.ensure_fixture("codec-f32"), 0)
    ref = c2.decode(codes)
    err = np.abs(pcm - ref).max()
    snr = 10 * np.log10(np.sum(ref.astype(np.float64) ** 2) / max(np.sum((pcm - ref).astype(np.float64) ** 2), 1e-300))
    print("vs CUDA-core conv: max abs diff %.3e, SNR %.1f dB" % (err, snr))
