"""Encode + prefill once more after a warm-up pass, for an ncu launch list: python tools/prefill_profile.py [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from magpie_tts_cpp_b200 import binding, fixtures
HELLO = [2378, 7, 4, 11, 11, 14, 32, 26, 22, 14, 17, 11, 3, 32, 28, 2379]
m = binding.Model(fixtures.ensure_fixture("model-f32"), 0, binding.PREC_BF16)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
s = m.session(batch=B, max_text=32, max_seq=110 + 16 + 16)
for _ in range(2):
    s.encode_text([HELLO] * B, want_output=False); s.prefill([0] * B)
