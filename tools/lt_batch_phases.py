import os, sys
import numpy as np
sys.path.insert(0, '/root/repo')
os.environ.setdefault("MGB_LT_DBG", "1")
from magpie_tts_cpp_b200 import binding, fixtures
m = binding.Model(fixtures.ensure_fixture("model-f32"), 0, binding.PREC_BF16)
B=int(os.environ.get("B64_B", "64"))
s = m.session(batch=B, max_text=32)
hid = np.random.default_rng(0).standard_normal((B,768)).astype(np.float32)
TEMP = float(os.environ.get("LT_TEMP", "0.0"))      # > 0: top-k sampling in the owner phase
for i in range(3): s.lt_sample(hid, temperature=TEMP, want_logits=False)
os.environ["MGB_LT_DBG_DUMP"]="1"
s.lt_sample(hid, temperature=TEMP, want_logits=False)
