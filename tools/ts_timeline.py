"""In-kernel timeline of the token-stationary GEMMs of the batched decoder step (gemm_ts.cu, MGB_TS_DBG): CTA (0, 0) of every launch
stamps globaltimer at entry | dependency wait over | first stage landed | last MMA issued | accumulator complete | epilogue done.
One teacher-forced run records the stamps during the CUDA-graph replays; a direct decoder step afterwards triggers the dump of the
last 160 launches (40 frames x 48 GEMM launches recorded; the ring keeps 1024).
    python tools/ts_timeline.py [utterances]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["MGB_TS_DBG"] = "1"
os.environ["MGB_TS_DBG_DUMP"] = "49"
os.environ["MGB_ATTN_DBG"] = "1"; os.environ["MGB_ATTN_DBG_DUMP"] = "19"      # 6 encoder launches + 12 during the capture; the 19th is the direct step          # 48 launch_linear_ts calls during the graph capture, the 49th is the direct step below
from magpie_tts_cpp_b200 import binding, fixtures
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
frames = 40
m = binding.Model(fixtures.ensure_fixture("model-f32"), 0, binding.PREC_BF16)
rng = np.random.default_rng(7)
texts = [[2378] + rng.integers(0, 90, int(rng.integers(18, 79))).tolist() + [2379] for _ in range(B)]
s = m.session(batch=B, max_text=96, max_seq=110 + frames + 16)
codes = np.repeat(np.random.default_rng(42).integers(0, 2016, (1, frames, 8)).astype(np.int32), B, axis=0)
s.encode_text(texts, want_output=False); s.prefill([b % 5 for b in range(B)])
s.teacher_forced(codes, want_hidden=False, want_logits=False)
print("B=%d: %.1f us/step" % (B, s.last_loop_ms * 1e3 / frames))
s.decoder_step(codes[:, 0], want_hidden=False)
