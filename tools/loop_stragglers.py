"""Profiling aid: per-CTA phase timestamps (globaltimer) of the last frame of the persistent frame-loop kernel:
who is late, per phase.   python tools/loop_stragglers.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["MGB_LOOP_DBG"] = "1"
from magpie_tts_cpp_b200 import binding, fixtures
HELLO = [2378, 7, 4, 11, 11, 14, 32, 26, 22, 14, 17, 11, 3, 32, 28, 2379]
T = 100
m = binding.Model(fixtures.ensure_fixture("model-f32"), 0, binding.PREC_BF16)
s = m.session(batch=1, max_text=32)
codes = np.random.default_rng(1).integers(0, 2016, (1, T, 8)).astype(np.int32)
for _ in range(2):
    s.encode_text([HELLO], want_output=False); s.prefill([0])
    s.teacher_forced(codes, want_hidden=False, want_logits=False)
print("%.1f us/frame" % (s.last_loop_ms * 1e3 / T))
G, PER = 148, 256
L = m.hp["dec_layers"]
st = s.debug_stamps(G * PER).astype(np.int64).reshape(G, PER)
t0 = st[:, 0].min()
nk = 110 + T                                   # keys at the last frame
n_att = 12 * min(6, max(1, (nk + 127) // 128))  # CTAs running attention items (H * S_split)
names = ["P1 in", "P1 out", "P2 in*", "P2 out", "P3 in", "P3 out", "P4 in", "P4 out", "P5 in", "P5 out", "P6 in", "P6 out"]
rows = []
for b in range(G):
    per_layer = 12 if b < n_att else 11     # non-attention CTAs lack the "P2 in" stamp
    base = 1
    arr = np.full((L, 12), np.nan)
    for l in range(L):
        seg = st[b, base:base + per_layer].astype(float)
        if per_layer == 12: arr[l] = seg
        else: arr[l, :2] = seg[:2]; arr[l, 3:] = seg[2:]
        base += per_layer
    rows.append(arr)
A = np.stack(rows) - t0          # [G][L][12]
print("n_att", n_att, "decoder layers: per stamp, spread over CTAs (ns), mean over layers 1..L-1")
for i, nm in enumerate(names):
    x = A[:, 1:, i]
    med = np.nanmedian(x, 0); mn = np.nanmin(x, 0); mx = np.nanmax(x, 0)
    worst = np.nanargmax(np.where(np.isnan(x), -1, x), 0)
    print(f"  {nm:7s} med-min {np.mean(med-mn):6.0f}  max-med {np.mean(mx-med):6.0f}   latest CTAs {np.bincount(worst, minlength=G).argsort()[-4:][::-1].tolist()}")
med_t = np.nanmedian(A, 0)        # [L][12]
print("median timeline of layer 5 (ns from P1 in):", (med_t[5] - med_t[5, 0]).astype(int).tolist(), " next layer P1 in:", int(med_t[6, 0] - med_t[5, 0]))
lat = A[:, 1:, 1::2]   # 'out' stamps
late_by_cta = np.nanmean(lat - np.nanmedian(lat, 0, keepdims=True), axis=(1, 2))
print("mean lateness of 'out' stamps vs median, by CTA (ns), every 4th:", [int(v) for v in late_by_cta[::4]])
order = np.argsort(late_by_cta)[::-1]
print("top-10 late CTAs:", order[:10].tolist(), [int(late_by_cta[i]) for i in order[:10]])
