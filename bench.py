#!/usr/bin/env python
"""bench.py -- decoder frames/s of the per-frame synthesis hot path (BASELINE.json metric).

Workload at N=1 (BASELINE.json configs[1]): teacher-forced decoder step + local transformer,
batch 1, one 500-frame utterance, bf16, random-init Magpie-357M GGUF, text "Hello, world!".
One *step* = one pass over the 500-frame utterance (500 decoder steps + 500 LT passes); the
KV cache is re-primed (encode + 110-frame prefill, untimed, like the reference's own fps
printout at magpie.cpp:4426-4429 which covers the generation loop only) before every step.

  value  device time of the loop (CUDA events on the session's stream), inputs resident in HBM
  e2e    the same loop through the C-ABI call with HOST buffers: H2D of the forced codes and D2H of
         the greedy codes inside the timed region (wall clock around mgb_teacher_forced)

N > 1: launched under torchrun, one rank per GPU, each rank runs the same per-GPU workload on its
own utterance (independent utterances, no collective on the data path -> weak scaling); value is
all ranks' frames / max-over-ranks time.

extra (N=1): configs[3] (64 utterances x 215 frames) and configs[4]'s per-GPU share (Q8_0 GGUF, 32 x 2600 frames), each with
its HBM step roofline (weights once + per-utterance mean K/V scan), and the codec on configs[2] with its tensor roofline.

--impl reference: the CPU restatement of the reference (oracle/, all host threads) on a bounded
sample of the same workload.  The real reference cannot be built here (needs ggml).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HELLO = [2378, 7, 4, 11, 11, 14, 32, 26, 22, 14, 17, 11, 3, 32, 28, 2379]   # "Hello, world!" (synthetic vocab)
FRAMES = int(os.environ.get("MGB_BENCH_FRAMES", "500"))   # override only for profiling runs (not a bench value)
METRIC = "decoder_frames_per_s"
UNIT = "frames/s"


def step_roofline(m, B, ctx, frames, text_len, s_per_step):
    """HBM roofline of one batched decoder+LT step: weights once + per utterance the mean K/V scan (W + B * KV(P))."""
    hp, wsz = m.hp, 2
    pbar = ctx + (frames + 1) / 2.0
    kv = hp["dec_layers"] * 2 * pbar * hp["d_model"] * wsz + hp["dec_layers"] * 2 * text_len * 128 * wsz
    by = m.step_weight_bytes + B * kv
    hbm, _, src = peaks()
    return {"bound": "hbm", "achieved": by / s_per_step / 1e9, "peak": hbm, "peak_source": src, "unit": "GB/s",
            "frac": by / s_per_step / 1e9 / hbm, "algorithmic_bytes_per_step": by, "step_us": s_per_step * 1e6}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1400.0))), "measured"
    return 6650.0, 1400.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop, self.th = index, [], False, None

    def _run(self):
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def __enter__(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.th.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def forced_codes(frames, seed=42):
    return np.random.default_rng(seed).integers(0, 2016, (1, frames, 8)).astype(np.int32)


def cpu_reference_run(frames, steps, warmup, threads=None):
    """Oracle (restated reference, CPU, OpenMP) teacher-forced decoder+LT loop; returns (frames/s, cores, seconds)."""
    from magpie_tts_cpp_b200 import fixtures
    from oracle import oracle
    oracle.build()
    if threads:
        oracle.set_num_threads(threads)
    cores = oracle.num_threads()
    o = oracle.OracleModel(fixtures.ensure_fixture("model-f32"))
    enc = o.encode_text(HELLO)
    codes = forced_codes(frames)[0]
    bos = np.full(8, o.hp["audio_bos_id"], np.int32)
    times = []
    for it in range(warmup + steps):
        st = o.new_state(enc, 0, max_seq=o.hp["context_frames"] + frames + 16)
        t0 = time.perf_counter()
        prev = bos
        for t in range(frames):
            h = st.step(prev)
            o.lt_sample(h, 0.0, 80, forced_codes=codes[t], want_logits=False)
            prev = codes[t]
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    total = sum(times)
    return frames * len(times) / total, cores, total


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    frames = 60          # bounded sample: 60 of the 500 frames per step (KV length 111..170)
    fps, cores, total = cpu_reference_run(frames, args.steps, min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "teacher-forced decoder+LT step, batch 1, random-init Magpie-357M f32 GGUF, 'Hello, world!'",
                   "frames_per_step": frames},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"first {frames} frames of the 500-frame utterance, {args.steps} passes, f32 oracle (OpenMP)"},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from magpie_tts_cpp_b200 import binding, fixtures
    rank, world, local = dist_env()
    if args.gpus > 1 and world == 1:
        # convenience: `python bench.py --gpus N` re-launches itself under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not os.path.exists(binding.LIB_PATH):
        binding.build()
    prec = binding.PREC_BF16 if args.dtype == "bf16" else binding.PREC_F32
    if rank == 0:
        fixtures.ensure_fixture("model-f32")
    if world > 1:
        dist.barrier()
    m = binding.Model(fixtures.ensure_fixture("model-f32"), local, prec)
    B = args.batch
    s = m.session(batch=B, max_text=32, max_seq=m.hp["context_frames"] + FRAMES + 16)
    codes = np.repeat(forced_codes(FRAMES), B, axis=0)
    toks = [HELLO] * B

    def prime():
        s.encode_text(toks, want_output=False)
        s.prefill([0] * B)

    def one_step():
        prime()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _, _, gr = s.teacher_forced(codes, want_hidden=False, want_logits=False, want_greedy=True)
        wall = time.perf_counter() - t0
        return s.last_loop_ms * 1e-3, wall, s.last_loop_launches, gr

    for _ in range(max(args.warmup, 3)):
        one_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dev_t = wall_t = 0.0
    launches = 0
    with ClockSampler(local) as clk:
        for _ in range(args.steps):
            d, w, l, gr = one_step()
            dev_t += d; wall_t += w; launches += l
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        t = torch.tensor([dev_t, wall_t], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_t, wall_t = float(t[0]), float(t[1])
    frames_total = FRAMES * B * args.steps * world
    value = frames_total / dev_t
    e2e = frames_total / wall_t

    # roofline of the frame loop (HBM bound): unique weight bytes + mean KV bytes per frame per utterance
    hbm, _, which = peaks()
    wsz = 2 if prec == binding.PREC_BF16 else 4
    hp = m.hp
    pbar = hp["context_frames"] + 1 + (FRAMES - 1) / 2.0
    kv_bytes = hp["dec_layers"] * 2 * pbar * hp["d_model"] * wsz + hp["dec_layers"] * 2 * len(HELLO) * 128 * wsz
    bytes_per_iter = m.step_weight_bytes + B * kv_bytes
    iter_s = dev_t / (FRAMES * args.steps)
    achieved = bytes_per_iter / iter_s / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": 1e3 * dev_t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": f"teacher-forced decoder+LT step, batch {B}/GPU, 500-frame utterance, random-init Magpie-357M, 'Hello, world!'",
                   "frames_per_step": FRAMES * B, "l2": "weights (185 MB bf16 per frame) exceed the 126 MB L2; no explicit flush"},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(codes.nbytes), "d2h_bytes_per_step": int(gr.nbytes)},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "one frame = decoder-step kernels + local-transformer kernel",
                     "achieved": achieved, "peak": hbm, "peak_source": which, "unit": "GB/s", "frac": achieved / hbm,
                     # DRAM bytes per frame of the frame-loop kernel from the committed ncu --set full capture (7.054 GB over a
                     # 40-frame launch, profiles/r1_frame_loop_kernel_ncu_full.txt); batch 1 / bf16 only
                     "traffic": 176.4e6 if (B == 1 and prec == binding.PREC_BF16) else None,
                     "traffic_source": "profiles/r1_frame_loop_kernel_ncu_full.txt (dram read+write of one 40-frame launch / 40)",
                     "algorithmic_bytes_per_launch": bytes_per_iter, "launch_us": iter_s * 1e6},
        "clocks": clk.summary(),
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        fps, cores, total = cpu_reference_run(40, 1, 0)
        line["cpu_baseline"] = {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": "first 40 frames of the same 500-frame utterance, f32 oracle (OpenMP), 1 pass"}
    if rank == 0 and world == 1 and not args.no_extra:
        line["extra"] = extra_measurements(binding, fixtures, m, args)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def extra_measurements(binding, fixtures, m, args):
    """Secondary numbers of the BASELINE metric string: batched decoder frames/s and codec audio-s/s."""
    out = {}
    try:
        # BASELINE configs[3]: 64 utterances per GPU, 10 s (215 frames) each, distinct random texts of 20..80 tokens,
        # speakers 0..4, EOS disabled (teacher-forced fixed length)
        B = 64
        frames = 215
        rng = np.random.default_rng(7)
        texts = [[2378] + rng.integers(0, 90, int(rng.integers(18, 79))).tolist() + [2379] for _ in range(B)]
        s = m.session(batch=B, max_text=96, max_seq=m.hp["context_frames"] + frames + 16)
        codes = np.repeat(forced_codes(frames), B, axis=0)
        for _ in range(2):
            s.encode_text(texts, want_output=False)
            s.prefill([b % 5 for b in range(B)])
            s.teacher_forced(codes, want_hidden=False, want_logits=False)
        out["decoder_b64_frames_per_s"] = B * frames / (s.last_loop_ms * 1e-3)
        out["decoder_b64_sample"] = "config 4: 64 utterances x 215 frames, random texts of 20..80 tokens, speakers 0..4"
        out["decoder_b64_launches_per_step"] = s.last_loop_launches / frames
        out["decoder_b64_roofline"] = step_roofline(m, B, m.hp["context_frames"], frames, 50, s.last_loop_ms * 1e-3 / frames)
        s.close()
    except Exception as e:  # noqa: BLE001
        out["decoder_b64_error"] = str(e)
    try:
        # BASELINE configs[4] (per-GPU share): Q8_0 GGUF, 32 utterances, 2600-frame utterances (KV length 110 -> 2710)
        mq = binding.Model(fixtures.ensure_fixture("model-long-q8"), m.device if hasattr(m, "device") else 0, binding.PREC_BF16)
        B, frames = 32, 2600
        s = mq.session(batch=B, max_text=32, max_seq=mq.hp["context_frames"] + frames + 16)
        codes = np.repeat(forced_codes(frames), B, axis=0)
        s.encode_text([HELLO] * B, want_output=False); s.prefill([b % 5 for b in range(B)])
        s.teacher_forced(codes[:, :64], want_hidden=False, want_logits=False)
        s.encode_text([HELLO] * B, want_output=False); s.prefill([b % 5 for b in range(B)])
        s.teacher_forced(codes, want_hidden=False, want_logits=False)
        out["decoder_q8_long_b32_frames_per_s"] = B * frames / (s.last_loop_ms * 1e-3)
        out["decoder_q8_long_sample"] = "config 5 per-GPU share: Q8_0 GGUF (dequantised to bf16 at load), 32 utterances x 2600 frames"
        out["decoder_q8_long_roofline"] = step_roofline(mq, B, mq.hp["context_frames"], frames, len(HELLO), s.last_loop_ms * 1e-3 / frames)
        s.close(); mq.close()
    except Exception as e:  # noqa: BLE001
        out["decoder_q8_long_error"] = str(e)
    try:
        # BASELINE configs[2]: nano-codec decode only, 60 s of 21.5 fps codes (1291 frames), batch 32, one call
        c = binding.Codec(fixtures.ensure_fixture("codec-f32"), 0)
        Bc, T = 32, 1291
        codes = np.random.default_rng(42).integers(0, 2016, (Bc, 8, T)).astype(np.int32)
        c.decode(codes[:2, :, :64])                  # warm-up: weight repack, allocations
        c.decode(codes)
        t0 = time.perf_counter()
        c.decode(codes)
        wall = time.perf_counter() - t0
        _, tf, _ = peaks()
        audio_s = Bc * T * 1024 / 22050.0
        flops = Bc * T * 2.447e9                       # 1.2234 GMAC per frame
        out["codec_audio_s_per_s"] = audio_s / (c.last_ms * 1e-3)
        out["codec_e2e_audio_s_per_s"] = audio_s / wall            # host codes in, host PCM out (169 MB D2H)
        out["codec_sample"] = f"config 3: batch {Bc} x {T} frames (60 s each), one mgb_codec_decode call"
        out["codec_launches"] = int(c.last_launches)
        out["codec_roofline"] = {"bound": "tensor", "achieved": flops / (c.last_ms * 1e-3) / 1e12, "peak": tf, "unit": "TFLOP/s",
                                 "frac": flops / (c.last_ms * 1e-3) / 1e12 / tf, "algorithmic_flops": flops}
        if not args.no_cpu_baseline:
            from oracle import oracle
            oc = oracle.OracleCodec(fixtures.ensure_fixture("codec-f32"))
            Tc = 12
            t0 = time.perf_counter()
            oc.decode(codes[0, :, :Tc])
            dt = time.perf_counter() - t0
            out["codec_cpu_baseline"] = {"value": Tc * 1024 / 22050.0 / dt, "unit": "audio-s/s", "cores": oracle.num_threads(), "kind": "port",
                                         "sample": f"first {Tc} frames of utterance 0, f32 oracle (OpenMP)"}
    except Exception as e:  # noqa: BLE001
        out["codec_error"] = str(e)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--batch", type=int, default=1, help="utterances per GPU (1 = BASELINE configs[1])")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
