#!/usr/bin/env python
"""bench.py -- decoder frames/s of the per-frame synthesis hot path (BASELINE.json metric:
"decoder frames/s (b=1; b=64/GPU @1/2/4/8 B200) + codec audio-s/s; vs ggml CPU").

N = 1 (default; BASELINE.json configs[1]): teacher-forced decoder step + local transformer, batch 1, one 500-frame
utterance, bf16, random-init Magpie-357M GGUF, text "Hello, world!".  One *step* = one pass over the 500-frame utterance
(500 decoder steps + 500 LT passes); the KV cache is re-primed (encode + 110-frame prefill, untimed, like the reference's own
fps printout at magpie.cpp:4426-4429 which covers the generation loop only) before every step.

N > 1 (launched under torchrun, one rank per GPU; BASELINE.json configs[3]): 64 utterances per GPU, 215 frames (10 s) each,
distinct random texts of 20..80 tokens, speakers 0..4, teacher-forced (fixed length).  Utterances are independent, so there is
no collective on the data path (weak scaling; NCCL is used only for the barrier and the max-over-ranks of the timings);
value = all ranks' frames / max-over-ranks device time.  `--batch B` forces B utterances per GPU at any N (B = 1: configs[1]
workload, B > 1: configs[3] shape).  rank 0 also reports one utterance per GPU (`extra.decoder_b1_per_gpu_frames_per_s`).

  value  device time of the loop (CUDA events on the session's stream), inputs resident in HBM
  e2e    the same loop through the C-ABI call with HOST buffers: H2D of the forced codes and D2H of the greedy codes inside
         the timed region (wall clock around mgb_teacher_forced)
  parity (N=1) greedy codes of the timed run against the oracle's on all 500 frames

--inproc: the same per-GPU workload in ONE process through the C-ABI pool (mgb_pool_*: one replica + one submission thread +
one stream per GPU, no torch.distributed, no NCCL) -- the form BASELINE.json's north_star names.

extra (N=1): configs[3] (64 utterances x 215 frames) and configs[4]'s per-GPU share (Q8_0 GGUF, 32 x 2600 frames), each with
its HBM step roofline (weights once + per-utterance mean K/V scan), the codec on configs[2] with its tensor
roofline, batch-1 frames/s at text lengths 16 / 70 / 256 and the wall time of the `magpie-tts` CLI.

--impl reference: the CPU restatement of the reference (oracle/, all host threads) on a bounded sample of the same workload.
The real reference cannot be built here (needs ggml).
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
B64_FRAMES = int(os.environ.get("MGB_BENCH_B64_FRAMES", "215"))
METRIC = "decoder_frames_per_s"
UNIT = "frames/s"


def config4_texts(nb, seed=7):
    """BASELINE configs[3]: distinct random texts of 20..80 tokens (BOS + 18..78 symbols + EOS), speakers 0..4."""
    rng = np.random.default_rng(seed)
    return [[2378] + rng.integers(0, 90, int(rng.integers(18, 79))).tolist() + [2379] for _ in range(nb)], [b % 5 for b in range(nb)]


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


def host_threads():
    """All host cores this process may use.  Under torch.distributed.run the environment carries OMP_NUM_THREADS=1, which the
    CPU arm must not inherit (round 1: the reference arm ran on one core at N >= 2)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_reference_run(tokens, speaker, frames, steps, warmup, want_codes=False):
    """Oracle (restated reference, CPU, OpenMP on all host cores) teacher-forced decoder+LT loop over the first `frames` frames
    of the config-2 code stream; returns dict(fps, cores, seconds[, greedy, logits])."""
    from magpie_tts_cpp_b200 import fixtures
    from oracle import oracle
    oracle.build()
    oracle.set_num_threads(host_threads())
    cores = oracle.num_threads()
    o = oracle.OracleModel(fixtures.ensure_fixture("model-f32"))
    enc = o.encode_text(tokens)
    codes = forced_codes(frames)[0]
    bos = np.full(8, o.hp["audio_bos_id"], np.int32)
    times, greedy, logits = [], None, None
    for it in range(warmup + steps):
        st = o.new_state(enc, speaker, max_seq=o.hp["context_frames"] + frames + 16)
        keep = want_codes and it == warmup + steps - 1
        gr, lgs = [], []
        t0 = time.perf_counter()
        prev = bos
        for t in range(frames):
            h = st.step(prev)
            _, a, lg = o.lt_sample(h, 0.0, 80, forced_codes=codes[t], want_logits=keep)
            if keep:
                gr.append(a); lgs.append(lg)
            prev = codes[t]
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
        if keep:
            greedy, logits = np.stack(gr), np.stack(lgs)
    total = sum(times)
    return {"fps": frames * len(times) / total, "cores": cores, "seconds": total, "greedy": greedy, "logits": logits}


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    B = args.batch if args.batch > 0 else (1 if max(world, args.gpus) == 1 else 64)
    warm = min(args.warmup, 1)
    if B == 1:
        frames, tokens, speaker = min(FRAMES, 250), HELLO, 0
        workload = "teacher-forced decoder+LT step, batch 1, random-init Magpie-357M f32 GGUF, 'Hello, world!'"
        sample = (f"first {frames} frames of the 500-frame utterance (KV length 111..{110 + frames}), {args.steps} timed passes after "
                  f"{warm} warm-up pass, f32 oracle (OpenMP)")
    else:
        texts, speakers = config4_texts(B)
        frames, tokens, speaker = B64_FRAMES, texts[0], speakers[0]
        workload = (f"teacher-forced decoder+LT step, {B} utterances/GPU x {B64_FRAMES} frames (config 4), random-init Magpie-357M f32 GGUF: "
                    "the reference synthesises utterances one at a time, so its frames/s is a per-utterance rate and does not grow with the utterance count")
        sample = (f"utterance 0 of the {B} ({len(tokens)} text tokens), all {frames} frames, {args.steps} timed passes after {warm} "
                  f"warm-up pass, f32 oracle (OpenMP)")
    r = cpu_reference_run(tokens, speaker, frames, args.steps, warm)
    fps = r["fps"]
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": warm, "ms_per_step": 1e3 * r["seconds"] / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "frames_per_step": frames, "utterances": 1},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def greedy_parity(gr, ref):
    """Greedy codes of the timed bf16 run against the f32 oracle's (all frames of the utterance)."""
    lg, rg = ref["logits"], ref["greedy"]
    n = min(len(gr), len(rg))
    agree = gr[:n] == rg[:n]
    fin = np.isfinite(lg[:n])
    rms = np.sqrt(np.sum(np.where(fin, lg[:n], 0.0).astype(np.float64) ** 2, axis=-1) / np.maximum(fin.sum(-1), 1))
    srt = np.sort(np.where(fin, lg[:n], -np.inf), axis=-1)
    top1, top2 = srt[..., -1].astype(np.float64), srt[..., -2].astype(np.float64)
    q = (top1 - top2) > 2.0 * (2e-2 * np.abs(top1) + 2e-2 * rms)       # picks a kernel inside the bf16 band cannot flip
    return {"frames": int(n), "parity_frames_equal": float(np.all(agree, axis=1).mean()), "codes_equal": float(agree.mean()),
            "margin_qualified_picks": int(q.sum()), "margin_qualified_equal": float(agree[q].mean()) if q.any() else None,
            "note": "bf16 kernel vs f32 oracle; random-init logits are nearly flat, so picks whose top-2 margin is inside the bf16 tolerance "
                    "band may flip: the margin-qualified picks are the ones the 2e-2 logit bar protects"}


def workload_for(B, m):
    """(texts, speakers, frames, max_text, codes, label)"""
    if B == 1:
        return [HELLO], [0], FRAMES, 32, forced_codes(FRAMES), \
            "teacher-forced decoder+LT step, batch 1/GPU, 500-frame utterance, random-init Magpie-357M, 'Hello, world!' (config 2)"
    texts, speakers = config4_texts(B)
    codes = np.repeat(forced_codes(B64_FRAMES), B, axis=0)
    return texts, speakers, B64_FRAMES, 96, codes, \
        (f"teacher-forced decoder+LT step, {B} utterances/GPU x {B64_FRAMES} frames (10 s), distinct random texts of 20..80 tokens, "
         "speakers 0..4, random-init Magpie-357M (config 4)")


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
    B = args.batch if args.batch > 0 else (1 if world == 1 else 64)

    def measure(B, steps, warmup, clocks=False):
        toks, spk, frames, max_text, codes, label = workload_for(B, m)
        s = m.session(batch=B, max_text=max_text, max_seq=m.hp["context_frames"] + frames + 16)

        def one_step():
            s.encode_text(toks, want_output=False)
            s.prefill(spk)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            _, _, gr = s.teacher_forced(codes, want_hidden=False, want_logits=False, want_greedy=True)
            wall = time.perf_counter() - t0
            return s.last_loop_ms * 1e-3, wall, s.last_loop_launches, gr

        for _ in range(warmup):
            one_step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        dev_t = wall_t = 0.0
        launches = 0
        gr = None
        if clocks:
            with ClockSampler(local) as clk:
                for _ in range(steps):
                    d, w, l, gr = one_step()
                    dev_t += d; wall_t += w; launches += l
            ck = clk.summary()
        else:
            ck = None
            for _ in range(steps):
                d, w, l, gr = one_step()
                dev_t += d; wall_t += w; launches += l
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            t = torch.tensor([dev_t, wall_t], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dev_t, wall_t = float(t[0]), float(t[1])
        s.close()
        mean_text = float(np.mean([len(t) for t in toks]))
        return dict(dev_t=dev_t, wall_t=wall_t, launches=launches, gr=gr, frames=frames, codes=codes, label=label, clocks=ck,
                    mean_text=mean_text)

    W = max(args.warmup, 3)
    r = measure(B, args.steps, W, clocks=True)
    frames_total = r["frames"] * B * args.steps * world
    value = frames_total / r["dev_t"]
    e2e = frames_total / r["wall_t"]

    # roofline of the frame loop (HBM bound): unique weight bytes + mean KV bytes per frame per utterance
    hbm, _, which = peaks()
    wsz = 2 if prec == binding.PREC_BF16 else 4
    hp = m.hp
    pbar = hp["context_frames"] + 1 + (r["frames"] - 1) / 2.0
    kv_bytes = hp["dec_layers"] * 2 * pbar * hp["d_model"] * wsz + hp["dec_layers"] * 2 * r["mean_text"] * 128 * wsz
    bytes_per_iter = m.step_weight_bytes + B * kv_bytes
    iter_s = r["dev_t"] / (r["frames"] * args.steps)
    achieved = bytes_per_iter / iter_s / 1e9
    b1 = B == 1 and prec == binding.PREC_BF16
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
        "ms_per_step": 1e3 * r["dev_t"] / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": r["label"], "utterances_per_gpu": B, "frames_per_step": r["frames"] * B,
                   "launch": "one process per GPU (torchrun); independent utterances, no data-path collective (NCCL only for the barrier "
                             "and the max-over-ranks of the timings)" if world > 1 else "single process",
                   "l2": f"weights (185 MB bf16) + K/V ({B} x {kv_bytes / 1e6:.1f} MB) per frame exceed the 126 MB L2; no explicit flush"},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(r["codes"].nbytes), "d2h_bytes_per_step": int(r["gr"].nbytes)},
        "gpu_launches": int(r["launches"]),
        "roofline": {"bound": "hbm",
                     "kernel": "frame_loop_kernel: one frame = 12-layer decoder step + local transformer, one persistent launch per 500 frames" if b1
                     else "one step = decoder-step kernels + local-transformer kernel for all utterances of the GPU",
                     "achieved": achieved, "peak": hbm, "peak_source": which, "unit": "GB/s", "frac": achieved / hbm,
                     # DRAM bytes per frame of the frame-loop kernel from the committed ncu --set full capture (7.054 GB over a
                     # 40-frame launch, profiles/r1_frame_loop_kernel_ncu_full.txt); batch 1 / bf16 only
                     "traffic": 176.4e6 if b1 else None,
                     "traffic_source": "profiles/r1_frame_loop_kernel_ncu_full.txt (dram read+write of one 40-frame launch / 40)" if b1 else None,
                     "algorithmic_bytes_per_launch": bytes_per_iter, "launch_us": iter_s * 1e6},
        "clocks": r["clocks"],
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # the oracle over the same utterance on the host cores (bounded: one pass; 500 frames take a few seconds), which also
        # yields the reference greedy codes the timed run's codes are compared with
        cb_frames = r["frames"] if B == 1 else min(r["frames"], 120)
        toks, spk = (HELLO, 0) if B == 1 else (config4_texts(B)[0][0], 0)
        ref = cpu_reference_run(toks, spk, cb_frames, 1, 0, want_codes=True)
        line["cpu_baseline"] = {"value": ref["fps"], "unit": UNIT, "cores": ref["cores"], "kind": "port",
                                "sample": f"{cb_frames} frames of utterance 0 of the same workload, f32 oracle (OpenMP, all host cores), 1 pass"}
        if prec == binding.PREC_BF16:
            line["parity"] = greedy_parity(r["gr"][0], ref)
            line["parity_frames_equal"] = line["parity"]["parity_frames_equal"]
    extra = {}
    if B != 1 and not args.no_extra:
        # one utterance per GPU (configs[1] workload) on every rank at once
        r1 = measure(1, max(1, min(args.steps, 3)), 3)
        extra["decoder_b1_per_gpu_frames_per_s"] = r1["frames"] * world * max(1, min(args.steps, 3)) / r1["dev_t"]
        extra["decoder_b1_sample"] = "config 2 workload (batch 1, 500 frames) on every GPU at once, aggregate frames/s"
    if rank == 0 and world == 1 and B == 1 and not args.no_extra:
        extra.update(extra_measurements(binding, fixtures, m, args))
    if extra and rank == 0:
        line["extra"] = extra
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_inproc(args):
    """One process, one model replica + one submission thread + one stream per GPU (C-ABI pool), no torch.distributed."""
    import torch
    from magpie_tts_cpp_b200 import binding, fixtures
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    n = args.gpus
    assert n <= torch.cuda.device_count(), "not enough GPUs"
    prec = binding.PREC_BF16 if args.dtype == "bf16" else binding.PREC_F32
    pool = binding.Pool(fixtures.ensure_fixture("model-f32"), devices=list(range(n)), precision=prec)
    B = args.batch if args.batch > 0 else 64
    if B == 1:
        texts, speakers, frames = [HELLO] * n, [0] * n, FRAMES
        codes = np.repeat(forced_codes(frames), n, axis=0)
    else:
        t1, s1 = config4_texts(B)
        # utterance i runs on device i mod n: interleave so that every device gets the same B texts
        texts = [t1[i // n] for i in range(B * n)]
        speakers = [s1[i // n] for i in range(B * n)]
        frames = B64_FRAMES
        codes = np.repeat(forced_codes(frames), B * n, axis=0)
    W = max(args.warmup, 3)
    for _ in range(W):
        pool.teacher_forced(texts, codes, speakers)
    dev_t = wall_t = 0.0
    with ClockSampler(0) as clk:
        for _ in range(args.steps):
            t0 = time.perf_counter()
            gr = pool.teacher_forced(texts, codes, speakers)
            wall_t += time.perf_counter() - t0
            dev_t += float(pool.last_device_ms.max()) * 1e-3
    total = frames * B * n * args.steps
    line = {
        "metric": METRIC, "value": total / dev_t, "unit": UNIT, "n_gpus": n, "steps": args.steps, "warmup": W,
        "ms_per_step": 1e3 * dev_t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic", "impl": "ours-inproc",
        "config": {"workload": f"teacher-forced decoder+LT step, {B} utterances/GPU x {frames} frames, in-process pool", "utterances_per_gpu": B,
                   "frames_per_step": frames * B * n,
                   "launch": "ONE process: mgb_pool (C++), one model replica + one submission thread + one stream per GPU, "
                             "no torch.distributed, no NCCL; value = frames / max-over-devices loop time (CUDA events per device)"},
        "e2e": {"value": total / wall_t, "unit": UNIT, "h2d_bytes_per_step": int(codes.nbytes), "d2h_bytes_per_step": int(gr.nbytes),
                "note": "wall clock around mgb_pool_teacher_forced: includes the text encoder and the 110-frame prefill of every device (sessions are cached by the pool)"},
        "clocks": clk.summary(),
    }
    print(json.dumps(line), flush=True)
    pool.close()


def extra_measurements(binding, fixtures, m, args):
    """Secondary numbers of the BASELINE metric string: batched decoder frames/s and codec audio-s/s."""
    out = {}
    try:
        # BASELINE configs[3]: 64 utterances per GPU, 10 s (215 frames) each, distinct random texts of 20..80 tokens,
        # speakers 0..4, EOS disabled (teacher-forced fixed length)
        B = 64
        frames = B64_FRAMES
        texts, speakers = config4_texts(B)
        s = m.session(batch=B, max_text=96, max_seq=m.hp["context_frames"] + frames + 16)
        codes = np.repeat(forced_codes(frames), B, axis=0)
        for _ in range(2):
            s.encode_text(texts, want_output=False)
            s.prefill(speakers)
            s.teacher_forced(codes, want_hidden=False, want_logits=False)
        out["decoder_b64_frames_per_s"] = B * frames / (s.last_loop_ms * 1e-3)
        out["decoder_b64_sample"] = "config 4: 64 utterances x 215 frames, random texts of 20..80 tokens, speakers 0..4"
        out["decoder_b64_launches_per_step"] = s.last_loop_launches / frames
        out["decoder_b64_roofline"] = step_roofline(m, B, m.hp["context_frames"], frames, 50, s.last_loop_ms * 1e-3 / frames)
        s.close()
    except Exception as e:  # noqa: BLE001
        out["decoder_b64_error"] = str(e)
    try:
        # batch-1 frame loop at longer texts (the reference's doc example is 70 tokens): same 500-frame run
        for n_text in (70, 256):
            rng = np.random.default_rng(n_text)
            text = [2378] + rng.integers(0, 90, n_text - 2).tolist() + [2379]
            s = m.session(batch=1, max_text=n_text, max_seq=m.hp["context_frames"] + FRAMES + 16)
            for _ in range(2):
                s.encode_text([text], want_output=False); s.prefill([0])
                s.teacher_forced(forced_codes(FRAMES), want_hidden=False, want_logits=False)
            out[f"decoder_b1_text{n_text}_frames_per_s"] = FRAMES / (s.last_loop_ms * 1e-3)
            out[f"decoder_b1_text{n_text}_launches"] = int(s.last_loop_launches)
            s.close()
    except Exception as e:  # noqa: BLE001
        out["decoder_b1_long_text_error"] = str(e)
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
        out["decoder_q8_long_sample"] = "config 5 per-GPU share: Q8_0 GGUF, 32 utterances x 2600 frames"
        out["decoder_q8_long_roofline"] = step_roofline(mq, B, mq.hp["context_frames"], frames, len(HELLO), s.last_loop_ms * 1e-3 / frames)
        s.close()
        # the same share as STREAMING synthesis (callback every 4 frames, each chunk decoded by the codec with 25 frames of
        # context = seamless audio), bounded to the first 400 of the 2600 frames
        c5 = binding.Codec(fixtures.ensure_fixture("codec-f32"), 0)
        sf, got = 400, [0]
        s = mq.session(batch=B, max_text=32, max_seq=mq.hp["context_frames"] + sf + 16)

        def on_audio(u, pcm, frames_done, is_last):
            got[0] += len(pcm)
            return False
        walls = []
        for steps_ in (16, sf, sf):      # warm-up, then two timed runs (wall clock with 3 200 host callbacks: the best of two, both reported)
            s.encode_text([HELLO] * B, want_output=False); s.prefill([b % 5 for b in range(B)])
            got[0] = 0
            t0 = time.perf_counter()
            s.stream_generate(c5, on_audio, max_steps=steps_, frames_per_chunk=4, codec_context_frames=25, ignore_eos=True)
            if steps_ == sf:
                walls.append(time.perf_counter() - t0)
        wall = min(walls)
        out["stream_q8_b32_frames_per_s"] = B * sf / wall
        out["stream_q8_b32_audio_s_per_s"] = got[0] / 22050.0 / wall
        out["stream_q8_b32_wall_samples_s"] = [round(w, 3) for w in walls]
        out["stream_q8_b32_sample"] = ("config 5 per-GPU share as streaming: 32 utterances in lock step, first 400 frames, chunks of 4 frames, every chunk "
                                       "decoded with 25 context frames (29 frames per utterance and chunk through the codec) and copied to the host callback; wall clock")
        s.close(); c5.close(); mq.close()
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
            oracle.set_num_threads(host_threads())
            oc = oracle.OracleCodec(fixtures.ensure_fixture("codec-f32"))
            Tc = 12
            t0 = time.perf_counter()
            oc.decode(codes[0, :, :Tc])
            dt = time.perf_counter() - t0
            out["codec_cpu_baseline"] = {"value": Tc * 1024 / 22050.0 / dt, "unit": "audio-s/s", "cores": oracle.num_threads(), "kind": "port",
                                         "sample": f"first {Tc} frames of utterance 0, f32 oracle (OpenMP)"}
    except Exception as e:  # noqa: BLE001
        out["codec_error"] = str(e)
    try:
        # the drop-in surface end to end: `magpie-tts -t "Hello, world!" -o out.wav` (process start, GGUF load, encode, prefill,
        # generation loop with EOS, 32-frame-chunk codec decode, WAV write), wall clock
        cli = os.path.join(ROOT, "magpie_tts_cpp_b200", "bin", "magpie-tts")
        if os.path.exists(cli):
            wav = "/tmp/magpie_b200_bench_cli.wav"
            cmd = [cli, "-m", fixtures.ensure_fixture("model-f32"), "-c", fixtures.ensure_fixture("codec-f32"), "-t", "Hello, world!",
                   "--temp", "0", "-o", wav, "-q"]
            subprocess.run(cmd, capture_output=True, text=True, timeout=300)          # page the files in
            walls = []
            for _ in range(3):     # process start, CUDA context, GGUF load and teardown vary by seconds between runs on a shared host: all samples reported
                t0 = time.perf_counter()
                pr = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
                walls.append(time.perf_counter() - t0)
            wall = min(walls)
            if pr.returncode == 0 and os.path.exists(wav):
                n_frames = (os.path.getsize(wav) - 44) // 2048
                out["cli_wall_s"] = wall
                out["cli_wall_samples_s"] = [round(w, 2) for w in walls]
                out["cli_frames"] = int(n_frames)
                out["cli_sample"] = ("magpie-tts -t 'Hello, world!' --temp 0 (bf16 default), best of 3: whole process incl. CUDA context, loading the 906 MB f32 GGUF and "
                                     "the codec; audio seconds produced = frames x 1024 / 22050")
                for ln in pr.stderr.splitlines():
                    if "frames/s" in ln:
                        out["cli_loop_line"] = ln.strip()
            else:
                out["cli_error"] = pr.stderr[-300:]
    except Exception as e:  # noqa: BLE001
        out["cli_error"] = str(e)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--batch", type=int, default=0, help="utterances per GPU (default: 1 at N=1 = configs[1]; 64 at N>1 = configs[3])")
    ap.add_argument("--inproc", action="store_true", help="one process, C-ABI pool with one stream per GPU (no torch.distributed)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.inproc:
        run_inproc(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
