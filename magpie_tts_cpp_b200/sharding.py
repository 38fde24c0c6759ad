"""Multi-GPU synthesis: independent utterances are sharded across devices (utterance i -> device i mod G,
`mgb_shard_device`), weights are replicated, and there is NO collective on the data path (SURVEY.md 8e;
the reference itself is single-utterance, single-device).  Two front ends:

* `MultiGpuSynthesizer` -- one process, one model replica + one host submission thread per GPU
  (the "process-less stream per GPU" form named in BASELINE.json's north_star);
* `local_shard` / `gather_in_order` -- one process per GPU under torchrun (what bench.py --gpus N uses); the only
  communication is the final host-side gather of codes, done with torch.distributed on CPU tensors/objects.
"""
from __future__ import annotations

import threading

import numpy as np

from . import binding


def shard_plan(n_utterances: int, world: int) -> "list[list[int]]":
    """Utterance indices per device/rank; uses the library's own assignment rule."""
    L = binding.lib()
    plan = [[] for _ in range(world)]
    for i in range(n_utterances):
        plan[L.mgb_shard_device(i, world)].append(i)
    return plan


def local_shard(n_utterances: int, world: int, rank: int) -> "list[int]":
    return shard_plan(n_utterances, world)[rank]


def gather_in_order(local: "dict[int, object]", n_utterances: int, group=None):
    """Gather {utterance index: result} from all ranks onto rank 0, ordered by utterance index.
    Works with any backend (gloo on CPU, nccl ranks gather python objects through the store)."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return [local[i] for i in range(n_utterances)]
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    parts = [None] * world if rank == 0 else None
    dist.gather_object(local, parts, dst=0, group=group)
    if rank != 0:
        return None
    merged = {}
    for p in parts:
        dup = set(merged) & set(p)
        if dup:
            raise RuntimeError(f"utterances assigned to more than one rank: {sorted(dup)}")
        merged.update(p)
    missing = [i for i in range(n_utterances) if i not in merged]
    if missing:
        raise RuntimeError(f"utterances not produced by any rank: {missing}")
    return [merged[i] for i in range(n_utterances)]


class MultiGpuSynthesizer:
    """One model (+ optional codec) replica per device; `synthesize` runs every device's shard concurrently."""

    def __init__(self, model_path: str, codec_path: str | None = None, devices=None, precision=binding.PREC_BF16):
        n = binding.device_count()
        if n == 0:
            raise binding.MagpieError("no CUDA device available (no CPU fallback)")
        self.devices = list(devices) if devices is not None else list(range(n))
        self.models = [binding.Model(model_path, d, precision) for d in self.devices]
        self.codecs = [binding.Codec(codec_path, d) for d in self.devices] if codec_path else None

    def synthesize(self, token_lists, speakers=None, max_steps=0, temperature=0.0, top_k=80, seed=0, ignore_eos=False):
        """token_lists: list of token-id lists. Returns list of [n_frames][8] code arrays, in input order."""
        n = len(token_lists)
        speakers = list(speakers) if speakers is not None else [0] * n
        plan = shard_plan(n, len(self.devices))
        results, errors = {}, []

        def work(di):
            try:
                idx = plan[di]
                if not idx:
                    return
                m = self.models[di]
                toks = [token_lists[i] for i in idx]
                s = m.session(batch=len(idx), max_text=max(len(t) for t in toks),
                              max_seq=m.hp["context_frames"] + (max_steps or m.hp["max_dec_steps"]) + 16)
                s.encode_text(toks, want_output=False)
                s.prefill([speakers[i] for i in idx])
                out = s.generate(max_steps=max_steps, temperature=temperature, top_k=top_k, seed=seed, ignore_eos=ignore_eos)
                for j, i in enumerate(idx):
                    results[i] = out[j]
                s.close()
            except Exception as e:  # noqa: BLE001
                errors.append(e)

        threads = [threading.Thread(target=work, args=(di,)) for di in range(len(self.devices))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        return [results[i] for i in range(n)]
