"""Multi-GPU host logic on CPU: utterance sharding + ordered gather with world_size 2 over gloo.
Each rank produces its shard's greedy codes with the CPU oracle (the checker standing in for the device
path, which needs a GPU); rank 0 verifies the gathered result equals the sequential run."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

HELLO = [2378, 7, 4, 11, 11, 14, 32, 26, 22, 14, 17, 11, 3, 32, 28, 2379]


def test_shard_plan_partitions():
    from magpie_tts_cpp_b200 import sharding
    for n, w in [(0, 2), (1, 2), (5, 2), (64, 8), (7, 3)]:
        plan = sharding.shard_plan(n, w)
        flat = sorted(i for p in plan for i in p)
        assert flat == list(range(n))
        assert max(len(p) for p in plan) - min(len(p) for p in plan) <= 1
        for r in range(w):
            assert plan[r] == [i for i in range(n) if i % w == r] == sharding.local_shard(n, w, r)
    from magpie_tts_cpp_b200 import binding
    assert binding.lib().mgb_shard_device(3, 0) < 0 and binding.lib().mgb_shard_device(-1, 2) < 0


def _worker(rank, world, port, model_path, q):
    import torch.distributed as dist
    from magpie_tts_cpp_b200 import sharding
    from oracle import oracle
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    utts = [HELLO, HELLO[:9] + [2379], HELLO[:5] + [2379], [2378, 3, 4, 2379], HELLO[:12] + [2379]]
    o = oracle.OracleModel(model_path)
    local = {i: o.synthesize(utts[i], speaker=i % 2, temperature=0.0, max_steps=6) for i in sharding.local_shard(len(utts), world, rank)}
    out = sharding.gather_in_order(local, len(utts))
    if rank == 0:
        ref = [o.synthesize(utts[i], speaker=i % 2, temperature=0.0, max_steps=6) for i in range(len(utts))]
        q.put(all(np.array_equal(a, b) for a, b in zip(out, ref)) and len(out) == len(utts))
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_synthesis(tiny_model_path, oracle_mod):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, tiny_model_path, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True


@pytest.mark.gpu
def test_multi_gpu_synthesizer_matches_single_session(tiny_model_path, oracle_mod):
    from magpie_tts_cpp_b200 import binding, sharding
    utts = [HELLO, HELLO[:9] + [2379], HELLO[:5] + [2379]]
    syn = sharding.MultiGpuSynthesizer(tiny_model_path, precision=binding.PREC_F32)
    out = syn.synthesize(utts, speakers=[0, 1, 0], max_steps=10)
    o = oracle_mod.OracleModel(tiny_model_path)
    for i, u in enumerate(utts):
        ref = o.synthesize(u, speaker=[0, 1, 0][i], temperature=0.0, max_steps=10)
        assert out[i].shape == ref.shape and np.mean(np.all(out[i] == ref, axis=1)) >= 0.99


def test_pool_fails_loudly_without_a_device(tiny_model_path):
    """The in-process pool has no CPU fallback either."""
    import torch
    from magpie_tts_cpp_b200 import binding
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(binding.MagpieError, match="no CUDA device"):
        binding.Pool(tiny_model_path)
    with pytest.raises(binding.MagpieError):
        binding.Pool(tiny_model_path, devices=[0, 1])


@pytest.mark.gpu
def test_inprocess_pool_matches_oracle_and_single_session(tiny_model_path, oracle_mod):
    """mgb_pool_* (C++: one replica + one submission thread + one stream per device slot, utterance i on slot i mod G).  Two
    slots on device 0 exercise the sharding, threading and scatter logic on a one-GPU box; results must equal the oracle's and
    a plain single-session run, in utterance order."""
    from magpie_tts_cpp_b200 import binding
    utts = [HELLO, HELLO[:9] + [2379], HELLO[:5] + [2379], [2378, 3, 4, 2379], HELLO[:12] + [2379]]
    spk = [0, 1, 0, 1, 0]
    pool = binding.Pool(tiny_model_path, devices=[0, 0], precision=binding.PREC_F32)
    assert pool.n_devices == 2
    out = pool.generate(utts, speakers=spk, max_steps=10)
    o = oracle_mod.OracleModel(tiny_model_path)
    for i, u in enumerate(utts):
        ref = o.synthesize(u, speaker=spk[i], temperature=0.0, max_steps=10)
        assert out[i].shape == ref.shape and np.mean(np.all(out[i] == ref, axis=1)) >= 0.99
    assert (pool.last_device_ms > 0).all()
    codes = np.random.default_rng(3).integers(0, 2016, (len(utts), 6, 8)).astype(np.int32)
    gr = pool.teacher_forced(utts, codes, speakers=spk)
    m = binding.Model(tiny_model_path, 0, binding.PREC_F32)
    s = m.session(batch=len(utts), max_text=16)
    s.encode_text(utts, want_output=False)
    s.prefill(spk)
    _, _, gr1 = s.teacher_forced(codes, want_hidden=False, want_logits=False)
    assert np.mean(gr == gr1) >= 0.99
    with pytest.raises(binding.MagpieError):
        pool.generate([list(range(5000))], max_steps=4)          # text longer than the encoder position table
    pool.close()
