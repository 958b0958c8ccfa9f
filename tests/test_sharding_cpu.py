"""CPU: host-side multi-GPU logic. The shard arithmetic, and - with two real processes over gloo - the episode
statistics all-reduce plus the sharding invariance it relies on (per-game streams keyed by the GLOBAL game index), with
the CPU oracle standing in for each rank's simulator (there is no GPU in this container)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hex_gym_env_b200.multi_gpu import allreduce_stats, shard_range


def test_shard_range_partitions():
    for total in (1, 7, 8, 1000, 1 << 20, (1 << 20) + 5):
        for world in (1, 2, 3, 4, 8):
            if total < world:
                continue
            spans = [shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (o0, c0), (o1, _) in zip(spans, spans[1:]):
                assert o0 + c0 == o1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_allreduce_identity_without_group():
    s = torch.arange(8, dtype=torch.int64)
    assert allreduce_stats(s) is None and s.tolist() == list(range(8))
    with pytest.raises(ValueError):
        allreduce_stats(torch.zeros(8, dtype=torch.int32))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, N, T, seed, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import hexref
    off, cnt = shard_range(total, world, rank)
    b = hexref.RefBatch(hexref.KIND_SELFPLAY_B, N, cnt, seed=seed, game_offset=off, agent_mode=2)
    b.reset()
    last = None
    for _ in range(T):
        last = b.step()
    stats = torch.from_numpy(b.stats().copy())
    allreduce_stats(stats)
    q.put((rank, off, cnt, stats.tolist(), last["obs"], last["reward"]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_gloo_stats_and_sharding_invariance():
    total, N, T, seed, world = 301, 5, 40, 11, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, N, T, seed, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from oracle import hexref
    whole = hexref.RefBatch(hexref.KIND_SELFPLAY_B, N, total, seed=seed, agent_mode=2)
    whole.reset()
    for _ in range(T):
        last = whole.step()
    want = whole.stats().tolist()
    assert got[0][3] == want and got[1][3] == want          # all-reduced statistics == the unsharded run's
    obs = np.concatenate([g[4] for g in got])
    rew = np.concatenate([g[5] for g in got])
    assert np.array_equal(obs, last["obs"]) and np.array_equal(rew, last["reward"])   # same trajectories, any sharding
