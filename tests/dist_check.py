"""Multi-GPU check, run under torchrun on a box with >= 2 GPUs (not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py

Every rank simulates its shard of ONE global batch on its own GPU and compares it bit for bit with the CPU oracle run on
the same global game indices; the NCCL all-reduced episode statistics must equal the oracle's statistics of the whole batch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hex_gym_env_b200.multi_gpu import ShardedHexBatch, init_from_env  # noqa: E402
from oracle import hexref  # noqa: E402


def main():
    rank, world, local = init_from_env("nccl")
    N, total, T, seed = 11, 10007, 120, 5
    sh = ShardedHexBatch(N, total, device=local, seed=seed)
    ref = hexref.RefBatch(hexref.KIND_SELFPLAY_B, N, sh.count, seed=seed, game_offset=sh.offset, agent_mode=2)
    o, m = sh.reset()
    ro, rm = ref.reset()
    assert np.array_equal(o.cpu().numpy(), ro) and np.array_equal(m.cpu().numpy(), rm)
    for t in range(T):
        out = sh.step(want_actions=True)
        r = ref.step()
        for k in ("obs", "mask", "reward", "done", "actions"):
            assert np.array_equal(out[k].cpu().numpy(), r[k]), "rank %d: %s differs at step %d" % (rank, k, t)
    g = sh.global_stats().cpu().numpy()
    mine = torch.from_numpy(ref.stats().copy()).cuda(local)
    dist.all_reduce(mine)
    assert np.array_equal(g, mine.cpu().numpy()), (g, mine)
    if rank == 0:
        whole = hexref.RefBatch(hexref.KIND_SELFPLAY_B, N, total, seed=seed, agent_mode=2)
        whole.reset()
        for _ in range(T):
            whole.step_fast()
        assert np.array_equal(g, whole.stats()), (g, whole.stats())
        print("dist_check ok: %d ranks, %d games, %d steps, global stats %s" % (world, total, T, g.tolist()))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
