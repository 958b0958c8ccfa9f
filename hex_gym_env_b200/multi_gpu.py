"""Games sharded over the GPUs of one box: one process per GPU, each owning a contiguous range of global game indices.

Games are independent, so the data path needs NO collective: rank r simulates games [offset_r, offset_r + count_r) with
its own HexBatch, and because every game's random stream is keyed by (seed, GLOBAL game index) the trajectories are the
same for any number of ranks. The only exchange is the sum of the int64[8] episode statistics (K7): one 64-byte
all-reduce per report interval over NCCL (NVLink / NVSwitch), issued on a side stream so that it never sits between two
step kernels. (The reference has no distributed code at all)
"""
import os

import torch
import torch.distributed as dist

from .batch import AGENT_RANDOM, STAT_NAMES, VARIANT_B, HexBatch


def shard_range(total_games, world_size, rank):
    """Contiguous balanced partition of [0, total_games): returns (offset, count) of `rank`. The first
    total_games % world_size ranks get one game more."""
    if not (0 <= rank < world_size) or total_games < 0:
        raise ValueError("bad shard request: total=%r world=%r rank=%r" % (total_games, world_size, rank))
    base, extra = divmod(int(total_games), int(world_size))
    count = base + (1 if rank < extra else 0)
    offset = rank * base + min(rank, extra)
    return offset, count


def init_from_env(backend=None):
    """Join the process group torchrun set up (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*). Returns (rank, world, local_rank).
    A single process (no WORLD_SIZE) needs no group."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def allreduce_stats(stats, group=None, async_op=False):
    """Sum an int64[8] statistics vector over all ranks in place (works for CUDA tensors over NCCL and CPU tensors over
    gloo). Without a process group it is the identity."""
    if stats.dtype != torch.int64 or stats.numel() != len(STAT_NAMES):
        raise ValueError("stats must be int64[%d]" % len(STAT_NAMES))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        return dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
    return None


class ShardedHexBatch(object):
    """This rank's shard of a global batch of games + the global statistics."""

    def __init__(self, board_size, total_games, rank=None, world_size=None, device=None, group=None, **kw):
        if rank is None:
            rank = dist.get_rank(group) if dist.is_initialized() else 0
        if world_size is None:
            world_size = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank, self.world_size, self.group = rank, world_size, group
        self.total_games = int(total_games)
        self.offset, self.count = shard_range(total_games, world_size, rank)
        if self.count == 0:
            raise ValueError("rank %d of %d would own no games (total %d)" % (rank, world_size, total_games))
        kw.setdefault("variant", VARIANT_B)
        kw.setdefault("agent_mode", AGENT_RANDOM)
        self.local = HexBatch(board_size, self.count, device=device, game_offset=self.offset, **kw)
        self._side = torch.cuda.Stream(device=self.local.device)
        self._stats = torch.zeros(len(STAT_NAMES), dtype=torch.int64, device=self.local.device)

    def reset(self, *a, **k):
        return self.local.reset(*a, **k)

    def step(self, *a, **k):
        return self.local.step(*a, **k)

    def global_stats_begin(self):
        """Start the global sum: the statistics kernel on the caller's stream, the all-reduce on a side stream that waits for
        that kernel only. Returns at once; step kernels issued on the caller's stream afterwards are NOT held up by the
        collective (this is the form bench.py uses around its timed bracket). global_stats_end() joins."""
        main = torch.cuda.current_stream(self.local.device)
        self.local.stats(out=self._stats)
        self._side.wait_stream(main)
        with torch.cuda.stream(self._side):
            allreduce_stats(self._stats, self.group)

    def global_stats_end(self):
        """Make the caller's stream wait for the all-reduce started by global_stats_begin(); returns the device int64[8] sums."""
        torch.cuda.current_stream(self.local.device).wait_stream(self._side)
        return self._stats

    def global_stats(self):
        """Episode statistics summed over all ranks (device int64[8]): begin + end. The caller's stream waits for the 64-byte
        all-reduce here, because the result is read on it; use the begin / end pair to keep it away from the step kernels."""
        self.global_stats_begin()
        return self.global_stats_end()

    def global_stats_dict(self):
        return dict(zip(STAT_NAMES, self.global_stats().cpu().tolist()))
