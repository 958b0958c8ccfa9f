#!/usr/bin/env python
"""bench.py - env-steps/s of 11x11 random self-play (BASELINE.json config 3) on N B200s, with the HBM roofline of the
fused step kernel and the CPU restatement of the reference loop timed beside it.

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path (one JSON line on rank 0)
  python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]   # the CPU reference arm (rank 0 only)

A "step" is one fused env step (agent ply + random-opponent reply + win check + reward/done + auto-reset + observation
and legal-action mask) over ALL games of the rank's shard: variant B (SelfPlayEnv + BaseRandomPolicy), agent colour
random per game, the agent itself a random policy drawing from the game's Philox stream ("random self-play").
Weak scaling: every GPU owns `--games-per-gpu` games (default 1,048,576 = config 3's one million games; at N GPUs the job
is N million games, sharded by global game index with no data-path collective; the only exchange is one NCCL all-reduce
of the int64[8] episode statistics).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BOARD = 11
GAMES_PER_GPU = 1 << 20
METRIC = "env_steps_per_sec_11x11_random_selfplay"
UNIT = "env-steps/s"


def contract_bytes(N):
    """B(N) = 4 + 2*S(N) + 2*C + 5, S(N) = 8W + 2C + 16 (algorithmic bytes per env-step per game)."""
    C = N * N
    W = (C + 31) // 32
    return 4 + 2 * (8 * W + 2 * C + 16) + 2 * C + 5


def moved_bytes(N, sampled=True):
    """Bytes this implementation really moves per env-step per game: packed state in + out (C label bytes + (W+2) u32
    record words), obs + mask + reward + done out (+ actions in when the agent is external)."""
    C = N * N
    W = (C + 31) // 32
    S = C + 4 * (W + 2)
    return 2 * S + 2 * C + 5 + (0 if sampled else 4)


class ClockSampler(object):
    """Samples SM clock + throttle reasons with NVML while the timed region runs."""

    BAD = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown"}
    NOTE = {0x4: "sw_power_cap", 0x80: "hw_power_brake", 0x2: "applications_clocks", 0x100: "display_clock", 0x10: "sync_boost"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self._stop = [], set(), None, threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as exc:  # no NVML: report that instead of inventing clocks
            self.nv, self.err = None, str(exc)

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in list(self.BAD.items()) + list(self.NOTE.items()):
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.01)

    def start(self):
        if self.nv:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join()
        if not self.nv:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml_unavailable"]}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ---------------------------------------------------------------------------------------------------- CPU legs
def cpu_port_rate(N, budget_s, threads):
    """The C restatement of the reference loop (oracle/hexref.c), `threads` pthreads over independent games."""
    import numpy as np  # noqa: F401
    from oracle import hexref
    hexref.set_threads(threads)
    G = 2048 * threads
    b = hexref.RefBatch(hexref.KIND_SELFPLAY_B, N, G, seed=0, agent_mode=2)
    b.reset()
    for _ in range(3):
        b.step()
    n, t0 = 0, time.perf_counter()
    while True:
        b.step()          # obs + mask + reward + done written every step, like the reference loop computes them
        n += 1
        dt = time.perf_counter() - t0
        if dt >= budget_s:
            break
    return G * n / dt, "%d games x %d steps (C port, %d threads)" % (G, n, threads)


def py_rate(N, seconds, procs):
    """oracle/pyloop.py in a child interpreter (keeps fork() away from this process's CUDA context)."""
    import subprocess
    out = subprocess.run([sys.executable, "-m", "oracle.pyloop", str(N), str(seconds), str(procs)], cwd=ROOT, check=True,
                         stdout=subprocess.PIPE, text=True, env=dict(os.environ, OMP_NUM_THREADS="1")).stdout
    j = json.loads(out.strip().splitlines()[-1])
    return j["value"], j["sample"]


def cpu_baseline(N, budget_s=12.0):
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    v, sample = py_rate(N, budget_s, cores)
    p1, _ = py_rate(N, min(budget_s, 4.0), 1)
    c1, _ = cpu_port_rate(N, 2.0, 1)
    cn, _ = cpu_port_rate(N, 3.0, cores)
    return {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "python_loop_1_core": p1,
            "c_port_1_thread": c1, "c_port_all_threads": cn}


def run_reference(args):
    """Reference arm: the reference's CPU loop (mask -> random action -> SelfPlayEnv.step, reset on done) restated in
    Python/numpy (oracle/pyloop.py, same per-step numpy work as minihex), one process per host core."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    steps = min(args.steps, 10)          # bounded: every step is a per_step_s sample on all cores
    per_step_s = 2.0
    for _ in range(min(max(args.warmup, 0), 3)):
        py_rate(BOARD, 0.3, cores)
    t0 = time.perf_counter()
    samples = []
    for _ in range(steps):
        v, sample = py_rate(BOARD, per_step_s, cores)
        samples.append(v)
    dt = time.perf_counter() - t0
    v = sum(samples) / len(samples)
    c_all, c_sample = cpu_port_rate(BOARD, 3.0, cores)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "samples_run": steps, "ms_per_step": 1e3 * dt / max(steps, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "11x11 SelfPlayEnv random self-play (BASELINE config 3), reference CPU loop",
                       "board_size": BOARD, "note": "each step = %.1f s sample of the loop on every host core" % per_step_s},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "python/numpy restatement of the minihex loop, %d processes x %.1f s x %d steps"
                                       % (cores, per_step_s, steps),
                             "c_port_all_threads": c_all, "c_port_sample": c_sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), file=args.out, flush=True)


# ---------------------------------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from hex_gym_env_b200 import HexBatch, VARIANT_B, AGENT_RANDOM

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = "unchanged"
    try:  # run this rank (and first-touch its pinned buffers) on the CPUs next to its GPU: the end-to-end leg is PCIe / host-memory bound
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
        numa = "nvmlDeviceSetCpuAffinity(gpu %d): %d cpus" % (local, len(os.sched_getaffinity(0)))
    except Exception as exc:
        numa = "not set (%s)" % (str(exc)[:60],)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    N, G, K, Wm = args.board, args.games_per_gpu, args.steps, args.warmup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ------------------------------------------------ device-resident leg (value + roofline)
    env = HexBatch(N, G, variant=VARIANT_B, device=local, seed=args.seed, game_offset=rank * G, agent_mode=AGENT_RANDOM,
                   auto_reset=True)
    env.reset()
    stats = torch.zeros(8, dtype=torch.int64, device=dev)
    for _ in range(Wm):
        env.step()
    # The step loop is captured in a CUDA graph (GRAPH_STEPS launches per replay) so that the ~3 us launch gap of a Python
    # loop does not sit between 110 us kernels; K steps = K kernel launches either way.
    GRAPH_STEPS = 50
    graph = None
    if not args.no_graph and K >= 2 * GRAPH_STEPS:
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            env.step()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                for _ in range(GRAPH_STEPS):
                    env.step()
        torch.cuda.current_stream(dev).wait_stream(side)
        graph.replay()      # (warm-up: these 51 steps are outside the timed region)

    def run_steps(n):
        if graph is not None:
            for _ in range(n // GRAPH_STEPS):
                graph.replay()
            n %= GRAPH_STEPS
        for _ in range(n):
            env.step()

    env.stats(out=stats)
    if world > 1:
        dist.all_reduce(stats)
    s0 = stats.clone()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    e0.record()
    run_steps(K)
    env.stats(out=stats)
    if world > 1:
        dist.all_reduce(stats)      # K7: the only collective of the path (64 bytes)
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = K + 1
    value = world * G * K / (ms * 1e-3)
    ds = (stats - s0).cpu().tolist()

    # kernel-only duration for the roofline: the K step launches alone, CUDA events on the launching stream
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    k0.record()
    run_steps(K)
    k1.record()
    torch.cuda.synchronize()
    kern_ms = k0.elapsed_time(k1) / K
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    B = contract_bytes(N)
    achieved = G * B / (kern_ms * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        ent = tj.get("N%d_G%d" % (N, G))
        if ent:
            traffic = ent["dram_bytes_per_launch"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": "hexb_step_kernel<%d, KIND_STEP, %s>" % (N, "several-rows sweep" if (G + 127) // 128 * 4 <= 32 * torch.cuda.get_device_properties(dev).multi_processor_count else "one-row sweep"), "kernel_ms": kern_ms,
                "bytes_per_env_step_contract": B, "bytes_per_env_step_moved": moved_bytes(N),
                "achieved_moved": G * moved_bytes(N) / (kern_ms * 1e-3) / 1e9,
                "frac_moved": G * moved_bytes(N) / (kern_ms * 1e-3) / 1e9 / peak,
                "note": "achieved/frac use SURVEY 8(d)'s algorithmic bytes (state 290 B per game each way); this implementation "
                        "packs the state into 145 B, so the bytes the kernel really requests (achieved_moved/frac_moved) are lower "
                        "and frac can exceed 1; 20 MiB of the state are kept L2-resident across steps, so HBM sees slightly less "
                        "than that again (traffic = ncu dram bytes per launch, measured with L2 flushed before the launch)",
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"}
    del env

    # ------------------------------------------------ end-to-end leg: host actions in, host obs/mask/reward/done out
    E = min(K, args.e2e_steps)
    Ew = 3
    # (untimed) record a legal random trajectory on the device so that the timed loop replays HOST actions
    rec = HexBatch(N, G, variant=VARIANT_B, device=local, seed=args.seed + 1, game_offset=rank * G, agent_mode=AGENT_RANDOM,
                   auto_reset=True)
    rec.reset()
    host_actions = torch.empty((E + Ew, G), dtype=torch.int32).pin_memory()
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    for t in range(E + Ew):
        u = torch.rand(G, dtype=torch.float64, device=dev, generator=gen)
        a = rec.sample_actions(u)
        host_actions[t].copy_(a)
        rec.step(a, outputs=False)
    torch.cuda.synchronize()
    del rec
    env = HexBatch(N, G, variant=VARIANT_B, device=local, seed=args.seed + 1, game_offset=rank * G, agent_mode=AGENT_RANDOM,
                   auto_reset=True)
    env.reset()
    io = env.pinned_io()
    for t in range(Ew):
        io["actions"].copy_(host_actions[t])
        env.step_host(io["actions"], io)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    checksum = 0.0
    for t in range(Ew, Ew + E):
        io["actions"].copy_(host_actions[t])           # this step's inputs, host memory
        env.step_host(io["actions"], io)               # H2D + kernel + D2H, returns with results in host memory
        checksum += float(io["reward"][0])             # host read of the step's result
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1))
    wall_ms = (time.perf_counter() - t0) * 1e3
    launches_e2e = E
    invalid = int(env.stats().cpu()[5])
    e2e = {"value": world * G * E / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 4 * G,
           "d2h_bytes_per_step": G * (2 * N * N + 5), "steps": E, "ms_per_step": e2e_ms / E, "wall_ms_per_step": wall_ms / E,
           "api": "hexb_step_host (C ABI, pinned host buffers)", "illegal_moves_in_replay": invalid, "cpu_affinity": numa}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "11x11 SelfPlayEnv (variant B) random self-play, random opponent, auto-reset (BASELINE config 3)"
                       if N == 11 else "%dx%d SelfPlayEnv random self-play" % (N, N),
                       "board_size": N, "games_per_gpu": G, "global_games": world * G, "parallelism": "games sharded by index x%d" % world,
                       "l2_policy": "working set %.0f MB per step > 126 MB L2" % (G * moved_bytes(N) / 1e6),
                       "agent": "fused on-device random policy (Philox stream per game)", "seed": args.seed,
                       "launch": ("CUDA graph of %d step launches per replay" % GRAPH_STEPS) if graph is not None else "one launch per step"},
            "roofline": roofline, "e2e": e2e, "gpu_launches": launches, "gpu_launches_detail": {"timed_region": "%d x hexb_step_kernel + 1 x hexb_stats_kernel" % K, "roofline_region": K, "e2e_region": launches_e2e}, "clocks": clocks,
            "plies_per_sec": ds[7] / (ms * 1e-3), "episodes_in_timed_region": ds[0],
            "episode_stats": dict(zip(("episodes", "black_wins", "white_wins", "agent_wins", "episode_plies", "invalid_ends",
                                       "env_steps", "plies"), ds))}
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(N, args.cpu_budget)
        print(json.dumps(line), file=args.out, flush=True)
    if world > 1:
        dist.destroy_process_group()


def _claim_stdout():
    """Keep stdout for the ONE JSON line: libraries (NCCL prints "NCCL version ..." there when NCCL_DEBUG is set) write to
    file descriptor 1, so point descriptor 1 at stderr and return a private handle to the real stdout."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--board", type=int, default=BOARD)
    ap.add_argument("--games-per-gpu", type=int, default=GAMES_PER_GPU)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=40)
    ap.add_argument("--cpu-budget", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every step from Python instead of replaying a CUDA graph")
    args = ap.parse_args()
    args.out = _claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
