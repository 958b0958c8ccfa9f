"""TEST INFRASTRUCTURE ONLY - the UNMODIFIED reference's random self-play loop, timed on host cores.

This is the CPU arm bench.py reports (`cpu_baseline.kind = "reference"`, `--impl reference`): the reference's own classes,
imported from the byte-identical copy under oracle/_ref (oracle/make_ref.py), driven through their stock code path exactly
like the reference's training scripts drive them (scripts/experiments/6x6_MLP-default_lr-0.0003.py:31-38 + SB3's loop):

    env  = selfplay_wrapper(HexEnv)(board_size=N)            minihex/SelfplayWrapper.py:37-67
    mask = env.legal_actions()                               minihex/HexSingleGame.py:205-206   (ActionMasker's mask_fn, every step)
    a    = BaseRandomPolicy().choose_action(obs)             minihex/SelfplayWrapper.py:17-22
    obs, r, done, _, _ = env.step(a)                         minihex/SelfplayWrapper.py:174-199
    if done: obs, _ = env.reset()                            minihex/SelfplayWrapper.py:69-89   (DummyVecEnv's auto-reset)

The reference is single-threaded; `procs` independent envs (one process per host core) give the aggregate rate.

    python -m oracle.ref_loop N seconds procs       ->  one JSON line {"value", "sample", "procs", "kind"}
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def available():
    from oracle import make_ref
    return make_ref.present()


def loop(n, seconds, seed=0):
    """One env, `seconds` of the loop above. Returns (env steps done, seconds used)."""
    import random
    from oracle import ref_harness
    ref_harness.use_copy()
    _, _, B, S = ref_harness.load()
    random.seed(seed)                       # the reference draws from the global `random`
    env = S.selfplay_wrapper(B.HexEnv)(board_size=n)
    pol = S.BaseRandomPolicy()
    obs, _ = env.reset()
    steps, t0 = 0, time.perf_counter()
    while True:
        env.legal_actions()
        obs, _, done, _, _ = env.step(pol.choose_action(obs))
        steps += 1
        if done:
            obs, _ = env.reset()
        if (steps & 63) == 0:
            dt = time.perf_counter() - t0
            if dt >= seconds:
                return steps, dt


def _worker(args):
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    return loop(*args)


def rate(n, seconds, procs):
    if procs <= 1:
        s, dt = loop(n, seconds, 0)
        return s / dt, "1 process x %.1f s, %d env steps (unmodified minihex SelfPlayEnv loop)" % (dt, s)
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    with ctx.Pool(procs) as pool:
        res = pool.map(_worker, [(n, seconds, i) for i in range(procs)])
    return (sum(s / dt for s, dt in res),
            "%d processes x %.1f s, %d env steps in total (unmodified minihex SelfPlayEnv loop)" % (procs, seconds, sum(s for s, _ in res)))


if __name__ == "__main__":
    import json
    _n, _sec, _procs = int(sys.argv[1]), float(sys.argv[2]), int(sys.argv[3])
    _v, _s = rate(_n, _sec, _procs)
    print(json.dumps({"value": _v, "sample": _s, "procs": _procs, "kind": "reference"}))
