"""Single-env drop-in speed (one game on the GPU, the reference's own rollout loop) next to the Python restatement of the reference."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hex_gym_env_b200.minihex.SelfplayWrapper import selfplay_wrapper, BaseRandomPolicy
from hex_gym_env_b200.minihex.HexSingleGame import HexEnv
from oracle import pyloop
for N in (5, 11):
    env = selfplay_wrapper(HexEnv)(board_size=N)
    obs, _ = env.reset()
    pol = BaseRandomPolicy()
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < 3.0:
        mask = env.legal_actions()
        obs, r, done, _, _ = env.step(pol.choose_action(obs))
        if done:
            obs, _ = env.reset()
        n += 1
    dt = time.perf_counter() - t0
    s, d = pyloop.loop(N, 3.0)
    print("N=%d: drop-in single env %.0f env-steps/s; python restatement of the reference %.0f env-steps/s" % (N, n / dt, s / d))
