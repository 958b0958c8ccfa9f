"""Build-container only (needs /root/reference): speed of the UNMODIFIED reference's random self-play loop next to the Python
restatement bench.py times on the GPU box (oracle/pyloop.py), same board size, one core each. Records how conservative the
reported cpu_baseline is."""
import os, sys, time, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_harness, pyloop
N = int(sys.argv[1]) if len(sys.argv) > 1 else 11
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 5.0
minihex, A, B, S = ref_harness.load()
env = S.selfplay_wrapper(B.HexEnv)(board_size=N)
pol = S.BaseRandomPolicy()
obs, _ = env.reset()
n, t0 = 0, time.perf_counter()
while time.perf_counter() - t0 < secs:
    mask = env.legal_actions()                       # what ActionMasker's mask_fn calls every step
    obs, r, done, _, _ = env.step(pol.choose_action(obs))
    if done:
        obs, _ = env.reset()
    n += 1
ref = n / (time.perf_counter() - t0)
s, d = pyloop.loop(N, secs)
print("N=%d: unmodified reference %.0f env-steps/s, oracle/pyloop.py %.0f env-steps/s (1 core each): ratio %.2f" % (N, ref, s / d, (s / d) / ref))
