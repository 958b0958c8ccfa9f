"""Probe: is the step loop launch-bound? Times K fused steps (a) as plain launches from Python, (b) as a captured CUDA graph."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hex_gym_env_b200 import HexBatch, VARIANT_B, AGENT_RANDOM

N = int(sys.argv[1]) if len(sys.argv) > 1 else 11
G = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
K = 400
env = HexBatch(N, G, variant=VARIANT_B, device=0, seed=0, agent_mode=AGENT_RANDOM)
env.reset()
for _ in range(100):
    env.step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(K):
    env.step()
e1.record()
t_issue = time.perf_counter() - t0
torch.cuda.synchronize()
print("plain: %.1f us/step device, %.1f us/step host issue" % (1e3 * e0.elapsed_time(e1) / K, 1e6 * t_issue / K))
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    env.step()
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s):
        for _ in range(50):
            env.step()
torch.cuda.synchronize()
g.replay()
torch.cuda.synchronize()
e0.record()
for _ in range(K // 50):
    g.replay()
e1.record()
torch.cuda.synchronize()
print("graph: %.1f us/step device" % (1e3 * e0.elapsed_time(e1) / K))
# no outputs (pure simulation)
e0.record()
for _ in range(K):
    env.step(outputs=False)
e1.record()
torch.cuda.synchronize()
print("no-output: %.1f us/step device" % (1e3 * e0.elapsed_time(e1) / K))
# multi-step launches (hexb_rollout): state stays on chip, outputs [T,G,..]
for T in (4, 16, 32):
    if T * G * (2 * N * N + 5) > 40e9:
        continue
    env.rollout(T)
    torch.cuda.synchronize()
    reps = max(1, 320 // T)
    e0.record()
    for _ in range(reps):
        env.rollout(T)
    e1.record()
    torch.cuda.synchronize()
    print("rollout T=%d: %.1f us/step device" % (T, 1e3 * e0.elapsed_time(e1) / (reps * T)))
