"""Workload for ncu captures of one configuration: reset, pre-roll, then `steps` single-step launches (optionally a forced
launch form). Usage: python tools/prof_small.py N G variant(A|B) form steps [rollout_T]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hex_gym_env_b200 import HexBatch, VARIANT_A, VARIANT_B, AGENT_RANDOM, AGENT_BLACK
N, G, var, form, steps = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
T = int(sys.argv[6]) if len(sys.argv) > 6 else 0
env = HexBatch(N, G, variant=VARIANT_A if var == "A" else VARIANT_B, device=0, seed=0, agent_mode=AGENT_BLACK if var == "A" else AGENT_RANDOM)
env.set_launch_form(form)
env.reset()
env.rollout(300, outputs=False)
for _ in range(steps):
    env.step()
if T:
    env.rollout(T)
torch.cuda.synchronize()
print("done")
