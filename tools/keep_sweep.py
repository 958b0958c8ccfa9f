"""Sweep of HEXB_L2_KEEP_MB (MiB of packed state kept L2-resident across steps) on one box: graph-timed step time per setting.
Usage: python tools/keep_sweep.py [N] [G] [keep values ...]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hex_gym_env_b200 import HexBatch, VARIANT_B, AGENT_RANDOM
from bench import capture_steps
N = int(sys.argv[1]) if len(sys.argv) > 1 else 11
G = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
keeps = [int(v) for v in sys.argv[3:]] or [0, 8, 16, 20, 24, 32, 40, 48, 64]
dev = torch.device("cuda", 0)
K = 100
for rnd in range(2):
    for keep in keeps:
        os.environ["HEXB_L2_KEEP_MB"] = str(keep)      # read by hexb_create
        env = HexBatch(N, G, variant=VARIANT_B, device=0, seed=0, agent_mode=AGENT_RANDOM)
        env.reset()
        env.rollout(600, outputs=False)
        for _ in range(3):
            env.step()
        g = capture_steps(env, dev, K)
        g.replay(); torch.cuda.synchronize()
        ts = []
        for rep in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            ts.append(1e3 * e0.elapsed_time(e1) / K)
        print(json.dumps({"N": N, "G": G, "keep_mb": keep, "round": rnd, "us_per_step_min": round(min(ts), 2), "us_per_step_med": round(sorted(ts)[2], 2)}), flush=True)
        env.close(); del env, g
