"""A/B of two builds of libhexb.so on the same box (HEXB_LIB selects the build; run once per build): graph-timed step time
of a few configurations. Usage: python tools/ab_probe.py [tag]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hex_gym_env_b200 import HexBatch, VARIANT_A, VARIANT_B, AGENT_RANDOM, AGENT_BLACK
from bench import capture_steps
tag = sys.argv[1] if len(sys.argv) > 1 else "build"
dev = torch.device("cuda", 0)
CONFIGS = [("11x11 32768", 11, 32768, VARIANT_B, AGENT_RANDOM, 200), ("11x11 65536", 11, 65536, VARIANT_B, AGENT_RANDOM, 200),
           ("7x7 131072 A", 7, 131072, VARIANT_A, AGENT_BLACK, 200), ("19x19 65536", 19, 65536, VARIANT_B, AGENT_RANDOM, 100),
           ("6x6 4096", 6, 4096, VARIANT_B, AGENT_RANDOM, 200), ("7x7 65536 A", 7, 65536, VARIANT_A, AGENT_BLACK, 200),
           ("11x11 131072", 11, 131072, VARIANT_B, AGENT_RANDOM, 200), ("11x11 1Mi", 11, 1 << 20, VARIANT_B, AGENT_RANDOM, 100),
           ("19x19 1Mi", 19, 1 << 20, VARIANT_B, AGENT_RANDOM, 40)]
for name, N, G, variant, am, K in CONFIGS:
    env = HexBatch(N, G, variant=variant, device=0, seed=0, agent_mode=am)
    env.reset()
    env.rollout(600, outputs=False)
    for _ in range(3):
        env.step()
    g = capture_steps(env, dev, K)
    g.replay(); torch.cuda.synchronize()
    ts = []
    for rep in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(1e3 * e0.elapsed_time(e1) / K)
    T = 32
    env.rollout(T, outputs=False); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        env.rollout(T, outputs=False)
    e1.record(); torch.cuda.synchronize()
    print(json.dumps({"build": tag, "config": name, "us_per_step_min": round(min(ts), 3), "us_per_step_med": round(sorted(ts)[2], 3),
                      "rollout_noout_us_per_step": round(1e3 * e0.elapsed_time(e1) / (4 * T), 3)}), flush=True)
    env.close(); del env, g
