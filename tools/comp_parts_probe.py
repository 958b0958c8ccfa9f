"""Which buffers should be compressible? Graph-timed step time with the packed state and / or the outputs in compressible memory.
Run once per setting: HEXB_COMPRESSIBLE=1 HEXB_COMPRESSIBLE_PARTS=state|outputs|state,outputs (or HEXB_COMPRESSIBLE=0)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hex_gym_env_b200 import HexBatch, VARIANT_B, AGENT_RANDOM
from bench import capture_steps
dev = torch.device("cuda", 0)
tag = "none" if os.environ.get("HEXB_COMPRESSIBLE") == "0" else os.environ.get("HEXB_COMPRESSIBLE_PARTS", "state,outputs")
for N, G, K in ((19, 1 << 20, 40), (11, 1 << 20, 100), (19, 1 << 22, 20)):
    env = HexBatch(N, G, variant=VARIANT_B, device=0, seed=0, agent_mode=AGENT_RANDOM)
    env.reset(); env.rollout(400, outputs=False)
    for _ in range(3): env.step()
    g = capture_steps(env, dev, K); g.replay(); torch.cuda.synchronize()
    ts = []
    for rep in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(1e3 * e0.elapsed_time(e1) / K)
    print(json.dumps({"compressible": tag, "N": N, "G": G, "us_min": round(min(ts), 1), "us_med": round(sorted(ts)[3], 1), "us_max": round(max(ts), 1)}), flush=True)
    env.close(); del env, g
