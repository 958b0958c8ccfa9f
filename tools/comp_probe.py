"""Experiment: the step kernel with ALL device memory compressible (tools/comp_alloc.cpp as torch's allocator) vs ordinary memory.
HEXB_COMP=1/0 python tools/comp_probe.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
so = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libcomp_alloc.so")
alloc = torch.cuda.memory.CUDAPluggableAllocator(so, "comp_malloc", "comp_free")
torch.cuda.memory.change_current_allocator(alloc)
from hex_gym_env_b200 import HexBatch, VARIANT_B, AGENT_RANDOM
from bench import capture_steps
dev = torch.device("cuda", 0)
tag = "comp%s" % os.environ.get("HEXB_COMP", "1")
# plain copy of obs-like bytes first: what the compression does for a streaming write
x = torch.randint(-1, 2, (1 << 28,), dtype=torch.int8, device=dev)
y = torch.empty_like(x)
for _ in range(3): y.copy_(x)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): y.copy_(x)
e1.record(); torch.cuda.synchronize()
print(json.dumps({"build": tag, "copy_256MB_GBps": round(2 * x.numel() * 10 / (e0.elapsed_time(e1) * 1e6), 1)}), flush=True)
z = torch.zeros_like(x)
e0.record()
for _ in range(10): y.copy_(z)
e1.record(); torch.cuda.synchronize()
print(json.dumps({"build": tag, "copy_zeros_GBps": round(2 * x.numel() * 10 / (e0.elapsed_time(e1) * 1e6), 1)}), flush=True)
del x, y, z
for name, N, G, K in (("11x11 1Mi", 11, 1 << 20, 100), ("19x19 1Mi", 19, 1 << 20, 40), ("11x11 131072", 11, 131072, 200)):
    env = HexBatch(N, G, variant=VARIANT_B, device=0, seed=0, agent_mode=AGENT_RANDOM)
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
    print(json.dumps({"build": tag, "config": name, "us_per_step_min": round(min(ts), 3), "us_per_step_med": round(sorted(ts)[2], 3)}), flush=True)
    env.close(); del env, g
