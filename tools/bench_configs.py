"""Throughput of the BASELINE configs that are not the bench.py headline (1 GPU): config 2 (7x7 variant-A HexEnv, 65,536 games,
launch-bound -> CUDA graph of 50 steps) and config 5 (19x19 SelfPlayEnv, 4,194,304 games). Prints one JSON line per config."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hex_gym_env_b200 import HexBatch, VARIANT_A, VARIANT_B, AGENT_RANDOM, AGENT_BLACK
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import contract_bytes, moved_bytes

def run(name, N, G, variant, agent_mode, K, graph):
    env = HexBatch(N, G, variant=variant, device=0, seed=0, agent_mode=agent_mode)
    env.reset()
    for _ in range(200):
        env.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if graph:
        g = torch.cuda.CUDAGraph(); s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            env.step(); torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=s):
                for _ in range(50):
                    env.step()
        g.replay(); torch.cuda.synchronize()
        e0.record()
        for _ in range(K // 50):
            g.replay()
        e1.record()
    else:
        e0.record()
        for _ in range(K):
            env.step()
        e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    st = dict(zip(("episodes", "black_wins", "white_wins", "agent_wins", "episode_plies", "invalid_ends", "env_steps", "plies"), env.stats().cpu().tolist()))
    print(json.dumps({"config": name, "board_size": N, "games": G, "steps": K, "cuda_graph": graph, "us_per_step": 1e3 * ms,
                      "env_steps_per_sec": G / (ms * 1e-3), "contract_GBps": G * contract_bytes(N) / (ms * 1e-3) / 1e9,
                      "moved_GBps": G * moved_bytes(N) / (ms * 1e-3) / 1e9, "stats": st}))

run("config2: 7x7 HexEnv (variant A) + random_policy opponent, 65,536 games", 7, 65536, VARIANT_A, AGENT_BLACK, 2000, True)
run("config2 without CUDA graph", 7, 65536, VARIANT_A, AGENT_BLACK, 2000, False)
run("config5: 19x19 SelfPlayEnv, 4,194,304 games", 19, 4194304, VARIANT_B, AGENT_RANDOM, 300, False)
run("config3-per-GPU-shard: 11x11 SelfPlayEnv, 131,072 games (L2 resident)", 11, 131072, VARIANT_B, AGENT_RANDOM, 2000, True)
run("6x6 SelfPlayEnv, 4,096 games (config 4 env side)", 6, 4096, VARIANT_B, AGENT_RANDOM, 2000, True)
