"""Small all-modes run for compute-sanitizer (memcheck): every kernel mode on ragged batch sizes, odd and even boards."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hex_gym_env_b200 import HexBatch, VARIANT_A, VARIANT_B
for N, G, variant, am in ((11, 333, VARIANT_B, 2), (6, 129, VARIANT_B, 1), (7, 200, VARIANT_A, 0), (19, 65, VARIANT_B, 2), (3, 1, VARIANT_B, 0)):
    for auto in (True, False):
        env = HexBatch(N, G, variant=variant, device=0, seed=1, agent_mode=am, auto_reset=auto)
        env.reset()
        for t in range(N * N // 2 + 5):
            env.step(want_term=True, want_actions=True)
        env.step(env.sample_actions(np.full(G, 0.3)))
        env.step_host(None)
        env.encode(1); env.export_state(); env.stats()
        env.reset(reset_mask=(np.arange(G) % 2).astype(np.uint8))
    raw = HexBatch(N, G, variant=variant, device=0, raw=True)
    raw.reset()
    raw.import_boards(np.random.RandomState(0).choice([0, 1, 2], size=(G, N, N)).astype(np.int8), np.zeros(G, np.int8))
    raw.ply(np.zeros(G, np.int32)); raw.ply(np.full(G, N * N - 1, np.int32)); raw.export_state()
torch.cuda.synchronize()
print("sanitize_smoke ok")
