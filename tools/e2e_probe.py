"""Distribution of the host-buffer step time per transport split (hexb_set_host_transport) on this box: 30 calls each,
min / median / p90 / max in ms. 1 Mi games of 11x11 per process; run under torchrun to see the contended case."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hex_gym_env_b200 import HexBatch, VARIANT_B, AGENT_RANDOM
rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import bench
    bench.pin_rank_cpus(local, world)
env = HexBatch(11, 1 << 20, variant=VARIANT_B, device=local, seed=1, game_offset=rank << 20, agent_mode=AGENT_RANDOM)
env.reset()
io = env.pinned_io()
for f in (1.0, 0.75, 0.5, 0.25, 0.0, -1.0):
    env.set_host_transport(f)
    for _ in range(24 if f < 0 else 3):
        env.step_host(None, io)
    if world > 1:
        dist.barrier()
    ts = []
    for _ in range(30):
        t0 = time.perf_counter()
        env.step_host(None, io)
        ts.append(1e3 * (time.perf_counter() - t0))
    ts.sort()
    if rank == 0:
        print(json.dumps({"world": world, "dma_fraction": f if f >= 0 else "adaptive -> %.2f" % env.host_transport(), "min": round(ts[0], 2),
                          "median": round(ts[15], 2), "p90": round(ts[27], 2), "max": round(ts[-1], 2), "threads": env._lib.hexb_host_threads()}), flush=True)
