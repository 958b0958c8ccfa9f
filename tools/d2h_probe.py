"""Diagnostic: host<->device copy bandwidth per rank, alone and concurrently (why the host-buffer leg does not scale with GPUs).
torchrun --nproc-per-node N tools/d2h_probe.py"""
import os, sys, time
import torch, torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
aff0 = sorted(os.sched_getaffinity(0))
note = ""
try:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(local)
    pynvml.nvmlDeviceSetCpuAffinity(h)
    note = "pci %s" % pynvml.nvmlDeviceGetPciInfo(h).busId
    try:
        note += " numa %s" % open("/sys/bus/pci/devices/%s/numa_node" % pynvml.nvmlDeviceGetPciInfo(h).busId.lower()[4:]).read().strip()
    except Exception as e:
        note += " numa ?"
except Exception as e:
    note = "nvml: %s" % e
aff1 = sorted(os.sched_getaffinity(0))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 256 << 20
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
host = torch.empty(n, dtype=torch.uint8).pin_memory()
def bw(dst, src, reps=8):
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    return reps * n / (time.perf_counter() - t0) / 1e9
bw(host, dev, 2)
both = (bw(host, dev), bw(dev, host))
alone = None
for r in range(world):       # one rank at a time
    if world > 1: dist.barrier()
    if r == rank:
        alone = (bw.__wrapped__(host, dev) if hasattr(bw, "__wrapped__") else None)
print("rank %d: cpus before %d (%s..), after affinity %d (%d..%d); %s; concurrent D2H %.1f GB/s H2D %.1f GB/s" %
      (rank, len(aff0), aff0[:2], len(aff1), aff1[0], aff1[-1], note, both[0], both[1]), flush=True)
if rank == 0:
    os.system("nvidia-smi topo -m 2>/dev/null | head -20; lscpu | grep -i 'numa\\|socket\\|model name' | head; free -g | head -2")
if world > 1:
    dist.barrier(); dist.destroy_process_group()
