"""Multi-process pinned host->device copy probe: N ranks (torchrun), each copying the bench's batch size from its own
pinned buffer to its own GPU at the same time, no kernels.  The ceiling of bench.py's e2e number at N GPUs.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/dev/h2d_probe_mp.py [--bind]
Prints one JSON line on rank 0: per-rank GB/s (min / median / max), aggregate GB/s, NUMA node of every GPU and of every
rank's CPU affinity.  --bind pins each rank to the CPUs NVML lists for its GPU before the pinned allocation."""
import json, os, sys, time
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
bind = "--bind" in sys.argv
stride = 1500 if "--packed" in sys.argv else 1536
cpus = None
if bind:
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1} & os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception as e:  # noqa: BLE001
        cpus = None
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
n = 1048576 * stride
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h.fill_(rank + 1)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(2):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 8
e0.record()
for _ in range(K):
    d.copy_(h, non_blocking=True)
e1.record()
torch.cuda.synchronize()
gbs = n * K / e0.elapsed_time(e1) / 1e6
def numa_of_gpu(i):
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(i)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        return int(open(f"/sys/bus/pci/devices/{bus.lower()[-12:]}/numa_node").read())
    except Exception:  # noqa: BLE001
        return None
info = {"rank": rank, "gbs": gbs, "gpu_numa": numa_of_gpu(local), "n_cpus_allowed": len(os.sched_getaffinity(0))}
if world > 1:
    out = [None] * world
    dist.all_gather_object(out, info)
else:
    out = [info]
if rank == 0:
    g = sorted(o["gbs"] for o in out)
    print(json.dumps({"n": world, "bind": bind, "bytes_per_copy": n, "per_rank_gbs": [round(o["gbs"], 1) for o in out],
                      "min": round(g[0], 1), "median": round(g[len(g) // 2], 1), "max": round(g[-1], 1), "aggregate_gbs": round(sum(g), 1),
                      "gpu_numa": [o["gpu_numa"] for o in out], "cpus_allowed": [o["n_cpus_allowed"] for o in out],
                      "host_cpus": os.cpu_count()}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
