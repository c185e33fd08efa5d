"""Host-to-device bandwidth per rank, alone and with all ranks copying at once (torchrun, one rank per
GPU): separates a shared-uplink / host-memory limit from anything the library does.  Rank 0 also
prints the box's topology.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/h2d_probe.py
"""
import os
import subprocess
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import bind_to_gpu_numa_node  # noqa: E402

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
node = bind_to_gpu_numa_node(local) if not os.environ.get("FN_NO_BIND") else None
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
if rank == 0:
    for cmd in (["nvidia-smi", "topo", "-m"], ["lscpu"], ["numactl", "--hardware"]):
        try:
            out = subprocess.run(cmd, capture_output=True, text=True, timeout=20).stdout
            print("$ %s\n%s" % (" ".join(cmd), "\n".join(out.splitlines()[:40])), flush=True)
        except Exception as exc:
            print("$ %s: %s" % (" ".join(cmd), exc), flush=True)
nbytes = 256 << 20
host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
host.fill_(1)
dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")


def copy_rate(reps=6):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    return reps * nbytes / (time.perf_counter() - t0) / 1e9


copy_rate(2)
alone = None
for r in range(world):
    dist.barrier()
    if r == rank:
        alone = copy_rate()
    dist.barrier()
dist.barrier()
together = copy_rate()
dist.barrier()
pairs = None
if world >= 4:                       # every second rank: do neighbours share an uplink?
    dist.barrier()
    pairs = copy_rate() if rank % 2 == 0 else None
    dist.barrier()
rows = [None] * world
dist.all_gather_object(rows, (rank, node, sorted(os.sched_getaffinity(0))[:4], alone, together, pairs))
if rank == 0:
    for r, nd, cpus, a, t, p in rows:
        print("rank %d numa %s cpus %s...: alone %.1f GB/s, all %d together %.1f GB/s, even ranks only %s" % (
            r, nd, cpus, a, world, t, "%.1f GB/s" % p if p else "-"), flush=True)
dist.destroy_process_group()
