"""cProfile of a warm fit + test through the public API (host-side overheads).

    python tools/profile_fit.py <config> [model]
"""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import fitnoise_b200 as fitnoise  # noqa: E402
from helpers import synthetic  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
if world > 1:                                      # under torchrun: genes sharded over the ranks by the session
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))

config = int(sys.argv[1]) if len(sys.argv) > 1 else 1
s = synthetic(config)
model = sys.argv[2] if len(sys.argv) > 2 else s["model"]
kw = {k: s[k] for k in ("controls", "control_design") if k in s}
if os.environ.get("FN_PIN"):                       # inputs in pinned host memory (one DMA instead of the staged copy)
    import torch
    pinned = torch.empty(s["y"].shape, dtype=torch.float64, pin_memory=True)
    pinned.numpy()[:] = s["y"]
    s["y"] = pinned.numpy()
data = fitnoise.Dataset(s["y"], s["context"])


def run():
    fitted = getattr(fitnoise, model)().fit(data, s["design"], **kw)
    tested = fitted.test(**s["test"])
    fitted._session.close()
    return tested


run()
for _ in range(3):
    t0 = time.perf_counter()
    run()
    if rank == 0:
        print("warm fit+test: %.2f ms" % ((time.perf_counter() - t0) * 1000))
pr = cProfile.Profile()
pr.enable()
run()
pr.disable()
if rank == 0:
    pstats.Stats(pr).sort_stats("tottime").print_stats(22)
if world > 1:
    dist.destroy_process_group()
