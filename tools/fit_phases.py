"""Where a warm fit + test spends its wall time: the Engine calls timed one by one (host clock), the
rest is Python / scipy.      python tools/fit_phases.py <config> [model]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import fitnoise_b200 as fitnoise  # noqa: E402
from fitnoise_b200 import _native  # noqa: E402
from helpers import synthetic  # noqa: E402

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
spent = {}


def timed(name):
    real = getattr(_native.Engine, name)

    def wrapper(self, *a, **k):
        t0 = time.perf_counter()
        try:
            return real(self, *a, **k)
        finally:
            spent[name] = spent.get(name, 0.0) + time.perf_counter() - t0
            spent[name + "#"] = spent.get(name + "#", 0) + 1
    setattr(_native.Engine, name, wrapper)


for name in ("upload", "nll_grad", "resident_begin", "resident_end", "noise_stats", "coef", "test_bh", "coef_fetch"):
    if hasattr(_native.Engine, name):
        timed(name)


def run():
    fitted = getattr(fitnoise, model)().fit(data, s["design"], **kw)
    tested = fitted.test(**s["test"])
    fitted._session.close()
    return tested


for _ in range(3):
    run()
reps = 20
spent.clear()
t0 = time.perf_counter()
for _ in range(reps):
    run()
total = (time.perf_counter() - t0) / reps
print("config %d %s: warm fit + test %.3f ms" % (config, model, total * 1e3))
inside = 0.0
for name in sorted(k for k in spent if not k.endswith("#")):
    ms = spent[name] / reps * 1e3
    inside += ms
    print("  %-16s %8.3f ms  (%d calls)" % (name, ms, spent[name + "#"] // reps))
print("  %-16s %8.3f ms" % ("python / scipy", total * 1e3 - inside))
