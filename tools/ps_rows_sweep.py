"""Per-sample variance models: the thread-per-gene kernel on padded rows against the lanes-per-gene
kernel (FN_NO_ROWS=1), over tile geometries.  Checks value and gradient agreement, prints kernel times.

    python tools/ps_rows_sweep.py [n] [m] [c] [missing fraction] [controls] [model=Model_t]

(model=Model_t / Model_normal: the unweighted SCALE1 models, whose wide designs (c >= 4) take the same layout)
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import fitnoise_b200 as fitnoise  # noqa: E402
from fitnoise_b200 import _native  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 48
c = int(sys.argv[3]) if len(sys.argv) > 3 else 6
missing = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0

dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(5)
y = torch.randn((n, m), dtype=torch.float64, device=dev, generator=gen)
if missing > 0:
    y[torch.rand((n, m), device=dev, generator=gen) < missing] = float("nan")
torch.cuda.synchronize()
design = np.array([[1.0, 0.0]] * (m // 2) + [[1.0, 1.0]] * (m - m // 2))
if c > 2:
    design = np.column_stack([design, np.random.default_rng(1).normal(size=(m, c - 2))])
design = design[:, :c]
model = ([a.split("=", 1)[1] for a in sys.argv if a.startswith("model=")] or ["Model_t_per_sample"])[0]
fam = getattr(fitnoise, model)()._family(m)
P = fam.theta_length()
head = [8.0] if fam.dist == _native.DIST_T else []
theta = np.concatenate([head, 0.8 + 0.5 * np.random.default_rng(2).random(P - len(head))])

e = _native.Engine(0)
if "controls" in sys.argv:                      # half of the genes are control genes with their own design (cfg4)
    ctl = (torch.rand(n, device=dev, generator=gen) < 0.5).to(torch.uint8)
    torch.cuda.synchronize()
    e.upload_dev(n, m, y.data_ptr(), 0, fam, ctl.data_ptr(), design, np.delete(design, 1, axis=1))
else:
    e.upload_dev(n, m, y.data_ptr(), 0, fam, 0, design)
e.set_timing(True)


def run(label, env_no_rows, tune):
    if env_no_rows:
        os.environ["FN_NO_ROWS"] = "1"
    else:
        os.environ.pop("FN_NO_ROWS", None)
    try:
        e.set_tuning(*tune)
        for _ in range(2):
            v = e.nll_grad(theta)
        one = []
        for _ in range(5):
            v = e.nll_grad(theta)
            one.append(e.last_kernel_ms())
        queued = e.time_nll_grad(theta, 50)
    except Exception as exc:  # geometry rejected
        print(json.dumps({"label": label, "tune": tune, "error": str(exc)[:120]}), flush=True)
        return None
    print(json.dumps({"label": label, "tune": tune, "kernel_us": round(float(np.median(one)) * 1e3, 2),
                      "queued_us": round(queued * 1e3, 2), "value": v[0], "used": int(v[2]), "bad": int(v[3])}), flush=True)
    return v


base = run("lanes-per-gene (FN_NO_ROWS)", True, (0, 0, 0, 0))
ref = run("rows default", False, (0, 0, 0, 0))
if base is not None and ref is not None:
    scale = abs(base[0])
    dv = abs(base[0] - ref[0]) / scale
    dg = float(np.max(np.abs(base[1] - ref[1]) / (np.abs(base[1]) + 1e-9 * scale)))
    print(json.dumps({"check": "rows vs lanes", "rel_value": dv, "rel_grad": dg, "ok": bool(dv < 1e-11 and dg < 1e-8)}), flush=True)
for threads in (128, 256):
    for ctas in (1, 2, 3, 4):
        for stages in (1, 2):
            run("rows", False, (1, threads, stages, ctas))
