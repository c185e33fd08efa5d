"""Minimal driver for ncu: upload a synthetic n x m matrix and run a few nll+grad evaluations.

    python tools/profile_eval.py [n] [m] [model] [evals] [design columns] [weights]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import fitnoise_b200 as fitnoise  # noqa: E402
from fitnoise_b200 import _native  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 96
model = sys.argv[3] if len(sys.argv) > 3 else "Model_t"
evals = int(sys.argv[4]) if len(sys.argv) > 4 else 6

dev = torch.device("cuda", 0)
y = torch.randn((n, m), dtype=torch.float64, device=dev)
torch.cuda.synchronize()
c = int(sys.argv[5]) if len(sys.argv) > 5 else 2
design = np.array([[1.0, 0.0]] * (m // 2) + [[1.0, 1.0]] * (m - m // 2))
if c > 2:
    design = np.column_stack([design, np.random.default_rng(1).normal(size=(m, c - 2))])
fam = getattr(fitnoise, model)()._family(m)
aux = None
if fam.kind == 2:                       # PAT-Seq: read counts; y = tail length, positive
    aux = torch.poisson(torch.full((n, m), 20.0, dtype=torch.float64, device=dev)) + 1.0
    y = y * 5.0 + 60.0
    torch.cuda.synchronize()
if len(sys.argv) > 6 and sys.argv[6] == "weights" and aux is None:      # precision weights (voom-like)
    aux = torch.exp(0.5 * torch.randn((n, m), dtype=torch.float64, device=dev))
    torch.cuda.synchronize()
e = _native.Engine(0)
e.upload_dev(n, m, y.data_ptr(), aux.data_ptr() if aux is not None else 0, fam, 0, design)
e.set_timing(True)
P = fam.theta_length()
theta = np.array(([8.0] if fam.dist == 1 else []) + [1.1] * (P - (1 if fam.dist == 1 else 0)))
if fam.kind == 2:
    theta = np.array(([8.0] if fam.dist == 1 else []) + [400.0] * fam.na + [0.005] * fam.nb)
for i in range(evals):
    v = e.nll_grad(theta)
    print("eval %d: value %.6f kernel %.4f ms" % (i, v[0], e.last_kernel_ms()))
