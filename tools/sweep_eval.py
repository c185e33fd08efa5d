"""Sweep the tile geometry of the nll+grad kernel (lanes per gene, threads, stages, CTAs/SM) on one
GPU and print the kernel time of each.  Development tool; not used by the product or the tests.

    python tools/sweep_eval.py [n] [m] [model]
"""
import itertools
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
weights = "weights" in sys.argv
counts = "patseq" in model
quick = "quick" in sys.argv

dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
with torch.cuda.stream(stream):
    y = torch.randn((n, m), dtype=torch.float64, device=dev)
    w = torch.exp(0.5 * torch.randn((n, m), dtype=torch.float64, device=dev)) if weights else None
    if counts:
        w = torch.poisson(torch.full((n, m), 20.0, dtype=torch.float64, device=dev))
        y = y * 5.0 + 50.0
        y[w == 0] = float("nan")
        weights = True
    design = np.array([[1.0, 0.0]] * (m // 2) + [[1.0, 1.0]] * (m - m // 2))
    fam = getattr(fitnoise, model)()._family(m)
    e = _native.Engine(0, stream.cuda_stream)
    e.upload_dev(n, m, y.data_ptr(), w.data_ptr() if weights else 0, fam, 0, design)
    e.set_timing(True)
    P = fam.theta_length()
    theta = np.array(([8.0] if fam.dist == 1 else []) + [1.1] * (P - (1 if fam.dist == 1 else 0)))
    if counts:
        theta = np.array(([8.0] if fam.dist == 1 else []) + [300.0] * fam.na + [0.004] * fam.nb)
    base = None
    rows = []
    arrays = 2 if weights else 1
    grid = itertools.product((128, 256), (1, 2, 4, 8), (1, 2, 3, 4), (1, 2, 3, 4))
    if quick:
        grid = [(0, 0, 0, 0), (256, 1, 2, 2), (256, 2, 2, 2), (256, 4, 2, 2), (256, 8, 2, 2), (128, 2, 2, 3), (128, 4, 3, 3), (256, 2, 3, 1), (256, 4, 3, 2)]
    for threads, lpg, stages, ctas in grid:
        if threads:
            genes = threads // lpg
            if genes % 16:
                continue
            stage = genes * m * 8 * arrays + genes
            if stage * stages * ctas > 215000 or stage * stages > 220000:
                continue
        try:
            e.set_tuning(lpg, threads, stages, ctas)
            for _ in range(2):
                v = e.nll_grad(theta)
            ts = []
            for _ in range(3):
                v = e.nll_grad(theta)
                ts.append(e.time_nll_grad(theta, 50))        # queued: kernel-only
        except Exception as exc:
            print(threads, lpg, stages, ctas, "ERR", exc)
            continue
        if base is None:
            base = v[0]
        assert abs(v[0] - base) <= 1e-9 * abs(base), (v[0], base)
        t = float(np.median(ts))
        gbs = n * (8 * m * arrays + 1) / (t * 1e-3) / 1e9
        rows.append((t, threads, lpg, stages, ctas, gbs))
        print("threads %3d lpg %2d stages %d ctas %d : %.4f ms  %.0f GB/s" % (threads, lpg, stages, ctas, t, gbs), flush=True)
    rows.sort()
    print("BEST:")
    for r in rows[:8]:
        print("  %.4f ms threads %d lpg %d stages %d ctas %d  %.0f GB/s" % r)
