"""Fit + test of the five BASELINE.json configs at full size through the public API, with wall
times per stage and the nll+grad kernel time.  Development / reporting tool.

    python tools/bench_configs.py [config ...]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import fitnoise_b200 as fitnoise  # noqa: E402
from helpers import synthetic  # noqa: E402

configs = [int(a) for a in sys.argv[1:]] or [1, 2, 3, 4, 5]
models = {1: ["Model_t"], 2: ["Model_independent", "Model_t"], 3: ["Model_t_patseq"],
          4: ["Model_t", "Model_t_per_sample"], 5: ["Model_independent", "Model_t"]}
for config in configs:
    s = synthetic(config)
    n, m = s["y"].shape
    kw = {k: s[k] for k in ("controls", "control_design") if k in s}
    data = fitnoise.Dataset(s["y"], s["context"])
    for name in models[config]:
        for rep in range(4):                     # rep 0: cold (context, allocator, graph capture); reps 2-3: steady state
            t0 = time.perf_counter()
            fitted = getattr(fitnoise, name)().fit(data, s["design"], **kw)
            t1 = time.perf_counter()
            tested = fitted.test(**s["test"])
            t2 = time.perf_counter()
            eng = fitted._session.local
            theta = fitted._theta
            ms = None
            if len(theta):
                eng.set_timing(True)
                ts = []
                for _ in range(7):
                    eng.nll_grad(theta)
                    ts.append(eng.last_kernel_ms())
                ms = float(np.median(ts))
                t3 = time.perf_counter()
                for _ in range(20):
                    eng.nll_grad(theta)
                call_ms = (time.perf_counter() - t3) * 1000 / 20
            arrays = 1 + ("weights" in s["context"]) + ("counts" in s["context"])
            bytes_eval = n * (8 * m * arrays + (8 if "counts" in s["context"] else 0) + 1)
            row = dict(config=config, model=name, rep=rep, n=n, m=m, c=s["design"].shape[1],
                       fit_s=round(t1 - t0, 4), test_s=round(t2 - t1, 4), evaluations=fitted._evaluations,
                       theta=[float("%.6g" % v) for v in theta[:4]],
                       q_lt_0_01=int((tested.q_values < 0.01).sum()),
                       eval_kernel_ms=ms, eval_call_ms=round(call_ms, 4) if ms else None,
                       eval_GBps=round(bytes_eval / (ms * 1e-3) / 1e9, 1) if ms else None)
            print(json.dumps(row), flush=True)
            fitted._session.close()
