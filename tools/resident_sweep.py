"""Per-evaluation latency of fn_nll_grad: one launch per evaluation vs the resident kernel, over a
range of sizes (builder's measurement tool; prints one JSON line per case).

    python tools/resident_sweep.py [--cases 20000x6,125000x96,1000000x96] [--steps 2000]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import fitnoise_b200 as fitnoise  # noqa: E402
from fitnoise_b200 import _native  # noqa: E402


def one_case(e, n, m, model, steps, tune):
    rng = np.random.default_rng(1)
    design = np.stack([np.ones(m), (np.arange(m) >= m // 2) * 1.0], axis=1)
    y = rng.standard_normal((n, m))
    aux = None
    mod = getattr(fitnoise, model)()
    fam = mod._family(m)
    if fam.kind == _native.KIND_PATSEQ:
        aux = rng.poisson(20.0, size=(n, m)).astype(float)
        aux[rng.random((n, m)) < 0.15] = 0
        y = y * 10 + 50
        y[aux == 0] = np.nan
    P = fam.theta_length()
    theta = np.full(P, 1.1)
    if fam.dist == _native.DIST_T:
        theta[0] = 8.0
    if fam.kind == _native.KIND_PATSEQ:
        theta[-fam.nb:] = 0.02
        theta[P - fam.na - fam.nb:P - fam.nb] = 100.0
    e.set_tuning(**tune)
    e.upload(y, aux, fam, None, design)
    out = {"n": n, "m": m, "model": model, "tune": tune}
    for mode in ("launch", "resident"):
        if mode == "resident":
            e.resident_begin()
        for _ in range(50):
            e.nll_grad(theta)
        t0 = time.perf_counter()
        for _ in range(steps):
            r = e.nll_grad(theta)
        dt = (time.perf_counter() - t0) / steps
        if mode == "resident":
            e.resident_end()
        out[mode + "_us"] = round(dt * 1e6, 2)
        out[mode + "_value"] = r[0]
    out["kernel_queued_us"] = round(e.time_nll_grad(theta, 100) * 1e3, 2)
    bytes_ = n * (8 * m * (2 if aux is not None else 1) + 1)
    out["GBps_resident"] = round(bytes_ / (out["resident_us"] * 1e-6) / 1e9, 1)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="20000x6,20000x12,50000x24,125000x96,250000x96,500000x96,1000000x96")
    ap.add_argument("--model", default="Model_t")
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--threads", default="0")
    ap.add_argument("--lpg", default="0")
    ap.add_argument("--stages", default="0")
    args = ap.parse_args()
    e = _native.Engine(0)
    for case in args.cases.split(","):
        n, m = (int(x) for x in case.split("x"))
        for threads in (int(x) for x in args.threads.split(",")):
            for lpg in (int(x) for x in args.lpg.split(",")):
                for stages in (int(x) for x in args.stages.split(",")):
                    try:
                        print(json.dumps(one_case(e, n, m, args.model, args.steps,
                                                  dict(lpg=lpg, threads=threads, stages=stages))), flush=True)
                    except _native.NativeError as exc:
                        print(json.dumps({"n": n, "m": m, "error": str(exc)}), flush=True)
    e.close()


if __name__ == "__main__":
    main()
