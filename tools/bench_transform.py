"""Time the fitnoise.transform sums kernel and a whole transform fit (synthetic negative-binomial counts)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fitnoise_b200 as fitnoise
from fitnoise_b200 import _native


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genes", type=int, default=1000000)
    ap.add_argument("--samples", type=int, default=96)
    ap.add_argument("--order", type=int, default=2)
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    rng = np.random.default_rng(7)
    n, m = args.genes, args.samples
    design = np.column_stack([np.ones(m), np.arange(m) % 2]).astype(float)
    mean = np.exp(rng.normal(3.0, 1.5, size=(n, 1))) * np.exp(0.5 * design[:, 1])[None, :]
    x = rng.poisson(mean * rng.gamma(10.0, 0.1, size=(n, m))).astype(np.float64)
    engine = _native.acquire_engine(0)
    engine.transform_upload(x, design)
    a = np.concatenate([np.zeros((args.order - 1) * m), np.ones(m)])
    for _ in range(3):
        engine.transform_sums(args.order, a)
    t0 = time.perf_counter()
    for _ in range(args.reps):
        engine.transform_sums(args.order, a)
    per = (time.perf_counter() - t0) / args.reps
    _native.release_engine(engine)
    t0 = time.perf_counter()
    fit = fitnoise.transform(x, design, order=args.order)
    wall = time.perf_counter() - t0
    print(json.dumps({"genes": n, "samples": m, "order": args.order, "sums_ms": per * 1e3,
                      "sums_gbs": n * m * 8 / per / 1e9, "fit_s": wall,
                      "cost": float(fit.optimization_result.fun)}))


if __name__ == "__main__":
    main()
