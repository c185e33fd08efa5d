"""Multi-GPU check (run under torchrun, one rank per GPU, NCCL): the sharded fit through the public
API gives the single-GPU result.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/dist_check.py
"""
import os
import sys

os.environ.setdefault("FN_SHARD_MIN_MB", "0")      # shard even the small cases: this tool checks the sharded path

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import fitnoise_b200 as fitnoise  # noqa: E402
from helpers import synthetic  # noqa: E402

rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
ok = True
cases = [(1, 20000), (3, 5000), (4, 3001), ("factors", 4000)]


def make_case(config, n):
    if config != "factors":
        s = synthetic(config, n=n)
        return s, getattr(fitnoise, s["model"])()
    rng = np.random.default_rng(77)
    m = 8
    design = np.array([[1.0, 0.0]] * 4 + [[1.0, 1.0]] * 4)
    batch = rng.normal(size=m)
    y = rng.normal(size=(n, m)) * 0.7 + np.outer(rng.normal(size=n) * 1.5, batch)
    y[: n // 10, 4:] += 2.0
    return dict(model="Model_t_factors", y=y, design=design, context={}, test=dict(coef=[1])), fitnoise.Model_t_factors(1)


single, forced = {}, {}
for config, n in cases:                 # before the process group exists: plain single-GPU fits
    s, model = make_case(config, n)
    kw = {k: s[k] for k in ("controls", "control_design") if k in s}
    f = model.fit(fitnoise.Dataset(s["y"], s["context"]), s["design"], **kw)
    t = f.test(**s["test"])
    single[config] = (f.optimization_result.x.copy(), f.optimization_result.fun, t.p_values.copy(), t.q_values.copy(),
                      f.noise_p_values.copy())
    # and everything else at a FORCED theta, so that the sharded run can be compared to the last bit
    f2 = model.fit(fitnoise.Dataset(s["y"], s["context"]), s["design"], force_param=f.optimization_result.x, **kw)
    t2 = f2.test(**s["test"])
    forced[config] = dict(coef=f2.coef.copy(), covar=f2._coef_covar.copy(), dfpost=f2._coef_df.copy(),
                          contrasts=t2.contrasts.copy(), p=t2.p_values.copy(), q=t2.q_values.copy(),
                          order=t2._order.copy(), noise_p=f2.noise_p_values.copy(), averages=f2.averages.copy(),
                          weights=f2.get_weights().copy(), combined=f2.noise_combined_p_value)
    f._session.close()
    f2._session.close()

dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for config, n in cases:
    s, model = make_case(config, n)
    kw = {k: s[k] for k in ("controls", "control_design") if k in s}
    f = model.fit(fitnoise.Dataset(s["y"], s["context"]), s["design"], **kw)
    t = f.test(**s["test"])
    x1, fun1, p1, q1, np1 = single[config]
    dx = np.max(np.abs(f.optimization_result.x - x1) / np.abs(x1))
    dfun = abs(f.optimization_result.fun - fun1) / abs(fun1)
    dp = np.nanmax(np.abs(t.p_values - p1) / np.maximum(p1, 1e-300))
    dnp = np.nanmax(np.abs(f.noise_p_values - np1) / np.maximum(np1, 1e-300))
    same_nan = np.array_equal(np.isnan(t.p_values), np.isnan(p1))
    good = dx < 1e-6 and dfun < 1e-11 and dp < 1e-4 and same_nan
    ok &= bool(good)
    print("rank %d config %s world %d shard [%d,%d): dtheta %.2e dfun %.2e dp %.2e dnoise_p %.2e nan-pattern %s -> %s" % (
        rank, config, f._session.world, f._session.lo, f._session.hi, dx, dfun, dp, dnp, same_nan, "OK" if good else "FAIL"),
        flush=True)
    # forced theta: per-gene results do not depend on how the genes are sharded -> identical arrays
    f2 = model.fit(fitnoise.Dataset(s["y"], s["context"]), s["design"], force_param=x1, **kw)
    t2 = f2.test(**s["test"])
    got = dict(coef=f2.coef, covar=f2._coef_covar, dfpost=f2._coef_df, contrasts=t2.contrasts, p=t2.p_values,
               q=t2.q_values, order=t2._order, noise_p=f2.noise_p_values, averages=f2.averages,
               weights=f2.get_weights(), combined=f2.noise_combined_p_value)
    # (not bit-identical: a gene's position inside its tile sets the order in which its samples are
    # visited, and the shard boundaries move genes to other positions)
    def same(key):
        a, b = np.asarray(got[key]), np.asarray(forced[config][key])
        if key == "order":
            return np.array_equal(a, b)
        if key in ("coef", "contrasts", "covar"):
            # coefficients near zero have no relative accuracy to speak of: the bar of the parity tests
            # (1e-9 relative, 1e-10 absolute) on the scale of the array
            scale = float(np.nanmax(np.abs(b))) if b.size else 1.0
            return a.shape == b.shape and np.allclose(a, b, rtol=1e-9, atol=1e-10 * max(scale, 1e-300), equal_nan=True)
        return a.shape == b.shape and np.allclose(a, b, rtol=1e-11, atol=0.0, equal_nan=True)
    def worst(key):
        a, b = np.asarray(got[key], dtype=float), np.asarray(forced[config][key], dtype=float)
        return float(np.nanmax(np.abs(a - b))) if a.shape == b.shape and a.size else float("nan")
    bad = ["%s (max abs diff %.2e)" % (k, worst(k)) for k in got if not same(k)]
    ok &= not bad
    print("rank %d config %s forced theta: %s" % (rank, config, "all per-gene results equal to one GPU (1e-11; coefficients 1e-9 / 1e-10 of their scale), top-table order identical" if not bad
                                                   else "DIFFER: %s" % bad), flush=True)
    f._session.close()
    f2._session.close()
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
