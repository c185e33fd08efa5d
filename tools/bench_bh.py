"""Benjamini-Hochberg on device-resident p-values: time per adjustment (CUDA graph replay) for the
one-sweep sort (default) and the classic five-launches-per-digit form (FN_BH_CLASSIC=1), checked
bit for bit against the oracle's numpy restatement of p_adjust.fdr.

    python tools/bench_bh.py [n ...]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from fitnoise_b200 import _native  # noqa: E402
from oracle import fitnoise_oracle as oracle  # noqa: E402

sizes = [int(a) for a in sys.argv[1:]] or [20000, 200000, 1000000, 4000000]
rng = np.random.default_rng(3)
for n in sizes:
    p = rng.random(n)
    p[: n // 10] = rng.random(n // 10) ** 8                 # a spike of small p-values
    p[rng.integers(0, n, n // 50)] = np.nan
    p[rng.integers(0, n, n // 50)] = 1.0                    # ties
    p[rng.integers(0, n, n // 100)] = p[rng.integers(0, n, n // 100)]
    want = oracle.fdr(p)
    order_want = np.argsort(np.where(np.isnan(p), np.inf, p), kind="stable")
    for classic in (False, True):
        os.environ.pop("FN_BH_CLASSIC", None)
        os.environ.pop("FN_BH_ONESWEEP", None)
        os.environ["FN_BH_CLASSIC" if classic else "FN_BH_ONESWEEP"] = "1"      # force either form at any size
        e = _native.Engine(0)
        pd = torch.from_numpy(p).cuda()
        qd = torch.empty_like(pd)
        torch.cuda.synchronize()
        optr = e.bh_ranked_dev(n, pd.data_ptr(), qd.data_ptr())
        e.synchronize()
        q = qd.cpu().numpy()
        order = np.empty(n, dtype=np.uint32)
        e.fetch(order, optr)
        ok = bool(np.array_equal(q, want) and np.array_equal(order.astype(np.int64), order_want))
        for _ in range(5):
            e.bh_ranked_dev(n, pd.data_ptr(), qd.data_ptr())
        e.synchronize()
        reps = 50
        t0 = time.perf_counter()
        for _ in range(reps):
            e.bh_ranked_dev(n, pd.data_ptr(), qd.data_ptr())
        e.synchronize()
        us = (time.perf_counter() - t0) / reps * 1e6
        print("n %8d  %-9s %8.1f us per adjustment   bit-exact q and order: %s" % (n, "classic" if classic else "one-sweep", us, ok), flush=True)
        e.close()
