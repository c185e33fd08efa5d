"""Wall time of K1 (fn_upload_dev: prepare kernel) and K3 (fn_coef with no host outputs) per geometry."""
import ctypes, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import fitnoise_b200 as fitnoise
from fitnoise_b200 import _native

n, m = int(sys.argv[1]), int(sys.argv[2])
weights = "weights" in sys.argv
dev = torch.device("cuda", 0)
y = torch.randn((n, m), dtype=torch.float64, device=dev)
w = torch.exp(0.5 * torch.randn((n, m), dtype=torch.float64, device=dev)) if weights else None
torch.cuda.synchronize()
design = np.array([[1.0, 0.0]] * (m // 2) + [[1.0, 1.0]] * (m - m // 2))
fam = fitnoise.Model_t()._family(m)
theta = np.array([8.0, 1.1])
for tuning in [None, dict(lpg=2, threads=256, stages=1, ctas_per_sm=2), dict(lpg=4, threads=256, stages=1, ctas_per_sm=2),
               dict(lpg=4, threads=256, stages=2, ctas_per_sm=2), dict(lpg=8, threads=256, stages=2, ctas_per_sm=2)]:
    e = _native.Engine(0)
    try:
        if tuning:
            e.set_tuning(**tuning)
        ts_up, ts_coef = [], []
        for rep in range(6):
            t0 = time.perf_counter()
            e.upload_dev(n, m, y.data_ptr(), w.data_ptr() if weights else 0, fam, 0, design)
            ts_up.append(time.perf_counter() - t0)
            t0 = time.perf_counter()
            e._check(e._lib.fn_coef(e._h, _native._dp(theta), 2, 2, _native._dp(_native._f64(design)), None, None, None))
            ts_coef.append(time.perf_counter() - t0)
        print("%-60s prepare %.3f ms  coef %.3f ms" % (tuning, min(ts_up) * 1e3, min(ts_coef) * 1e3), flush=True)
    except Exception as exc:
        print(tuning, "ERR", exc)
    finally:
        e.close()
