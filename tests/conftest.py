import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))
                  if not p.endswith("fdr.npz") and not p.endswith("transform_poly.npz")
                  and not os.path.basename(p).startswith("full_"))


def load_golden(name):
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def golden_context(g):
    ctx = {}
    if "weights" in g:
        ctx["weights"] = g["weights"]
    if "counts" in g:
        ctx["counts"] = g["counts"]
    return ctx


def golden_fit_kwargs(g):
    kw = {}
    for key in ("controls", "control_design", "noise_design"):
        if key in g:
            kw[key] = g[key]
    return kw


def golden_test_kwargs(g):
    if "test_coef" in g:
        return dict(coef=[int(i) for i in g["test_coef"]])
    return dict(contrasts=g["test_contrasts"])


def assert_close(a, b, rtol, atol=0.0, what=""):
    a = np.asarray(a, dtype="float64")
    b = np.asarray(b, dtype="float64")
    assert a.shape == b.shape, "%s: shape %s vs %s" % (what, a.shape, b.shape)
    nan_a, nan_b = np.isnan(a), np.isnan(b)
    assert np.array_equal(nan_a, nan_b), "%s: NaN pattern differs at %s" % (
        what, np.argwhere(nan_a != nan_b)[:5].tolist())
    ok = ~nan_a
    inf = np.isinf(a) | np.isinf(b)
    assert np.array_equal(a[inf], b[inf]), "%s: infinities differ" % what
    ok &= ~inf
    if not ok.any():
        return
    err = np.abs(a[ok] - b[ok])
    tol = atol + rtol * np.abs(b[ok])
    worst = np.argmax(err - tol)
    assert np.all(err <= tol), "%s: max violation |%r - %r| = %.3e (rtol %.1e)" % (
        what, a[ok][worst], b[ok][worst], err[worst], rtol)


@pytest.fixture(scope="session")
def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
