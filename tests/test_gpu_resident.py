"""The resident evaluation kernel (fn_resident_begin / fn_resident_end) and the hand-over protocol.

The resident form must give the totals of the one-launch form for every variance family, survive
its own idle timeout, hand over millions of results without ever passing a stale or torn value,
and leave the GPU when any other entry point is called.  Needs a B200: ``-m gpu``.
"""
import os
import time

import numpy as np
import pytest

import fitnoise_b200 as fitnoise
from conftest import assert_close, load_golden
from fitnoise_b200 import _native

pytestmark = pytest.mark.gpu


def _case(model, n=3000, m=24, seed=5):
    rng = np.random.default_rng(seed)
    design = np.stack([np.ones(m), (np.arange(m) >= m // 2) * 1.0], axis=1)
    y = rng.normal(size=(n, m)) * np.sqrt(8.0 / rng.chisquare(8.0, size=(n, 1))) + 3.0
    aux = None
    mod = getattr(fitnoise, model[:-2] if model.endswith("_w") else model)()
    if "patseq" in model:
        aux = rng.poisson(15.0, size=(n, m)).astype(float)
        aux[rng.random((n, m)) < 0.15] = 0
        y = y * 10 + 40
        y[aux == 0] = np.nan
    elif model.endswith("_w"):
        aux = np.exp(rng.normal(size=(n, m)) * 0.5)
        aux[rng.random((n, m)) < 0.02] = 0.0
    else:
        y[rng.random((n, m)) < 0.03] = np.nan
    fam = mod._family(m)
    P = fam.theta_length()
    theta = np.empty(P)
    at = 0
    if fam.dist == _native.DIST_T:
        theta[at] = 7.5
        at += 1
    if fam.kind == _native.KIND_PATSEQ:
        theta[at:at + fam.na] = 120.0 * (1 + 0.1 * rng.random(fam.na))
        theta[at + fam.na:] = 0.02 * (1 + 0.1 * rng.random(fam.nb))
    else:
        theta[at:] = 1.3 * (1 + 0.2 * rng.random(P - at))
    return y, aux, fam, design, theta


MODELS = ["Model_t", "Model_normal", "Model_t_w", "Model_t_per_sample", "Model_normal_per_sample_w",
          "Model_t_patseq", "Model_t_patseq_v1", "Model_t_patseq_v3", "Model_t_patseq_per_sample",
          "Model_t_patseq_v1_per_sample"]


@pytest.mark.parametrize("model", MODELS)
def test_resident_matches_one_launch(model):
    y, aux, fam, design, theta = _case(model)
    e = _native.Engine(0)
    try:
        e.upload(y, aux, fam, None, design)
        thetas = [theta * (1.0 + 0.05 * k) for k in range(4)]
        base = [e.nll_grad(t) for t in thetas]
        e.resident_begin()
        assert e.resident_stats()[0]
        got = [e.nll_grad(t) for t in thetas] + [e.nll_grad(thetas[0])]
        on, launches, evals = e.resident_stats()
        e.resident_end()
        assert not e.resident_stats()[0]
        after = e.nll_grad(thetas[1])
    finally:
        e.close()
    assert evals >= 5 and launches >= 1
    for b, g in zip(base + [base[0], base[1]], got + [after]):
        assert b[2] == g[2] and b[3] == g[3] == 0
        assert_close(g[0], b[0], 1e-13, what="value")
        assert_close(g[1], b[1], 1e-10, atol=1e-13 * abs(b[0]), what="gradient")


def test_resident_survives_idle_timeout(monkeypatch):
    monkeypatch.setenv("FN_RESIDENT_IDLE_MS", "2")
    y, aux, fam, design, theta = _case("Model_t")
    e = _native.Engine(0)
    try:
        e.upload(y, aux, fam, None, design)
        base = e.nll_grad(theta)
        e.resident_begin()
        first = e.nll_grad(theta)
        for _ in range(3):
            time.sleep(0.05)                  # 25 idle timeouts: the kernel has left the GPU
            again = e.nll_grad(theta * 1.0)
            assert again[0] == first[0]
        _, launches, evals = e.resident_stats()
        e.resident_end()
    finally:
        e.close()
    assert launches >= 3 and evals == 4
    assert_close(first[0], base[0], 1e-13)


def test_other_entry_points_end_the_resident_phase():
    y, aux, fam, design, theta = _case("Model_t")
    e = _native.Engine(0)
    try:
        e.upload(y, aux, fam, None, design)
        e.resident_begin()
        v1 = e.nll_grad(theta)
        score, d2, p = e.per_gene(theta)          # must stop the resident kernel by itself, not hang
        assert not e.resident_stats()[0]
        coef, _, _ = e.coef(theta, design)
        e.resident_begin()
        v2 = e.nll_grad(theta)
        e.upload(y, aux, fam, None, design)       # re-upload while resident
        v3 = e.nll_grad(theta)
    finally:
        e.close()
    assert v1[0] == v2[0]
    assert_close(v3[0], v1[0], 1e-13)
    assert_close(np.nansum(score), v1[0], 1e-11)


def test_handover_litmus():
    """Hundreds of thousands of hand-overs with changing values: every result must be exactly the
    value that theta produces (a stale or torn {bits, tag} slot would pass on another theta's
    value, or a mix of two evaluations' accumulators)."""
    y, aux, fam, design, theta = _case("Model_t", n=64, m=6, seed=9)
    e = _native.Engine(0)
    reps = int(os.environ.get("FN_LITMUS_EVALS", 300000))
    try:
        e.upload(y, aux, fam, None, design)
        K = 97
        thetas = [np.array([3.0 + 0.37 * k, 0.5 + 0.11 * k]) for k in range(K)]
        e.resident_begin()
        expected = [e.nll_grad(t) for t in thetas]
        values = np.array([x[0] for x in expected])
        assert len(np.unique(values)) == K
        rng = np.random.default_rng(3)
        order = rng.integers(0, K, size=reps)
        bad = 0
        for k in order:
            v, g, used, nbad = e.nll_grad(thetas[k])
            ex = expected[k]
            if v != ex[0] or g[0] != ex[1][0] or g[1] != ex[1][1] or used != ex[2] or nbad != 0:
                bad += 1
        _, launches, evals = e.resident_stats()
        e.resident_end()
        # and the same through one launch per evaluation (the slots are shared by both forms)
        for k in order[:20000]:
            v, g, used, nbad = e.nll_grad(thetas[k])
            ex = expected[k]
            if v != ex[0] or g[0] != ex[1][0] or g[1] != ex[1][1] or used != ex[2]:
                bad += 1
    finally:
        e.close()
    assert bad == 0
    assert evals == K + reps


def test_fit_uses_the_resident_kernel():
    g = load_golden("t_twogroup")
    warm = fitnoise.Model_t().fit(g["y"], g["design"], force_param=g["theta"])     # the pooled engine
    _, launches0, evals0 = warm._session.local.resident_stats()
    warm._session.close()
    fitted = fitnoise.Model_t().fit(g["y"], g["design"])
    _, launches, evals = fitted._session.local.resident_stats()
    assert launches - launches0 >= 1 and evals - evals0 == fitted._evaluations
    os.environ["FN_NO_RESIDENT"] = "1"
    try:
        plain = fitnoise.Model_t().fit(g["y"], g["design"])
    finally:
        del os.environ["FN_NO_RESIDENT"]
    assert_close(fitted.optimization_result.x, plain.optimization_result.x, 1e-9)
    assert_close(fitted.optimization_result.fun, plain.optimization_result.fun, 1e-12)


def test_staged_upload_of_two_large_pageable_arrays():
    """y and weights of 64 MB each from pageable memory: the second staged copy must not reuse a pinned
    staging buffer whose DMA into y is still queued (fn_api.cu copy_h2d)."""
    rng = np.random.default_rng(21)
    n, m = 90000, 96
    y = rng.normal(size=(n, m))
    w = np.exp(rng.normal(size=(n, m)) * 0.3)
    design = np.stack([np.ones(m), (np.arange(m) >= m // 2) * 1.0], axis=1)
    fam = fitnoise.Model_t()._family(m)
    theta = np.array([6.0, 1.2])
    e = _native.Engine(0)
    try:
        e.upload(y, w, fam, None, design)
        big = e.per_gene(theta)[0]
        rows = np.r_[0:300, n // 2:n // 2 + 300, n - 300:n]
        e.upload(y[rows], w[rows], fam, None, design)
        small = e.per_gene(theta)[0]
    finally:
        e.close()
    # (the two uploads use different tile geometries, so the sums associate differently)
    assert_close(big[rows], small, 1e-12, what="per-gene score after two staged uploads")


@pytest.mark.parametrize("model", ["Model_t", "Model_t_per_sample", "Model_t_patseq"])
def test_cpython_fast_call_equals_ctypes(model):
    """fn_nll_grad through the thin CPython entry (csrc/fn_pycall.c, taken for float64 ndarrays) and
    through ctypes (taken for anything else, e.g. a list) is the same C-ABI call: identical bits."""
    if _native._pycall() is None:
        pytest.skip("the optional _fn_pycall module was not built")
    y, aux, fam, design, theta = _case(model, n=1500)
    e = _native.Engine(0)
    try:
        e.upload(y, aux, fam, None, design)
        for resident in (False, True):
            if resident:
                e.resident_begin()
            fast = e.nll_grad(np.ascontiguousarray(theta, dtype=np.float64))
            slow = e.nll_grad([float(t) for t in theta])
            if resident:
                e.resident_end()
            assert fast[0] == slow[0] and np.array_equal(fast[1], slow[1]) and fast[2:] == slow[2:]
            assert isinstance(fast[1], np.ndarray) and fast[1].shape == (len(theta),)
        with pytest.raises(_native.NativeError):
            e.nll_grad(np.zeros(len(theta) + 1))                # wrong length: the library's own error, raised
    finally:
        e.close()
