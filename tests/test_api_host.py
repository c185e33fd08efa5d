"""Host-side logic of the package on a CPU-only box: the optimiser driver loop, result attributes,
lazy views, error behaviour.  The evaluator is the host build of the kernel mathematics
(tests/hostcheck); the CUDA path runs the same tests in test_gpu_parity.py."""
import json

import numpy as np
import pytest

import fitnoise_b200 as fitnoise
import hostcheck
from conftest import assert_close, golden_context, golden_fit_kwargs, load_golden
from helpers import fit_golden
from oracle import fitnoise_oracle as oracle

OPT_CASES = ["t_twogroup", "t_patseq_v2", "t_controls_opt", "t_factors1"]


@pytest.mark.parametrize("name", OPT_CASES)
def test_fitted_hyperparameters_match_oracle(name):
    """north_star: fitted hyper-parameters within 1e-6 relative at the same optimiser tolerance.
    The oracle runs the reference's driver loop (fitnoise.py:143-158) on its own score_row."""
    g = load_golden(name)
    fitted, tested = fit_golden(g, engine_factory=hostcheck.HostEvaluator)
    ref = oracle.fit_and_test(str(g["model"]), g["y"], g["design"], golden_context(g),
                              coef_idx=[int(i) for i in g["test_coef"]], nan_aware=True,
                              skip_rank_deficient=True, n_factors=int(g.get("n_factors", 0)),
                              **golden_fit_kwargs(g))
    x, rx = fitted.optimization_result.x, ref["x"]
    free = ~(np.isclose(rx, 100.0) | np.isclose(rx, 1e-3))          # df may sit on its bound
    assert_close(x[free], rx[free], 1e-6, what="theta")
    assert_close(x[~free], rx[~free], 1e-9, what="theta at bound")
    assert_close(fitted.optimization_result.fun, ref["fun"], 1e-10, what="fun")
    assert_close(fitted.deviance, ref["deviance"], 1e-10, what="deviance")
    assert_close(fitted.score, ref["score"], 1e-10, what="score")
    # and it is no worse than what the reference's own (loose, numeric-gradient) optimiser found
    # (comparable only when no row is excluded as rank deficient)
    from helpers import rank_deficient_rows
    if not rank_deficient_rows(g):
        assert fitted.optimization_result.fun <= float(g["ref_opt_fun"]) + 1e-9 * abs(float(g["ref_opt_fun"]))
    # statistics downstream of theta, against the oracle at ITS theta
    assert_close(tested.p_values, ref["p_values"], 1e-5, what="p at fitted theta")


def test_result_attributes_and_lazy_views():
    g = load_golden("t_twogroup")
    fitted, tested = fit_golden(g, engine_factory=hostcheck.HostEvaluator, force_param=g["theta"])
    for name in ("data", "noise_design", "control_design", "controls", "optimization_result", "param",
                 "df", "deviance", "score", "noise_dists", "noise_p_values", "noise_combined_p_value",
                 "averages", "design", "coef", "coef_dists", "description"):
        assert hasattr(fitted, name), name
    for name in ("contrasts", "contrast_dists", "p_values", "q_values"):
        assert hasattr(tested, name), name
    assert fitted.param.df == g["theta"][0] and fitted.param.variance == g["theta"][1]
    n, m = g["y"].shape
    assert len(fitted.noise_dists) == n and len(fitted.coef_dists) == n
    nd = fitted.noise_dists[3]
    assert isinstance(nd, fitnoise.Mvt) and nd.covar.shape == (m, m) and nd.df == g["theta"][0]
    cd = fitted.coef_dists[5]
    assert_close(cd.mean, g["ref_coef"][5], 1e-9)
    assert_close(cd.covar, g["ref_coef_covar"][5], 1e-9)
    td = tested.contrast_dists[5]
    assert_close(td.covar, g["ref_contrast_covar"][5], 1e-9)
    assert_close(td.p_value(np.zeros(1)), g["ref_p_values"][5], 1e-8)
    # the R bridge reads attributes through as_jsonic + json.dumps (R/fitnoise.R:75-88)
    for name in ("coef", "averages", "noise_p_values", "score", "df", "description"):
        json.dumps(fitnoise.as_jsonic(getattr(fitted, name)))
    json.dumps(fitnoise.as_jsonic(fitted.param))
    # fits are immutable records: test() returned a new object
    assert not hasattr(fitted, "p_values")


def test_status_rows_nan_pattern():
    """Which outputs are NaN for k>c / k==c / k<c / all-missing rows."""
    g = load_golden("independent_status")
    fitted, tested = fit_golden(g, engine_factory=hostcheck.HostEvaluator, force_param=g["theta"])
    assert np.array_equal(np.isnan(fitted.coef), np.isnan(g["ref_coef"]))
    assert np.array_equal(np.isnan(tested.p_values), np.isnan(g["ref_p_values"]))
    assert fitted.coef_dists[4] is None and fitted.coef_dists[0] is not None
    assert tested.q_values[4] == 1.0 or np.isnan(g["ref_p_values"][4])


def test_errors():
    y = np.random.default_rng(0).normal(size=(10, 6))
    design = np.array([[1.0, 0]] * 3 + [[1.0, 1]] * 3)
    m = fitnoise.Model_t()._with(_engine_factory=hostcheck.HostEvaluator)
    with pytest.raises(AssertionError):
        m.fit(y)                                                  # design is required (fitnoise.py:212)
    with pytest.raises(AssertionError):
        m.fit(y, design).test()                                   # "No coefficients or contrasts"
    with pytest.raises(AssertionError):
        m.fit(y, design, force_param=[1.0])                       # wrong length
    with pytest.raises(KeyError):
        fitnoise.Model_t_patseq()._with(_engine_factory=hostcheck.HostEvaluator).fit(y, design)
    counts = np.ones((10, 6))
    counts[0, 0] = 0
    with pytest.raises(AssertionError):                           # zero count with finite tail
        fitnoise.Model_t_patseq()._with(_engine_factory=hostcheck.HostEvaluator).fit(
            fitnoise.Dataset(y, {"counts": counts}), design)
    with pytest.raises(AssertionError):                           # factors need a SCALE1 variance model
        type("Bad", (fitnoise.Model_factors_mixin, fitnoise.Model_normal_per_sample), {})(1)._family(6)

    class Custom(fitnoise.Model):
        def _get_dist(self, param, aux):
            return None
    with pytest.raises(NotImplementedError):                      # no variance family -> no fallback
        Custom()._with(_engine_factory=hostcheck.HostEvaluator).fit(y, design)


def test_no_cpu_fallback_without_gpu(has_cuda):
    if has_cuda:
        pytest.skip("box has a GPU")
    from fitnoise_b200 import _native
    y = np.random.default_rng(0).normal(size=(10, 6))
    design = np.array([[1.0, 0]] * 3 + [[1.0, 1]] * 3)
    with pytest.raises(_native.NativeError):
        fitnoise.Model_t().fit(y, design)
    with pytest.raises(_native.NativeError):
        fitnoise.fdr([0.1, 0.2])


def test_nan_aware_initial_variance():
    g = load_golden("t_nan_quirk")
    fitted, _ = fit_golden(g, engine_factory=hostcheck.HostEvaluator)
    assert np.isfinite(fitted.optimization_result.x).all()
    assert np.isfinite(fitted.noise_p_values).sum() >= 11


def test_top_table():
    g = load_golden("t_per_sample_controls")
    fitted, tested = fit_golden(g, engine_factory=hostcheck.HostEvaluator, force_param=g["theta"])
    table = tested.top_table()
    order = np.argsort(g["ref_p_values"], kind="stable")
    assert list(table.columns) == ["coef2", "AveExpr", "P.Value", "adj.P.Val", "noise.P.Value", "control"]
    assert np.array_equal(np.asarray(table.index), order)
    assert_close(table["P.Value"].values, g["ref_p_values"][order], 1e-8)
    assert_close(table["adj.P.Val"].values, g["ref_q_values"][order], 1e-8)
    assert_close(table["coef2"].values, g["ref_contrasts"][order, 0], 1e-9)
    assert np.array_equal(table["control"].values, g["controls"][order])
    assert np.all(np.diff(table["P.Value"].values) >= 0)
    unsorted = tested.top_table(sort=False)
    assert np.array_equal(np.asarray(unsorted.index), np.arange(len(order)))
    # status rows: NaN p-values go last, in input order
    g = load_golden("independent_status")
    fitted, tested = fit_golden(g, engine_factory=hostcheck.HostEvaluator, force_param=g["theta"])
    table = tested.top_table()
    nan_rows = np.where(np.isnan(g["ref_p_values"]))[0]
    assert np.array_equal(np.asarray(table.index)[-len(nan_rows):], nan_rows)
    assert "control" not in table.columns


def test_top_table_of_the_noise_fit():
    """Legacy ``test.fit(fit)`` with neither coefs nor contrasts (backyard/old_fitnoise.R:776-795):
    the table of the noise p-values, BH-adjusted, ordered by them."""
    g = load_golden("t_per_sample_controls")
    fitted, _ = fit_golden(g, engine_factory=hostcheck.HostEvaluator, force_param=g["theta"])
    table = fitted.top_table()
    p = g["ref_noise_p_values"]
    key = np.where(np.isnan(p), np.inf, p)
    order = np.argsort(key, kind="stable")
    assert list(table.columns) == ["AveExpr", "P.Value", "adj.P.Val", "noise.P.Value", "control"]
    assert np.array_equal(np.asarray(table.index), order)
    assert_close(table["P.Value"].values, p[order], 1e-8)
    assert_close(table["adj.P.Val"].values, oracle.fdr(fitted.noise_p_values)[order], 1e-12)
    assert np.array_equal(table["P.Value"].values, table["noise.P.Value"].values, equal_nan=True)


def test_results_are_released_without_the_cycle_collector():
    """A fit's result arrays must die with the last reference to them (their pinned host blocks go
    back to the pool that the next fit draws from): no frame chain, closure or finaliser may keep
    them waiting for the garbage collector (regression: a lazy ``import torch.distributed`` inside
    Session.__init__ pinned the whole calling stack of a process's first fit)."""
    import gc
    import weakref
    g = load_golden("t_twogroup")
    gc.collect()
    gc.disable()
    try:
        fitted, tested = fit_golden(g, engine_factory=hostcheck.HostEvaluator, force_param=g["theta"])
        refs = [weakref.ref(o) for o in (fitted, tested, fitted._session, fitted.coef, tested.p_values,
                                         tested.q_values, fitted.noise_p_values)]
        fitted._session.close()
        del fitted, tested
        assert [r() is not None for r in refs] == [False] * len(refs)
    finally:
        gc.enable()


def test_cpython_fast_call_marshals_like_the_abi():
    """csrc/fn_pycall.c (the optional thin entry for fn_nll_grad) against a stand-in with the ABI's
    signature: theta and P go in, value / gradient / counts come out, the return code is passed on,
    and anything but contiguous float64 buffers is refused.  No GPU involved."""
    import ctypes
    from fitnoise_b200 import _native, build
    if build.build_pycall() is None:
        pytest.skip("no Python.h / C compiler here: the ctypes path is used instead")
    fast = _native._pycall()
    if fast is None:
        pytest.skip("FN_NO_PYCALL is set")
    c_double_p = ctypes.POINTER(ctypes.c_double)
    proto = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, c_double_p, ctypes.c_int32, c_double_p, c_double_p,
                             ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64))
    seen = {}

    def stand_in(handle, theta, P, value, grad, used, bad):
        seen.update(handle=handle, P=P, theta=[theta[i] for i in range(P)])
        value[0] = sum(seen["theta"])
        for i in range(P):
            grad[i] = 10.0 * theta[i]
        used[0], bad[0] = 2**40 + 7, 3
        return -5 if P == 1 else 0

    callback = proto(stand_in)
    address = ctypes.cast(callback, ctypes.c_void_p).value
    theta, grad = np.array([1.5, 2.5, 4.0]), np.empty(3)
    assert fast(address, 0xABCDEF, theta, grad) == (0, 8.0, 2**40 + 7, 3)
    assert seen == dict(handle=0xABCDEF, P=3, theta=[1.5, 2.5, 4.0]) and np.array_equal(grad, [15.0, 25.0, 40.0])
    assert fast(address, 1, np.array([2.0]), np.empty(1))[0] == -5                  # the library's error code comes through
    assert fast(address, 1, np.zeros(0), np.empty(0))[:2] == (0, 0.0)               # P = 0 (Model_independent)
    with pytest.raises(ValueError):
        fast(0, 1, theta, grad)
    for bad_theta, bad_grad in ((np.array([1, 2]), np.empty(2)), (np.ones(4)[::2], np.empty(2)),
                                (np.ones(3), np.empty(2)), (np.ones(2), np.empty(2, dtype=np.float32))):
        with pytest.raises((TypeError, BufferError, ValueError)):
            fast(address, 1, bad_theta, bad_grad)


def test_shard_threshold_setting(monkeypatch):
    """FN_SHARD_MIN_MB: matrices below it are fitted whole by every rank (session.py); default 8 MB,
    garbage falls back to the default, 0 shards everything."""
    from fitnoise_b200 import session
    monkeypatch.delenv("FN_SHARD_MIN_MB", raising=False)
    assert session._shard_min_bytes() == 8 * 1048576
    monkeypatch.setenv("FN_SHARD_MIN_MB", "0")
    assert session._shard_min_bytes() == 0
    monkeypatch.setenv("FN_SHARD_MIN_MB", "2.5")
    assert session._shard_min_bytes() == 2.5 * 1048576
    monkeypatch.setenv("FN_SHARD_MIN_MB", "many")
    assert session._shard_min_bytes() == 8 * 1048576
