"""Randomised parity of the kernels' arithmetic with the oracle (the numpy restatement of the
reference, pinned by tests/golden): many small random shapes, missing patterns, weights, counts,
control genes and every model family -- far more sub-designs than the golden fixtures hold.  CPU only
(tests/hostcheck is the same per-gene code the CUDA kernels compile; test_gpu_parity.py and
tools/fuzz_parity.py tie the kernels to it on the GPU)."""
import zlib

import numpy as np
import pytest

import fitnoise_b200 as fitnoise
import hostcheck
from conftest import assert_close
from oracle import fitnoise_oracle as oracle

MODELS = ["Model_t", "Model_normal", "Model_independent", "Model_t_per_sample", "Model_normal_per_sample",
          "Model_t_patseq", "Model_t_v1_patseq", "Model_normal_patseq_v3", "Model_t_patseq_per_sample",
          "Model_t_patseq_v1_per_sample"]


def _case(rng, model):
    m = int(rng.integers(4, 15))
    c = int(rng.integers(1, min(5, m - 2) + 1))
    n = int(rng.integers(3, 14))
    design = np.column_stack([np.ones(m)] + [rng.normal(size=m) for _ in range(c - 1)])
    y = rng.normal(size=(n, m)) * rng.uniform(0.5, 2.0) + rng.uniform(5, 60)
    context = {}
    patseq = "patseq" in model
    if patseq:
        counts = rng.poisson(rng.uniform(3, 30), size=(n, m)).astype(float)
        counts[rng.random((n, m)) < 0.15] = 0
        y = np.abs(y) + 1.0
        y[counts == 0] = np.nan
        context = {"counts": counts}
    else:
        y[rng.random((n, m)) < 0.1] = np.nan
        if rng.random() < 0.6 or model == "Model_independent":
            w = np.exp(rng.normal(0, 0.5, size=(n, m)))
            w[rng.random((n, m)) < 0.05] = 0.0
            context = {"weights": w}
    kwargs = {}
    if c > 1 and rng.random() < 0.4:
        kwargs = dict(controls=rng.random(n) < 0.5, control_design=design[:, :-1])
    return dict(model=model, y=y, design=design, context=context, kwargs=kwargs, n=n, m=m, c=c)


def _theta(rng, cfg):
    lo = np.array([b[0] for b in cfg.bounds], dtype=float)
    hi = np.array([b[1] for b in cfg.bounds], dtype=float)
    init = np.asarray(cfg.initial, dtype=float)
    theta = init * np.exp(rng.normal(0, 0.3, size=len(init)))
    return np.minimum(np.maximum(theta, lo * 1.001), hi * 0.999) if len(init) else theta


@pytest.mark.parametrize("model", MODELS)
def test_random_cases_against_oracle(model):
    rng = np.random.default_rng(zlib.crc32(model.encode()))
    checked = 0
    for trial in range(12):
        case = _case(rng, model)
        y, design, context, kwargs = case["y"], case["design"], case["context"], case["kwargs"]
        cfg0 = oracle.Configured(model, y, context, nan_aware=True, design=design,
                                 control_design=kwargs.get("control_design", np.ones((case["m"], 1))),
                                 controls=kwargs.get("controls"))
        theta = _theta(rng, cfg0)
        what = "%s trial %d n=%d m=%d c=%d" % (model, trial, case["n"], case["m"], case["c"])
        fitted = getattr(fitnoise, model)()._with(_engine_factory=hostcheck.HostEvaluator).fit(
            fitnoise.Dataset(y, context), design, force_param=theta, **kwargs)
        tested = fitted.test(coef=[case["c"] - 1])
        ref = oracle.fit_and_test(model, y, design, context, coef_idx=[case["c"] - 1], force_param=theta,
                                  nan_aware=True, skip_rank_deficient=True, **kwargs)
        score, d2, p = fitted._session.per_gene(fitted._theta)
        want = np.full(case["n"], np.nan)
        for row, item in ref["prepared"].items():
            want[row] = oracle.score_row(ref["cfg"], np.asarray(theta), row, *item)
        assert_close(score, want, 1e-9, what=what + " score")
        if len(theta) and len(ref["prepared"]):
            value, grad, used, bad = fitted._objective(theta)
            ov, og = oracle.total_score_grad(ref["cfg"], ref["prepared"], np.asarray(theta))
            assert used == len(ref["prepared"]) and bad == 0, what
            assert_close(value, ov, 1e-10, what=what + " objective")
            assert_close(grad, og, 1e-7, atol=1e-9 * max(abs(ov), 1.0), what=what + " gradient")
        # rows whose retained sub-design is nearly collinear amplify rounding: compare where the
        # oracle's own coefficients are moderate
        ok = np.all(np.abs(np.nan_to_num(ref["coef"])) < 1e4, axis=1)
        assert np.array_equal(np.isnan(fitted.coef), np.isnan(ref["coef"])), what + " coef NaN pattern"
        assert_close(fitted.coef[ok], ref["coef"][ok], 1e-7, atol=1e-8, what=what + " coef")
        assert_close(tested.p_values[ok], ref["p_values"][ok], 1e-6, atol=1e-300, what=what + " p")
        assert_close(fitted.noise_p_values, ref["noise_p_values"], 1e-7, atol=1e-300, what=what + " noise p")
        assert_close(fitted.averages, ref["averages"], 1e-12, atol=1e-13, what=what + " averages")
        assert_close(fitted.get_weights(), oracle.get_weights(ref["cfg"], np.asarray(theta)), 1e-12, what=what + " weights")
        checked += 1
    assert checked == 12


@pytest.mark.parametrize("model", ["Model_t_factors", "Model_normal_factors", "Model_independent_factors"])
def test_random_factor_cases_against_oracle(model):
    """Unknown-factor models: the augmented ridge system of the kernels against the oracle's dense
    covariance diag(sigma2) + F F' (plus_covar), at a perturbed starting point."""
    rng = np.random.default_rng(zlib.crc32(model.encode()))
    for trial in range(6):
        nf = int(rng.integers(1, 3))
        m = int(rng.integers(8, 15))
        c = int(rng.integers(1, 4))
        n = int(rng.integers(6, 14))
        design = np.column_stack([np.ones(m)] + [rng.normal(size=m) for _ in range(c - 1)])
        F = rng.normal(size=(m, nf))
        y = rng.normal(size=(n, m)) + rng.normal(size=(n, nf)) @ F.T + 10.0
        y[rng.random((n, m)) < 0.05] = np.nan
        context = {}
        if rng.random() < 0.5 or model == "Model_independent_factors":
            context = {"weights": np.exp(rng.normal(0, 0.4, size=(n, m)))}
        cfg0 = oracle.Configured(model, y, context, nan_aware=True, n_factors=nf, design=design,
                                 control_design=np.ones((m, 1)), controls=None)
        init = np.asarray(cfg0.initial, dtype=float)
        theta = init + rng.normal(0, 0.05, size=len(init)) * np.maximum(np.abs(init), 0.1)
        for i, (lo, hi) in enumerate(cfg0.bounds):          # keep strictly inside the finite bounds
            if lo is not None:
                theta[i] = max(theta[i], lo * 1.001)
            if hi is not None:
                theta[i] = min(theta[i], hi * 0.999)
        what = "%s trial %d nf=%d n=%d m=%d c=%d" % (model, trial, nf, n, m, c)
        fitted = getattr(fitnoise, model)(nf)._with(_engine_factory=hostcheck.HostEvaluator).fit(
            fitnoise.Dataset(y, context), design, force_param=theta)
        ref = oracle.fit_and_test(model, y, design, context, coef_idx=[c - 1], force_param=theta,
                                  nan_aware=True, skip_rank_deficient=True, n_factors=nf)
        value, grad, used, bad = fitted._objective(theta)
        ov, og = oracle.total_score_grad(ref["cfg"], ref["prepared"], np.asarray(theta))
        assert used == len(ref["prepared"]) and bad == 0, what
        assert_close(value, ov, 1e-10, what=what + " objective")
        assert_close(grad, og, 1e-6, atol=1e-8 * max(abs(ov), 1.0), what=what + " gradient")
        assert_close(fitted.coef, ref["coef"], 1e-7, atol=1e-8, what=what + " coef")
        assert_close(fitted.noise_p_values, ref["noise_p_values"], 1e-7, atol=1e-300, what=what + " noise p")
