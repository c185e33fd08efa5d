"""The oracle (oracle/fitnoise_oracle.py) against the reference's own outputs (tests/golden).

This is what pins the oracle: every fixture key ``ref_*`` was produced by the unmodified
reference (see tests/golden/make_golden.py).  CPU only.
"""
import numpy as np
import pytest

from conftest import (assert_close, golden_context, golden_fit_kwargs, golden_names,
                      load_golden)
from oracle import fitnoise_oracle as oracle

NAMES = golden_names()
# The oracle restates the same LAPACK/scipy calls; differences are rounding only.
RTOL = 1e-10


def run_oracle(g, **extra):
    kw = golden_fit_kwargs(g)
    if "test_coef" in g:
        kw["coef_idx"] = [int(i) for i in g["test_coef"]]
    else:
        kw["contrasts"] = g["test_contrasts"]
    kw.update(extra)
    return oracle.fit_and_test(str(g["model"]), g["y"], g["design"], golden_context(g),
                               n_factors=int(g.get("n_factors", 0)), **kw)


@pytest.mark.parametrize("name", NAMES)
def test_forced_param_outputs(name):
    g = load_golden(name)
    out = run_oracle(g, force_param=g["theta"])
    cfg = out["cfg"]
    assert_close(cfg.initial, g["ref_initial"], 1e-14, what="initial")
    bounds = np.asarray([[np.nan if v is None else v for v in b] for b in cfg.bounds], float).reshape(-1, 2)
    assert_close(bounds, g["ref_bounds"], 1e-14, what="bounds")
    if "ref_factors" in g:
        assert_close(np.asarray(cfg.factors(g["theta"])), g["ref_factors"], 1e-13, what="factors")
        assert_close(cfg.model_cost(g["theta"]), g["ref_model_cost"], 1e-12, atol=1e-300, what="model cost")

    n = g["y"].shape[0]
    scores = np.repeat(np.nan, n)
    for row, item in out["prepared"].items():
        scores[row] = oracle.score_row(cfg, g["theta"], row, *item)
    assert_close(scores, g["ref_score_rows"], RTOL, what="score_row")
    assert_close(out["noise_p_values"], g["ref_noise_p_values"], RTOL, what="noise_p")
    assert_close(out["averages"], g["ref_averages"], 1e-14, what="averages")
    assert_close(out["noise_combined_p_value"], g["ref_noise_combined_p_value"], RTOL, what="combined")
    assert out["df"] == g["ref_df"]
    assert out["fun"] == 0.0 and out["deviance"] == 0.0 and out["score"] == 0.0
    assert_close(out["coef"], g["ref_coef"], RTOL, atol=1e-12, what="coef")
    assert_close(out["coef_covar"], g["ref_coef_covar"], RTOL, atol=1e-14, what="coef_covar")
    assert_close(out["coef_df"], g["ref_coef_df"], 1e-14, what="coef_df")
    assert_close(out["contrasts"], g["ref_contrasts"], RTOL, atol=1e-12, what="contrasts")
    assert_close(out["contrast_covar"], g["ref_contrast_covar"], RTOL, atol=1e-14, what="contrast_covar")
    assert_close(out["p_values"], g["ref_p_values"], 1e-9, what="p_values")
    assert_close(out["q_values"], g["ref_q_values"], 1e-9, what="q_values")
    # BH of the reference's own p-values is bit-exact
    assert np.array_equal(oracle.fdr(g["ref_p_values"]), g["ref_q_values"])
    w = oracle.get_weights(cfg, g["theta"])
    assert_close(w, g["ref_weights"], 1e-14, what="weights")


@pytest.mark.parametrize("name", NAMES)
def test_gradient_matches_central_differences(name):
    g = load_golden(name)
    theta = g["theta"]
    if len(theta) == 0:
        pytest.skip("no hyper-parameters")
    out = run_oracle(g, force_param=theta)
    cfg, prepared = out["cfg"], out["prepared"]
    if not prepared:
        pytest.skip("reference skips every gene (NaN quirk)")
    value, grad = oracle.total_score_grad(cfg, prepared, theta)
    assert_close(value, oracle.total_score(cfg, prepared, theta), 1e-13, what="value")
    for i in range(len(theta)):
        h = 1e-5 * max(abs(theta[i]), 1e-3)
        tp, tm = theta.copy(), theta.copy()
        tp[i] += h
        tm[i] -= h
        num = (oracle.total_score(cfg, prepared, tp) - oracle.total_score(cfg, prepared, tm)) / (2 * h)
        assert abs(num - grad[i]) <= 2e-6 * max(abs(grad[i]), abs(value) / max(abs(theta[i]), 1e-3) * 1e-3, 1e-6), \
            "theta[%d]: analytic %r vs numeric %r" % (i, grad[i], num)


@pytest.mark.parametrize("name", [n for n in NAMES if "ref_opt_x" in load_golden(n)])
def test_optimiser_against_reference(name):
    """The reference's numpy optimiser (numeric gradient, default tolerance) stops ~1e-5 from the
    optimum; the tight driver loop (fitnoise.py:143-158 restated) must land within that and at a
    value no worse than the reference's."""
    g = load_golden(name)
    out = run_oracle(g)
    ref_x, ref_fun = g["ref_opt_x"], float(g["ref_opt_fun"])
    assert out["fun"] <= ref_fun + 1e-9 * abs(ref_fun)
    at_bound = np.isclose(ref_x, 100.0)     # df runs into its upper bound on some data sets
    assert np.all(np.abs(out["x"] - ref_x) <= 2e-3 * np.abs(ref_x) + 1e-12) or at_bound.any()
    assert abs(out["fun"] - ref_fun) <= 1e-7 * abs(ref_fun)
    # the loose restatement reproduces the reference optimiser itself
    loose = run_oracle(g, tight=False)
    # (numeric-gradient L-BFGS-B amplifies last-bit differences in the objective, hence 1e-3)
    assert_close(loose["x"], ref_x, 1e-3, what="loose x")
    assert_close(loose["fun"], ref_fun, 1e-8, what="loose fun")


def test_nan_quirk_documented():
    g = load_golden("t_nan_quirk")
    assert np.all(np.isnan(g["ref_score_rows"])) and np.all(np.isnan(g["ref_noise_p_values"]))
    assert np.all(np.isnan(g["ref_initial"][1:]))
    cfg = oracle.Configured("Model_t", g["y"], nan_aware=True)
    assert np.isfinite(cfg.initial).all()


def test_fdr_golden():
    g = load_golden("fdr")
    for i in "1234":
        q = oracle.fdr(g["p" + i])
        assert np.array_equal(q, g["ref_q" + i]), "fdr vector %s" % i
