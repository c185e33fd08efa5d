"""The kernels' arithmetic (csrc/fn_gene.cuh, fn_special.cuh), compiled for the host by
tests/hostcheck, against scipy / mpmath and against the golden fixtures.  CPU only -- the same
checks run against the real CUDA kernels in test_gpu_parity.py."""
import numpy as np
import pytest
import scipy.special
import scipy.stats

import hostcheck
from conftest import golden_names, load_golden
from helpers import check_against_golden, fit_golden

NAMES = [n for n in golden_names() if n != "t_nan_quirk"]


def test_digamma():
    rng = np.random.default_rng(0)
    x = np.concatenate([10 ** rng.uniform(-4, 3, 20000), [5e-11, 0.5, 1, 2, 47.00000000005]])
    got, ref = hostcheck.digamma(x), scipy.special.digamma(x)
    assert np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-3)) < 1e-12


def test_lgamma_digamma_pair():
    """The combined routine behind the per-launch density tables, over every argument the tables can
    take: (v + p)/2 with v in [1e-10, 100] and p = 0 ... 2000."""
    import mpmath
    rng = np.random.default_rng(1)
    x = np.concatenate([10 ** rng.uniform(-11, 3.1, 40000), np.arange(1, 400) * 0.5, [5e-11, 1e-3, 9.999999, 10.0, 1000.5]])
    lg, dg = hostcheck.lgamma_digamma(x)
    ref_l, ref_d = scipy.special.gammaln(x), scipy.special.digamma(x)
    assert np.max(np.abs(lg - ref_l) / np.maximum(np.abs(ref_l), 1.0)) < 2e-14
    assert np.max(np.abs(dg - ref_d) / np.maximum(np.abs(ref_d), 1e-3)) < 1e-12
    for xv in [5e-11, 0.3, 1.5, 2.0, 9.5, 47.25, 180.0]:
        l, d = hostcheck.lgamma_digamma(np.array([xv]))
        assert abs(l[0] - float(mpmath.loggamma(xv))) <= 4e-15 * max(1.0, abs(float(mpmath.loggamma(xv))))
        assert abs(d[0] - float(mpmath.digamma(xv))) <= 1e-13 * max(1.0, abs(float(mpmath.digamma(xv))))


def test_f_sf_against_scipy():
    rng = np.random.default_rng(1)
    N = 100000
    dfn = rng.integers(1, 9, N).astype(float)
    dfd = np.where(rng.random(N) < 0.3, 1e-10 + rng.integers(0, 100, N),
                   10 ** rng.uniform(-3, 2.5, N) + rng.integers(0, 100, N))
    f = 10 ** rng.uniform(-8, 6, N)
    got, ref = hostcheck.f_sf(f, dfn, dfd), scipy.stats.f.sf(f, dfn, dfd)
    ok = ref > 1e-300
    assert np.max(np.abs(got - ref)[ok] / ref[ok]) < 1e-10          # north_star: p-values within 1e-8
    assert hostcheck.f_sf([0.0], [2.0], [5.0])[0] == 1.0
    assert hostcheck.f_sf([np.inf], [2.0], [5.0])[0] == 0.0
    assert np.isnan(hostcheck.f_sf([np.nan], [2.0], [5.0])[0])


def test_f_sf_deep_tail_against_mpmath():
    mpmath = pytest.importorskip("mpmath")
    mpmath.mp.dps = 60
    for f, dfn, dfd in [(1e4, 1, 12.5), (3e5, 2, 99.0), (250.0, 1, 94.0000000001), (1e9, 3, 7.25),
                        (5000.0, 1, 47.3), (80.0, 6, 1e-10 + 94)]:
        x = mpmath.mpf(dfd) / (mpmath.mpf(dfd) + mpmath.mpf(dfn) * mpmath.mpf(f))
        ref = mpmath.betainc(mpmath.mpf(dfd) / 2, mpmath.mpf(dfn) / 2, 0, x, regularized=True)
        got = hostcheck.f_sf([f], [dfn], [dfd])[0]
        assert abs(got - float(ref)) / float(ref) < 1e-10, (f, dfn, dfd, got, float(ref))


def test_chi2_sf_against_scipy():
    rng = np.random.default_rng(2)
    N = 100000
    q = 10 ** rng.uniform(-6, 3.2, N)
    k = rng.integers(1, 200, N).astype(float)
    got, ref = hostcheck.chi2_sf(q, k), scipy.stats.chi2.sf(q, k)
    ok = ref > 1e-300
    assert np.max(np.abs(got - ref)[ok] / ref[ok]) < 1e-10


@pytest.mark.parametrize("name", NAMES)
def test_kernel_math_matches_reference_golden(name):
    """Model.fit(force_param) + test through the package's host logic and the host build of the
    kernel mathematics, against the reference's outputs."""
    g = load_golden(name)
    fitted, tested = fit_golden(g, engine_factory=hostcheck.HostEvaluator, force_param=g["theta"])
    check_against_golden(g, fitted, tested)
    score, d2, p = fitted._session.per_gene(fitted._theta)
    ref = g["ref_score_rows"].copy()
    from helpers import rank_deficient_rows
    ref[rank_deficient_rows(g)] = np.nan
    from conftest import assert_close
    assert_close(score, ref, 1e-9, what="per-gene score_row")


@pytest.mark.parametrize("model", ["Model_t", "Model_t_per_sample", "Model_t_patseq"])
def test_blocked_sample_order(model):
    """m a multiple of 16 takes the kernels' blocked (bank-rotated) sample order."""
    from oracle import fitnoise_oracle as oracle
    from conftest import assert_close
    import fitnoise_b200 as fitnoise
    rng = np.random.default_rng(11)
    n, m = 40, 32
    design = np.stack([np.ones(m), (np.arange(m) >= 16) * 1.0, rng.normal(size=m)], axis=1)
    y = rng.normal(size=(n, m)) * 1.3 + 20.0
    context = {}
    if model == "Model_t_patseq":
        counts = rng.poisson(20.0, size=(n, m)).astype(float)
        counts[rng.random((n, m)) < 0.1] = 0
        y[counts == 0] = np.nan
        context = {"counts": counts}
        theta = [9.0, 300.0, 0.004]
    elif model == "Model_t_per_sample":
        theta = [7.0] + list(rng.uniform(0.5, 2.0, size=m))
    else:
        y[3, 5] = np.nan
        theta = [7.0, 1.5]
    fitted = getattr(fitnoise, model)()._with(_engine_factory=hostcheck.HostEvaluator).fit(
        fitnoise.Dataset(y, context), design, force_param=theta)
    ref = oracle.fit_and_test(model, y, design, context, coef_idx=[1], force_param=theta, nan_aware=True)
    score, d2, p = fitted._session.per_gene(fitted._theta)
    want = np.full(n, np.nan)
    for row, item in ref["prepared"].items():
        want[row] = oracle.score_row(ref["cfg"], np.asarray(theta), row, *item)
    assert_close(score, want, 1e-9, what="score")
    value, grad, used, bad = fitted._objective(theta)
    ov, og = oracle.total_score_grad(ref["cfg"], ref["prepared"], np.asarray(theta))
    assert_close(value, ov, 1e-11)
    assert_close(grad, og, 1e-8, atol=1e-9 * abs(ov))
    assert_close(fitted.coef, ref["coef"], 1e-9, atol=1e-10)


@pytest.mark.parametrize("c", [9, 12, 16])
def test_wide_designs(c):
    """Designs with 9 ... 16 columns run in the template widths 12 and 16."""
    from oracle import fitnoise_oracle as oracle
    from conftest import assert_close
    import fitnoise_b200 as fitnoise
    rng = np.random.default_rng(100 + c)
    n, m = 25, 40
    design = np.column_stack([np.ones(m), (np.arange(m) % 2) * 1.0, rng.normal(size=(m, c - 2))])
    y = rng.normal(size=(n, m)) * 1.2 + 8.0
    weights = np.exp(rng.normal(0, 0.3, size=(n, m)))
    weights[rng.random((n, m)) < 0.05] = 0.0
    theta = [6.0, 1.4]
    contrasts = np.zeros((c, 3))
    contrasts[1, 0] = 1.0; contrasts[c - 1, 1] = 1.0; contrasts[2, 2] = 1.0; contrasts[c - 2, 2] = -1.0
    fitted = fitnoise.Model_t()._with(_engine_factory=hostcheck.HostEvaluator).fit(
        fitnoise.Dataset(y, {"weights": weights}), design, force_param=theta)
    tested = fitted.test(contrasts=contrasts)
    ref = oracle.fit_and_test("Model_t", y, design, {"weights": weights}, contrasts=contrasts, force_param=theta)
    score, d2, p = fitted._session.per_gene(fitted._theta)
    want = np.full(n, np.nan)
    for row, item in ref["prepared"].items():
        want[row] = oracle.score_row(ref["cfg"], np.asarray(theta), row, *item)
    assert_close(score, want, 1e-9, what="score")
    value, grad, used, bad = fitted._objective(theta)
    ov, og = oracle.total_score_grad(ref["cfg"], ref["prepared"], np.asarray(theta))
    assert_close(value, ov, 1e-11)
    assert_close(grad, og, 1e-8, atol=1e-9 * abs(ov))
    assert_close(fitted.coef, ref["coef"], 1e-8, atol=1e-9)
    assert_close(tested.p_values, ref["p_values"], 1e-8)
    assert_close(tested.q_values, ref["q_values"], 1e-8)


@pytest.mark.parametrize("name", NAMES)
def test_objective_and_gradient_match_oracle(name):
    """Packed-theta objective (incl. the factor penalty and the chain rule through Q2) against the
    oracle's matrix-calculus gradient on the reference's own covariance."""
    from conftest import assert_close, golden_context, golden_fit_kwargs
    from helpers import rank_deficient_rows
    from oracle import fitnoise_oracle as oracle
    g = load_golden(name)
    theta = g["theta"]
    if len(theta) == 0:
        pytest.skip("no hyper-parameters")
    fitted, _ = fit_golden(g, engine_factory=hostcheck.HostEvaluator, force_param=theta)
    value, grad, used, bad = fitted._objective(theta)
    kw = golden_fit_kwargs(g)
    nd = g.get("noise_design", g["design"])
    cfg = oracle.Configured(str(g["model"]), g["y"], golden_context(g), n_factors=int(g.get("n_factors", 0)),
                            design=nd, control_design=kw.get("control_design", np.ones((g["y"].shape[1], 1))),
                            controls=kw.get("controls"))
    out = oracle.fit_noise(cfg, nd, kw.get("control_design"), kw.get("controls"), force_param=theta,
                           skip_rank_deficient=True)
    ov, og = oracle.total_score_grad(cfg, out["prepared"], theta)
    assert used == len(out["prepared"]) and bad == 0
    assert_close(value, ov, 1e-11, what="objective")
    assert_close(grad, og, 1e-8, atol=1e-9 * abs(ov), what="gradient")
