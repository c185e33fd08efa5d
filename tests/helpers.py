"""Shared helpers of the test-suite: synthetic data of the BASELINE.json shapes and the
comparison of a fitted/tested model against the golden fixtures / the oracle."""
import numpy as np

import fitnoise_b200 as fitnoise
from conftest import assert_close, golden_context, golden_fit_kwargs, golden_test_kwargs

# Rows whose retained sub-design is rank deficient with k > c: the reference's REML term for them
# depends on LAPACK's Householder details; documented divergence -- such rows
# are excluded from the REML sum and get noise_p = NaN.
def rank_deficient_rows(g):
    y, design = g["y"], g.get("noise_design", g["design"])
    rows = []
    fin = np.isfinite(y)
    if "weights" in g:
        with np.errstate(divide="ignore"):
            fin &= np.isfinite(1.0 / g["weights"])
    for i in range(y.shape[0]):
        d = design
        if "controls" in g and g["controls"][i]:
            d = g["control_design"]
        k = fin[i].sum()
        if k > d.shape[1] and np.linalg.matrix_rank(d[fin[i]]) < d.shape[1]:
            rows.append(i)
    return rows


def check_against_golden(g, fitted, tested, rtol=1e-9, p_rtol=1e-8):
    skip = rank_deficient_rows(g)
    ref_np = g["ref_noise_p_values"].copy()
    got_np = fitted.noise_p_values.copy()
    if skip:
        assert np.all(np.isnan(got_np[skip]))
        ref_np[skip] = np.nan
    assert_close(got_np, ref_np, p_rtol, what="noise_p_values")
    assert_close(fitted.averages, g["ref_averages"], 1e-12, what="averages")
    if not skip:
        assert_close(fitted.noise_combined_p_value, g["ref_noise_combined_p_value"], p_rtol, what="combined p")
    assert fitted.df == g["ref_df"]
    assert_close(fitted.coef, g["ref_coef"], rtol, atol=1e-10, what="coef")
    assert_close(fitted._coef_covar, g["ref_coef_covar"], rtol, atol=1e-13, what="coef covar")
    assert_close(fitted._coef_df, g["ref_coef_df"], 1e-12, what="coef df")
    assert_close(tested.contrasts, g["ref_contrasts"], rtol, atol=1e-10, what="contrasts")
    assert_close(tested.p_values, g["ref_p_values"], p_rtol, what="p_values")
    assert_close(tested.q_values, g["ref_q_values"], p_rtol, what="q_values")
    assert_close(fitted.get_weights(), g["ref_weights"], 1e-12, what="weights")
    if not skip:
        assert fitted.description == str(g["ref_description"])


def fit_golden(g, engine_factory=None, **extra):
    nf = int(g.get("n_factors", 0))
    model = getattr(fitnoise, str(g["model"]))(nf) if nf else getattr(fitnoise, str(g["model"]))()
    if engine_factory is not None:
        model = model._with(_engine_factory=engine_factory)
    data = fitnoise.Dataset(g["y"], golden_context(g))
    fitted = model.fit(data, design=g["design"], **golden_fit_kwargs(g), **extra)
    tested = fitted.test(**golden_test_kwargs(g))
    return fitted, tested


def synthetic(config, n=None, seed=None):
    """Synthetic data of the five BASELINE.json configs (SURVEY 8d recipes), at n genes."""
    rng = np.random.default_rng({1: 2015, 2: 2016, 3: 2017, 4: 2018, 5: 2019}[config] if seed is None else seed)

    def t_noise(n, m, df):
        return rng.normal(size=(n, m)) * np.sqrt(df / rng.chisquare(df, size=(n, 1)))

    if config == 1:
        n = n or 20000; m = 6
        design = np.array([[1.0, 0.0]] * 3 + [[1.0, 1.0]] * 3)
        y = t_noise(n, m, 10.0)
        y[: n // 10, 3:] += np.where(rng.random(n // 10) < 0.5, 5.0, -5.0)[:, None]
        return dict(model="Model_t", y=y, design=design, context={}, test=dict(coef=[1]))
    if config == 2:
        n = n or 20000; m = 12
        g = np.arange(m) // 4
        design = np.stack([np.ones(m), (g == 1) * 1.0, (g == 2) * 1.0], axis=1)
        w = np.exp(rng.normal(0, 0.5, size=(n, m)))
        y = rng.normal(size=(n, m)) / np.sqrt(w)
        y[: n // 10] += 2.0 * design[:, 1]
        return dict(model="Model_independent", y=y, design=design, context={"weights": w},
                    test=dict(contrasts=np.array([[0.0, 0], [1, 0], [0, 1]])))
    if config == 3:
        n = n or 50000; m = 24
        design = np.array([[1.0, 0.0]] * 12 + [[1.0, 1.0]] * 12)
        counts = rng.poisson(np.exp(rng.normal(3, 1, size=(n, m)))).astype(float)
        counts[rng.random(size=(n, m)) < 0.15] = 0.0
        mu = rng.uniform(20, 100, size=(n, 1))
        var = 28.0 ** 2 / np.maximum(counts, 1) + (0.05 * mu) ** 2
        y = mu + rng.normal(size=(n, m)) * np.sqrt(var) * np.sqrt(12.6 / rng.chisquare(12.6, size=(n, 1)))
        y[: n // 20, 12:] += 15.0
        y[counts == 0] = np.nan
        return dict(model="Model_t_patseq", y=y, design=design, context={"counts": counts}, test=dict(coef=[1]))
    if config == 4:
        n = n or 200000; m = 48
        treat = (np.arange(m) >= m // 2) * 1.0
        batch = np.arange(m) % 4
        design = np.stack([np.ones(m), treat, (batch == 1) * 1.0, (batch == 2) * 1.0, (batch == 3) * 1.0,
                           rng.normal(size=m)], axis=1)
        control_design = np.delete(design, 1, axis=1)
        y = t_noise(n, m, 8.0) * 0.7 + 6.0
        y[: n // 10] += 2.0 * treat
        controls = rng.random(n) < 0.5
        controls[: n // 10] = False
        return dict(model="Model_t", y=y, design=design, context={}, test=dict(coef=[1]),
                    controls=controls, control_design=control_design)
    if config == 5:
        n = n or 1000000; m = 96
        design = np.array([[1.0, 0.0]] * 48 + [[1.0, 1.0]] * 48)
        y = t_noise(n, m, 10.0)
        y[: n // 10, 48:] += 1.0
        return dict(model="Model_independent", y=y, design=design, context={}, test=dict(coef=[1]))
    raise ValueError(config)
