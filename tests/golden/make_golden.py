"""Generate the golden fixtures ``tests/golden/*.npz`` from the UNMODIFIED reference.

Run in the build container (needs ``/root/reference``):

    python tests/golden/make_golden.py

The reference ships no tests or golden vectors, so parity is pinned by
running the reference's own numpy path (``use_theano=False``; py2->py3 syntax patched in
memory by ``refimport.py``) on small seeded inputs and storing inputs + outputs.  Every
number stored under a key starting with ``ref_`` was produced by reference code:

  * ``ref_score_rows``  per-gene ``score_row`` (fitnoise.py:44-50) at theta = ``theta``,
                        NaN for genes the REML prep skips (fitnoise.py:28)
  * ``ref_noise_p_values, ref_averages, ref_noise_combined_p_value, ref_df``
                        Model.fit_noise(force_param=theta) (fitnoise.py:207-291)
  * ``ref_coef, ref_coef_covar, ref_coef_df``  Model.fit_coef (fitnoise.py:294-347)
  * ``ref_contrasts, ref_contrast_covar, ref_p_values, ref_q_values``  Model.test (:368-396)
  * ``ref_weights``     Model.get_weights (:399-404)
  * ``ref_description`` repr(fitted)  (:185-196)
  * ``ref_opt_x, ref_opt_fun``  (some cases) the reference's own optimiser result,
                        scipy L-BFGS-B with numeric gradient and default tolerances (:73-76)
  * ``ref_initial, ref_bounds``  Model._configured output

Inputs are stored under ``y, design, weights, counts, controls, control_design, theta,
test_coef | test_contrasts`` and ``model`` (reference class name).
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refimport  # noqa: E402

warnings.simplefilter("ignore")


def two_group(m):
    h = m // 2
    return np.array([[1.0, 0.0]] * h + [[1.0, 1.0]] * (m - h))


def three_group(m):
    g = np.arange(m) % 3
    return np.stack([np.ones(m), (g == 1) * 1.0, (g == 2) * 1.0], axis=1)


def t_noise(rng, n, m, df):
    return rng.normal(size=(n, m)) * np.sqrt(df / rng.chisquare(df, size=(n, 1)))


ONLY = None          # set by `--only name,name`: write just these fixtures (the others stay as committed)


def run_case(ref, name, model, y, design, theta, context=None, test_coef=None,
             test_contrasts=None, controls=None, control_design=None, noise_design=None,
             optimise=False, n_factors=0):
    """`theta` may be a function of the reference's configured initial vector (factor models: the
    pack coordinates depend on the reference's own Q2)."""
    if ONLY is not None and name not in ONLY:
        return
    context = context or {}
    base_cls = getattr(ref, model)
    cls = (lambda: base_cls(n_factors)) if n_factors else base_cls
    data = ref.Dataset(y, context)
    n, m = y.shape
    kwargs = dict(design=design, use_theano=False)
    if noise_design is not None:
        kwargs["noise_design"] = noise_design
    if controls is not None:
        kwargs["controls"] = list(controls)
    if control_design is not None:
        kwargs["control_design"] = control_design

    if callable(theta):
        nd0 = design if noise_design is None else noise_design
        cd0 = np.ones((m, 1)) if control_design is None else control_design
        fl0 = [False] * n if controls is None else list(controls)
        conf0 = cls()._with(data=data, noise_design=nd0, control_design=cd0, controls=np.asarray(fl0, bool),
                            _aux=np.zeros((n, 0)))._configured()
        theta = theta(np.asarray(conf0._initial, float))
    fitted = cls().fit(data, force_param=theta, **kwargs)
    out = dict(model=model, y=y, design=design, theta=np.asarray(theta, float), n_factors=np.int64(n_factors))
    for key, value in context.items():
        out[key] = np.asarray(value, float)
    if controls is not None:
        out["controls"] = np.asarray(controls, bool)
    if control_design is not None:
        out["control_design"] = control_design
    if noise_design is not None:
        out["noise_design"] = noise_design

    # per-gene score_row through the reference's own objects (fitnoise.py:23-50)
    nd = design if noise_design is None else noise_design
    cd = np.ones((m, 1)) if control_design is None else control_design
    flags = [False] * n if controls is None else list(controls)
    configured = cls()._with(data=data, noise_design=nd, control_design=cd,
                             controls=np.asarray(flags, bool),
                             _aux=np.zeros((n, 0)))._configured()
    get_dist = lambda param, aux: configured._get_dist(configured._unpack(param), aux)
    designs = [cd if f else nd for f in flags]
    _, prepared = ref.fitnoise._fit_noise(
        y=data.y, designs=designs, get_model_cost=lambda p: 0.0, get_dist=get_dist,
        initial=ref.env.as_vector(configured._initial), aux=configured._aux,
        bounds=configured._bounds, use_theano=False, verbose=False,
        force_param=np.asarray(theta, float))
    scores = np.repeat(np.nan, n)
    th = ref.env.as_vector(theta)
    for row, (retain, tQ2, z2) in prepared.items():
        scores[row] = -(get_dist(th, configured._aux[row]).marginal(retain)
                        .transformed(tQ2).log_density(z2))
    out["ref_score_rows"] = scores
    out["ref_initial"] = np.asarray(configured._initial, float)
    out["ref_bounds"] = np.asarray([[np.nan if v is None else v for v in b] for b in configured._bounds],
                                   float).reshape(-1, 2)
    if n_factors:
        out["ref_factor_Q2"] = configured._factor_Q2
        out["ref_factors"] = np.asarray(fitted.param.factors, float)
        out["ref_model_cost"] = np.float64(configured._model_cost(configured._unpack(th)))

    out["ref_noise_p_values"] = fitted.noise_p_values
    out["ref_averages"] = fitted.averages
    out["ref_noise_combined_p_value"] = np.float64(fitted.noise_combined_p_value)
    out["ref_df"] = np.float64(fitted.df)
    out["ref_coef"] = fitted.coef
    c = design.shape[1]
    cov = np.tile(np.nan, (n, c, c))
    cdf = np.tile(np.nan, n)
    for i, dist in enumerate(fitted.coef_dists):
        if dist is not None:
            cov[i] = dist.covar
            if hasattr(dist, "df"):
                cdf[i] = dist.df
    out["ref_coef_covar"] = cov
    out["ref_coef_df"] = cdf
    out["ref_weights"] = fitted.get_weights()
    out["ref_description"] = np.array(repr(fitted))

    if test_coef is not None:
        tested = fitted.test(coef=list(test_coef))
        out["test_coef"] = np.asarray(test_coef, int)
    else:
        tested = fitted.test(contrasts=test_contrasts)
        out["test_contrasts"] = np.asarray(test_contrasts, float)
    out["ref_contrasts"] = tested.contrasts
    kc = tested.contrasts.shape[1]
    ccov = np.tile(np.nan, (n, kc, kc))
    for i, dist in enumerate(tested.contrast_dists):
        if dist is not None:
            ccov[i] = dist.covar
    out["ref_contrast_covar"] = ccov
    out["ref_p_values"] = tested.p_values
    out["ref_q_values"] = tested.q_values

    if optimise:
        opt = cls().fit(data, **kwargs)
        out["ref_opt_x"] = np.asarray(opt.optimization_result.x, float)
        out["ref_opt_fun"] = np.float64(opt.optimization_result.fun)
        out["ref_opt_score"] = np.float64(opt.score)
        out["ref_opt_deviance"] = np.float64(opt.deviance)
        out["ref_opt_description"] = np.array(repr(opt))

    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print("%-34s n=%d m=%d finite scores=%d  p finite=%d  %.1f KB" % (
        name, n, m, np.isfinite(scores).sum(), np.isfinite(tested.p_values).sum(),
        os.path.getsize(path) / 1024.0))


def punch_status_rows(y, m, c_groups):
    """Rows of every status in SURVEY 8a's quirk table (two-group design, first half group 1)."""
    h = m // 2
    y[1, h:] = np.nan                 # k > c, a whole group missing: rank-deficient sub-design
    y[2, :] = np.nan
    y[2, 0] = 0.3; y[2, h] = -0.2     # k == c
    y[3, 1:] = np.nan                 # k < c
    y[4, :] = np.nan                  # all missing
    y[5, 0] = np.nan                  # one missing
    y[6, [0, h]] = np.nan             # one missing per group
    y[7, :h] = np.nan                 # the other whole group missing
    y[8, :] = np.nan
    y[8, [0, 1, h]] = [1.0, 1.5, 0.2]  # k == c + 1
    return y


def main():
    ref = refimport.load()
    rng = np.random.default_rng(20151019)

    # 1. Model_t, two groups, complete data; the reference optimiser too.
    n, m = 120, 6
    y = t_noise(rng, n, m, 10.0)
    y[:12, 3:] += 5.0
    run_case(ref, "t_twogroup", "Model_t", y, two_group(m), [7.3, 0.8], test_coef=[1], optimise=True)

    # 2. Model_normal with precision weights (one zero weight, one NaN weight), 3 groups, F-test.
    n, m = 90, 9
    w = np.exp(rng.normal(0, 0.5, size=(n, m)))
    y = rng.normal(size=(n, m)) / np.sqrt(w)
    y[:9, 1::3] += 3.0
    w[3, 4] = 0.0
    w[5, 0] = 0.0; w[5, 8] = 0.0
    w[7, 2] = np.nan
    run_case(ref, "normal_weights", "Model_normal", y, three_group(m), [1.3],
             context={"weights": w}, test_contrasts=np.array([[0, 0], [1, 0], [0, 1.0]]))

    # 3. Model_independent with weights (voom style), contrast F-test: cfg2 in miniature.
    n, m = 100, 12
    w = np.exp(rng.normal(0, 0.5, size=(n, m)))
    y = rng.normal(size=(n, m)) / np.sqrt(w)
    y[:10, 2::3] -= 2.5
    run_case(ref, "independent_weights", "Model_independent", y, three_group(m), [],
             context={"weights": w}, test_contrasts=np.array([[0, 0], [1, 0], [-1, 1.0]]))

    # 4. Model_independent with every row status of the quirk table.
    n, m = 40, 6
    y = punch_status_rows(rng.normal(size=(n, m)), m, 2)
    run_case(ref, "independent_status", "Model_independent", y, two_group(m), [], test_coef=[1])

    # 5. Model_t_patseq (= v2), 15 % missing, cfg3 in miniature; with the reference optimiser.
    def patseq_data(n, m, frac=0.15):
        counts = rng.poisson(np.exp(rng.normal(3, 1, size=(n, m)))).astype(float)
        counts[rng.random(size=(n, m)) < frac] = 0.0
        mu = rng.uniform(20, 100, size=(n, 1))
        var = 28.0 ** 2 / np.maximum(counts, 1) + (0.05 * mu) ** 2
        y = mu + rng.normal(size=(n, m)) * np.sqrt(var) * np.sqrt(12.6 / rng.chisquare(12.6, size=(n, 1)))
        y[:, m // 2:] += np.where(np.arange(n)[:, None] < n // 20, 15.0, 0.0)
        y[counts == 0] = np.nan
        return y, counts
    n, m = 150, 10
    y, counts = patseq_data(n, m)
    y[3, :] = np.nan; counts[3, :] = 0
    y[4, 2:] = np.nan; counts[4, 2:] = 0
    y[5, m // 2:] = np.nan; counts[5, m // 2:] = 0
    run_case(ref, "t_patseq_v2", "Model_t_patseq", y, two_group(m), [12.6, 700.0, 0.003],
             context={"counts": counts}, test_coef=[1], optimise=True)

    # 6. The other PAT-Seq families at fixed theta.
    n, m = 60, 8
    y, counts = patseq_data(n, m)
    run_case(ref, "t_patseq_v1", "Model_t_v1_patseq", y, two_group(m), [9.0, 600.0, 0.004],
             context={"counts": counts}, test_coef=[1])
    y, counts = patseq_data(n, m)
    run_case(ref, "normal_patseq_v3", "Model_normal_patseq_v3", y, two_group(m), [0.2, 0.006],
             context={"counts": counts}, test_coef=[1])
    y, counts = patseq_data(n, m)
    th = [11.0] + list(rng.uniform(400, 900, size=m)) + list(rng.uniform(0.002, 0.006, size=m))
    run_case(ref, "t_patseq_v2_per_sample", "Model_t_patseq_per_sample", y, two_group(m), th,
             context={"counts": counts}, test_coef=[1])
    y, counts = patseq_data(n, m)
    th = [8.0, 650.0] + list(rng.uniform(0.002, 0.006, size=m))
    run_case(ref, "t_patseq_v1_per_sample", "Model_t_patseq_v1_per_sample", y, two_group(m), th,
             context={"counts": counts}, test_coef=[1])
    y, counts = patseq_data(n, m)
    run_case(ref, "normal_patseq_v2", "Model_normal_patseq", y, two_group(m), [800.0, 0.002],
             context={"counts": counts}, test_coef=[1])

    # 7. Model_t_per_sample with control genes and a control design (cfg4 in miniature),
    #    design != noise_design.
    n, m = 80, 8
    treat = (np.arange(m) >= m // 2) * 1.0
    batch = (np.arange(m) % 2) * 1.0
    covar = rng.normal(size=m)
    design = np.stack([np.ones(m), treat, batch, covar], axis=1)
    control_design = np.stack([np.ones(m), batch, covar], axis=1)
    sds = rng.uniform(0.6, 1.6, size=m)
    y = t_noise(rng, n, m, 6.0) * sds
    y[:8] += 3.0 * treat
    controls = np.arange(n) % 2 == 1
    th = [6.0] + list(sds ** 2)
    run_case(ref, "t_per_sample_controls", "Model_t_per_sample", y, design, th,
             test_coef=[1], controls=controls, control_design=control_design)
    # Model_t on the same data, with the optimiser, controls on.
    run_case(ref, "t_controls_opt", "Model_t", y, design, [6.0, 1.1],
             test_coef=[1], controls=controls, control_design=control_design, optimise=True)
    # Model_normal_per_sample with weights; noise design differs from the coefficient design.
    w = np.exp(rng.normal(0, 0.4, size=(n, m)))
    run_case(ref, "normal_per_sample_weights", "Model_normal_per_sample", y / np.sqrt(w), design,
             list(sds ** 2 * 1.1), context={"weights": w},
             test_contrasts=np.array([[0.0], [1.0], [-1.0], [0.0]]),
             noise_design=control_design)

    # 8. Model_t at m=24, c=6: a wider design.
    n, m = 50, 24
    X = np.stack([np.ones(m), (np.arange(m) >= 12) * 1.0, (np.arange(m) % 4 == 1) * 1.0,
                  (np.arange(m) % 4 == 2) * 1.0, (np.arange(m) % 4 == 3) * 1.0, rng.normal(size=m)], axis=1)
    y = t_noise(rng, n, m, 4.0) * 1.4 + 7.0
    y[:5] += 2.0 * X[:, 1]
    w = np.ones((n, m))
    w[9, 3] = 0.0; w[10, 20:] = 0.0; w[11, ::2] = 0.0
    run_case(ref, "t_wide_design", "Model_t", y, X, [4.0, 2.0], context={"weights": w},
             test_coef=[1, 5])

    # 8b. The reference quirk: NaN in y with Model_t makes the initial variance NaN
    #     (fitnoise.py:546), so `good` is all-False at the initial theta and the REML prep skips
    #     every gene.  Stored to document the behaviour the product deliberately diverges from.
    n, m = 12, 6
    y = rng.normal(size=(n, m))
    y[2, 1] = np.nan
    run_case(ref, "t_nan_quirk", "Model_t", y, two_group(m), [5.0, 1.0], test_coef=[1])

    # 10. Unknown batch-effect factors (Model_factors_mixin, fitnoise.py:439-518).
    def factor_data(n, m, nf, sd=0.7):
        load = rng.normal(size=(n, nf)) * 1.5
        batch = rng.normal(size=(nf, m))
        return rng.normal(size=(n, m)) * sd + load @ batch

    perturb = lambda init: init * (1.0 + 0.1 * np.cos(np.arange(len(init)) + 1.0))
    n, m = 80, 6
    y = factor_data(n, m, 1)
    y[:8, 3:] += 3.0
    run_case(ref, "t_factors1", "Model_t_factors", y, two_group(m), perturb, test_coef=[1],
             n_factors=1, optimise=True)
    n, m = 60, 8
    w = np.exp(rng.normal(0, 0.4, size=(n, m)))
    y = factor_data(n, m, 2) / np.sqrt(w)
    w[4, 1] = 0.0
    run_case(ref, "normal_factors2_weights", "Model_normal_factors", y, two_group(m), perturb,
             context={"weights": w}, test_coef=[1], n_factors=2)
    n, m = 60, 6
    y = factor_data(n, m, 1)
    y[3, 0] = np.nan; y[7, 4:] = np.nan; y[9, :] = np.nan; y[11, 3:] = np.nan
    run_case(ref, "independent_factors1_nan", "Model_independent_factors", y, two_group(m), perturb,
             test_coef=[1], n_factors=1)
    n, m = 70, 8
    treat = (np.arange(m) >= m // 2) * 1.0
    design = np.stack([np.ones(m), treat, (np.arange(m) % 2) * 1.0], axis=1)
    control_design = np.stack([np.ones(m), (np.arange(m) % 2) * 1.0], axis=1)
    y = factor_data(n, m, 1)
    y[:7] += 2.0 * treat
    controls = np.arange(n) % 3 == 0
    run_case(ref, "t_factors1_controls", "Model_t_factors", y, design, perturb, test_coef=[1],
             controls=controls, control_design=control_design, n_factors=1)

    # 9. fdr vectors (p_adjust.py:3-21).
    p1 = rng.random(500); p1[::17] = np.nan; p1[5] = p1[6] = p1[7]; p1[20] = 0.0; p1[21] = 1.0
    p2 = rng.random(37) ** 6
    p3 = np.array([np.nan, np.nan])
    p4 = np.array([0.5])
    if ONLY is None or "fdr" in ONLY:
        np.savez_compressed(os.path.join(HERE, "fdr.npz"),
                            p1=p1, ref_q1=ref.fdr(p1), p2=p2, ref_q2=ref.fdr(p2),
                            p3=p3, ref_q3=ref.fdr(p3), p4=p4, ref_q4=ref.fdr(p4))
        print("fdr")

    # 11. Designs with 9 ... 13 columns (the kernels' template widths 12 and 16).  Own generator, so
    #     that the fixtures above keep their random draws.
    rng2 = np.random.default_rng(20261020)

    def wide(m, c):
        cols = [np.ones(m), (np.arange(m) >= m // 2) * 1.0]
        cols += [(np.arange(m) % 5 == k) * 1.0 for k in range(1, 4)]
        cols += [rng2.normal(size=m) for _ in range(c - len(cols))]
        return np.stack(cols[:c], axis=1)

    n, m, c = 40, 30, 10
    X = wide(m, c)
    y = t_noise(rng2, n, m, 6.0) * 1.2 + 5.0
    y[:6] += 2.5 * X[:, 1]
    w = np.exp(rng2.normal(0, 0.4, size=(n, m)))
    w[3, 4] = 0.0; w[8, 20:] = 0.0
    contrasts = np.zeros((c, 2)); contrasts[1, 0] = 1.0; contrasts[c - 1, 1] = 1.0; contrasts[2, 1] = -1.0
    run_case(ref, "t_design_c10", "Model_t", y, X, [6.0, 1.3], context={"weights": w}, test_contrasts=contrasts)

    n, m, c = 36, 40, 13
    X = wide(m, c)
    y = rng2.normal(size=(n, m)) * 0.9 + 3.0
    y[:5] -= 1.5 * X[:, 1]
    y[4, 7] = np.nan; y[6, ::3] = np.nan; y[9, 5:] = np.nan; y[12, :] = np.nan
    w = np.exp(rng2.normal(0, 0.3, size=(n, m)))
    run_case(ref, "independent_design_c13", "Model_independent", y, X, [], context={"weights": w},
             test_coef=[1, 12])

    n, m, c = 40, 32, 9
    X = wide(m, c)
    counts = rng2.poisson(18.0, size=(n, m)).astype(float)
    counts[rng2.random((n, m)) < 0.12] = 0
    mu = rng2.uniform(30, 90, size=(n, 1))
    y = mu + 4.0 * X[:, 1][None, :] * (np.arange(n) < 5)[:, None] \
        + t_noise(rng2, n, m, 9.0) * np.sqrt(26.0 ** 2 / np.maximum(counts, 1) + (0.05 * mu) ** 2)
    y[counts == 0] = np.nan
    run_case(ref, "t_patseq_design_c9", "Model_t_patseq", y, X, [11.0, 650.0, 0.0028],
             context={"counts": counts}, test_coef=[1])

    # 12. Many genes: 1500 distinct missing patterns per family, straight from the reference.
    n, m = 1500, 12
    y = rng2.normal(size=(n, m)) * 1.1 + 6.0
    y[:150, 6:] += rng2.normal(0, 2.0, size=(150, 1))
    y[rng2.random((n, m)) < 0.15] = np.nan
    w = np.exp(rng2.normal(0, 0.5, size=(n, m)))
    w[rng2.random((n, m)) < 0.03] = 0.0
    run_case(ref, "independent_many_rows", "Model_independent", y, two_group(m), [], context={"weights": w},
             test_coef=[1])
    counts = rng2.poisson(8.0, size=(n, m)).astype(float)
    counts[rng2.random((n, m)) < 0.15] = 0
    mu = rng2.uniform(30, 90, size=(n, 1))
    y = mu + 5.0 * two_group(m)[:, 1][None, :] * (np.arange(n) < 100)[:, None] \
        + t_noise(rng2, n, m, 8.0) * np.sqrt(25.0 ** 2 / np.maximum(counts, 1) + (0.06 * mu) ** 2)
    y[counts == 0] = np.nan
    run_case(ref, "t_patseq_many_rows", "Model_t_patseq", y, two_group(m), [9.5, 600.0, 0.0035],
             context={"counts": counts}, test_coef=[1])


def transform_fixture():
    """fitnoise.transform: what of the reference runs without Theano.  `_unpack_transform`,
    `_apply_transform`, `_configured` are called as they are; the objective is the reference's own
    expression (fittransform.py:103-121) evaluated on arrays with the reference's `qr_complete`, `dot`
    and `log` instead of Theano symbols."""
    ref = refimport.load()
    ft = refimport.load_transform()
    rng = np.random.default_rng(20261021)
    n, m = 60, 6
    design = np.stack([np.ones(m), (np.arange(m) >= 3) * 1.0], axis=1)
    depth = rng.uniform(0.5, 2.0, size=m)
    mu = np.exp(rng.normal(3.0, 1.5, size=(n, 1)))
    x = rng.poisson(mu * depth[None, :] * np.exp(0.5 * design[:, 1])[None, :]).astype(float)
    out = dict(x=x, design=design)
    q, r = ref.env.qr_complete(design)
    q_null = q[:, design.shape[1]:]
    for order in (1, 2, 3):
        t = ft.Transform_polynomial(order)
        cfg = t._with(x=x, design=design, null=q_null)._configured()
        pack = np.concatenate([rng.uniform(0.0, 3.0, size=(order - 1) * m), rng.uniform(0.2, 5.0, size=m)])
        vy = cfg._apply_transform(cfg._unpack_transform(m, pack), x)
        vcentered = vy - vy.mean(axis=0)[None, :]
        vresidual = ft.dot(vcentered, q_null)
        cost = ft.log((vresidual * vresidual).flatten().sum()) - ft.log((vcentered * vcentered).flatten().sum())
        # The reference's objective as a function of pack (its own methods and numpy-level functions), its
        # central-difference gradient (what Theano's autodiff of the same expression returns, to ~1e-9),
        # and the reference's OWN driver loop (fittransform.py:140-156: up to 100 L-BFGS-B runs, stop when
        # the improvement is <= 1e-6) fed with that value and gradient: the fitted parameters and
        # objective the reference's Transform.fit arrives at, up to the noise of the numerical gradient.
        def ref_cost(pk, cfg=cfg):
            y_ = cfg._apply_transform(cfg._unpack_transform(m, pk), x)
            c_ = y_ - y_.mean(axis=0)[None, :]
            r_ = ft.dot(c_, q_null)
            return float(ft.log((r_ * r_).flatten().sum()) - ft.log((c_ * c_).flatten().sum()))

        def ref_grad(pk):
            g_ = np.zeros(len(pk))
            for i in range(len(pk)):
                h = 1e-6 * max(1.0, abs(pk[i]))
                up, dn = pk.copy(), pk.copy()
                up[i] += h
                dn[i] -= h
                g_[i] = (ref_cost(up) - ref_cost(dn)) / (2 * h)
            return g_

        assert abs(ref_cost(pack) - float(cost)) <= 1e-14
        out["ref_grad%d" % order] = ref_grad(pack)
        import scipy.optimize
        a, last = np.asarray(cfg._param_initial, dtype=float), None
        for _ in range(100):
            fit = scipy.optimize.minimize(lambda pk: (ref_cost(pk), ref_grad(pk)), a, method="L-BFGS-B", jac=True,
                                          bounds=cfg._param_bounds)
            a = fit.x
            if last is not None and last - fit.fun <= 1e-6:
                break
            last = fit.fun
        out["ref_fit_pack%d" % order] = fit.x.copy()
        out["ref_fit_fun%d" % order] = float(fit.fun)
        out["pack%d" % order] = pack
        out["ref_y%d" % order] = vy
        out["ref_cost%d" % order] = float(cost)
        out["ref_initial%d" % order] = np.asarray(cfg._param_initial, dtype=float)
        out["ref_lower%d" % order] = np.array([b[0] for b in cfg._param_bounds], dtype=float)
        out["ref_upper_is_none%d" % order] = np.array([b[1] is None for b in cfg._param_bounds])
    np.savez_compressed(os.path.join(HERE, "transform_poly.npz"), **out)
    print("transform_poly")


if __name__ == "__main__":
    import sys
    if len(sys.argv) > 1 and sys.argv[1] == "--transform":
        transform_fixture()
        sys.exit(0)
    if len(sys.argv) > 2 and sys.argv[1] == "--only":
        ONLY = set(sys.argv[2].split(","))
    main()
