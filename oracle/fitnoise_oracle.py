"""CPU oracle for the Fitnoise hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain numpy/scipy restatement of the reference algorithm (pfh/fitnoise 2.2, numpy path),
written to be *the checker* for the CUDA implementation in ``fitnoise_b200``:

  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
    ``--impl reference`` legs may import this module;
  * the product (``fitnoise_b200``) never imports it and has no CPU fallback.

It deliberately follows the reference's own (slow) route -- a Python loop over genes, a
complete QR of every retained sub-design, explicit ``inv`` and ``det`` of the
(k-c)x(k-c) transformed covariance -- and NOT the closed-form REML identities the CUDA
kernels use, so that agreement between the two is evidence and not tautology.  Because
it performs the same numpy/LAPACK/scipy calls per gene as the reference does, it is also
the "port" CPU baseline timed by ``bench.py``.

Parity status: PINNED against the reference itself.  ``tests/golden/make_golden.py``
imports the unmodified reference (py2->py3 syntax patched in memory) in the build
container and stores its outputs as ``tests/golden/*.npz``; ``tests/test_oracle_golden.py``
checks every function here against those fixtures.  The reference ships no tests or
golden vectors of its own.

Citations are to files under the reference tree, ``fitnoise/<file>:<lines>``.
"""
import numpy as np
import scipy.optimize
import scipy.special
import scipy.stats

INDEPENDENT_DF = 1e-10      # fitnoise.py:577  "Approximately zero"


# --------------------------------------------------------------------------------------
# Noise models: _configured / _unpack / _get_dist restated as plain functions.
# Each spec answers: aux rows, initial theta, bounds, sigma^2_j(theta, aux_row) and its
# partial derivatives (the latter is NOT in the reference -- Theano autodiff'd it,
# fitnoise.py:93 -- and is cross-checked against central differences in the tests).
# --------------------------------------------------------------------------------------

def _weights_aux(y, context):
    # fitnoise.py:539-544 / 581-584 / 620-625
    if "weights" in context:
        aux = 1.0 / np.asarray(context["weights"], dtype="float64")
    else:
        aux = np.ones(y.shape)
    assert aux.shape == y.shape
    return aux


def row_variance_mean(y, nan_aware=False):
    """fitnoise.py:546 / 627: ``numpy.mean(numpy.var(y, axis=1))``.

    The reference has no NaN handling here, so any NaN in y yields NaN (SURVEY 8a quirk
    table).  ``nan_aware=True`` is the documented divergence the product uses: population
    variance over the finite entries of each row, averaged over rows with >=1 finite
    entry.  Identical to the reference when y has no NaN.
    """
    if not nan_aware:
        return float(np.mean(np.var(y, axis=1)))
    fin = np.isfinite(y)
    cnt = fin.sum(axis=1)
    y0 = np.where(fin, y, 0.0)
    keep = cnt > 0
    mean = y0[keep].sum(axis=1) / cnt[keep]
    dev = np.where(fin[keep], y0[keep] - mean[:, None], 0.0)
    var = (dev * dev).sum(axis=1) / cnt[keep]
    return float(np.mean(var))


def _check_counts(y, counts):
    # fitnoise.py:647, 693, 745, 795, 851
    assert counts.shape == y.shape
    assert np.all(~np.isfinite(y)[counts == 0]), \
        "Tail lengths with zero count should be NaN (or NA in R)"


def _avg_tail(y, counts, v3):
    # fitnoise.py:747 / 797 (v2), 853 (v3: no max(1,.) on the numerator)
    y0 = y.copy()
    y0[~np.isfinite(y0)] = 0.0
    num = (y0 * counts).sum(axis=1)
    if not v3:
        num = np.maximum(1.0, num)
    return num / np.maximum(1.0, counts.sum(axis=1))


class Spec(object):
    """One normal-family noise model (the *_t_* variants add df in front, Model_t_mixin)."""
    def __init__(self, name, configure, sigma2, dsigma2, names):
        self.name, self.configure, self.sigma2, self.dsigma2, self.names = \
            name, configure, sigma2, dsigma2, names


def _cfg_normal(y, context, nan_aware):
    aux = _weights_aux(y, context)
    var = row_variance_mean(y, nan_aware)
    return aux, [var], [(var * 1e-6, var * 1e6)]           # fitnoise.py:548-552


def _cfg_independent(y, context, nan_aware):
    return _weights_aux(y, context), [], []                 # fitnoise.py:586-590


def _cfg_per_sample(y, context, nan_aware):
    m = y.shape[1]
    aux = _weights_aux(y, context)
    var = row_variance_mean(y, nan_aware)
    return aux, [var] * m, [(var * 1e-6, var * 1e6)] * m    # fitnoise.py:629-633


def _cfg_patseq_v1(per_sample):
    def configure(y, context, nan_aware):
        m = y.shape[1]
        counts = np.asarray(context["counts"], dtype="float64")
        _check_counts(y, counts)
        tails = y.copy()
        tails[~np.isfinite(tails)] = 0.0
        aux = np.concatenate([np.maximum(counts, 1.0), tails], axis=1)   # :648-652
        big = (1e-12, 1e12)
        if per_sample:
            return aux, [100.0] + [0.01] * m, [big] * (m + 1)            # :701-702
        return aux, [250.0, 0.01], [big, big]                            # :655-656
    return configure


def _cfg_patseq_v23(per_sample, v3):
    def configure(y, context, nan_aware):
        m = y.shape[1]
        counts = np.asarray(context["counts"], dtype="float64").copy()
        _check_counts(y, counts)
        avg = _avg_tail(y, counts, v3)
        aux = np.concatenate([np.maximum(counts, 1.0), avg[:, None]], axis=1)  # :749-750
        big = (1e-12, 1e12)
        if per_sample:
            return aux, [100.0] * m + [0.01] * m, [big] * (2 * m)        # :803-804
        if v3:
            return aux, [1.0, 0.01], [big, big]                          # :859-860
        return aux, [100.0, 0.01], [big, big]                            # :753-754
    return configure


def _s2_normal(t, aux):            # fitnoise.py:530-535
    return t[0] * aux


def _ds2_normal(t, aux):
    return aux[None, :].copy()


def _s2_independent(t, aux):       # fitnoise.py:573-578
    return aux.copy()


def _ds2_independent(t, aux):
    return np.zeros((0, len(aux)))


def _s2_per_sample(t, aux):        # fitnoise.py:609-614
    return t * aux


def _ds2_per_sample(t, aux):
    return np.diag(aux)


def _s2_patseq_v1(t, aux):         # fitnoise.py:671-678 and 717-724 (t[1:] scalar or vector)
    m = len(aux) // 2
    count, tail = aux[:m], aux[m:]
    s = t[1] if len(t) == 2 else t[1:]
    return t[0] / count + s * (tail * tail)


def _ds2_patseq_v1(t, aux):
    m = len(aux) // 2
    count, tail = aux[:m], aux[m:]
    if len(t) == 2:
        return np.stack([1.0 / count, tail * tail])
    return np.concatenate([(1.0 / count)[None, :], np.diag(tail * tail)], axis=0)


def _s2_patseq_v2(t, aux):         # fitnoise.py:769-776 and 825-832
    m = len(aux) - 1
    count, tail = aux[:m], aux[m]
    if len(t) == 2:
        return t[0] / count + t[1] * (tail * tail)
    return t[:m] / count + t[m:] * (tail * tail)


def _ds2_patseq_v2(t, aux):
    m = len(aux) - 1
    count, tail = aux[:m], aux[m]
    if len(t) == 2:
        return np.stack([1.0 / count, np.repeat(tail * tail, m)])
    return np.concatenate([np.diag(1.0 / count), np.eye(m) * (tail * tail)], axis=0)


def _s2_patseq_v3(t, aux):         # fitnoise.py:875-882
    m = len(aux) - 1
    count, tail = aux[:m], aux[m]
    return (t[0] / count + t[1]) * (tail * tail)


def _ds2_patseq_v3(t, aux):
    m = len(aux) - 1
    count, tail = aux[:m], aux[m]
    return np.stack([(tail * tail) / count, np.repeat(tail * tail, m)])


SPECS = {
    "normal": Spec("normal", _cfg_normal, _s2_normal, _ds2_normal, ["variance"]),
    "independent": Spec("independent", _cfg_independent, _s2_independent, _ds2_independent, []),
    "per_sample": Spec("per_sample", _cfg_per_sample, _s2_per_sample, _ds2_per_sample, ["variances"]),
    "patseq_v1": Spec("patseq_v1", _cfg_patseq_v1(False), _s2_patseq_v1, _ds2_patseq_v1,
                      ["read_variance", "sample_variance"]),
    "patseq_v1_per_sample": Spec("patseq_v1_per_sample", _cfg_patseq_v1(True), _s2_patseq_v1,
                                 _ds2_patseq_v1, ["read_variance", "sample_variance"]),
    "patseq_v2": Spec("patseq_v2", _cfg_patseq_v23(False, False), _s2_patseq_v2, _ds2_patseq_v2,
                      ["read_variance", "sample_variance"]),
    "patseq_v2_per_sample": Spec("patseq_v2_per_sample", _cfg_patseq_v23(True, False),
                                 _s2_patseq_v2, _ds2_patseq_v2, ["read_variance", "sample_variance"]),
    "patseq_v3": Spec("patseq_v3", _cfg_patseq_v23(False, True), _s2_patseq_v3, _ds2_patseq_v3,
                      ["read_variance", "sample_variance"]),
}

# Reference class name -> (spec, is_t).  fitnoise.py:555, 569, 635, 680, 726, 778, 834, 884, 889-892
MODELS = {
    "Model_normal": ("normal", False), "Model_t": ("normal", True),
    "Model_independent": ("independent", False),
    "Model_normal_per_sample": ("per_sample", False), "Model_t_per_sample": ("per_sample", True),
    "Model_normal_patseq_v1": ("patseq_v1", False), "Model_t_v1_patseq": ("patseq_v1", True),
    "Model_normal_patseq_v1_per_sample": ("patseq_v1_per_sample", False),
    "Model_t_patseq_v1_per_sample": ("patseq_v1_per_sample", True),
    "Model_normal_patseq_v2": ("patseq_v2", False), "Model_t_patseq_v2": ("patseq_v2", True),
    "Model_normal_patseq_v2_per_sample": ("patseq_v2_per_sample", False),
    "Model_t_patseq_v2_per_sample": ("patseq_v2_per_sample", True),
    "Model_normal_patseq_v3": ("patseq_v3", False), "Model_t_patseq_v3": ("patseq_v3", True),
    "Model_normal_patseq": ("patseq_v2", False), "Model_t_patseq": ("patseq_v2", True),
    "Model_normal_patseq_per_sample": ("patseq_v2_per_sample", False),
    "Model_t_patseq_per_sample": ("patseq_v2_per_sample", True),
    # Model_factors_mixin on top (fitnoise.py:558-565, 592-594); need n_factors at configure time
    "Model_normal_factors": ("normal", False), "Model_t_factors": ("normal", True),
    "Model_independent_factors": ("independent", False),
}


class Configured(object):
    """A model bound to a data set: what Model.fit_noise holds after _configured().

    ``n_factors`` > 0 adds Model_factors_mixin (fitnoise.py:439-518): unknown batch-effect
    factors f_i = Q2 . pack_i in the orthogonal complement of the design, covariance
    diag(sigma2) + sum_i f_i f_i', penalty sum_{i<j} (f_i . f_j)^2.  theta is then ordered
    [df if t] + [pack_1 .. pack_nf] + [the normal model's parameters] (MRO of
    Model_t_factors: t-mixin outermost, :563-565)."""
    def __init__(self, model, y, context=None, nan_aware=False, n_factors=0, design=None,
                 control_design=None, controls=None):
        spec_name, self.is_t = MODELS[model]
        self.model = model
        self.spec = SPECS[spec_name]
        self.y = np.asarray(y, dtype="float64")
        assert self.y.ndim == 2
        context = context or {}
        aux, initial, bounds = self.spec.configure(self.y, context, nan_aware)
        self.n_factors = int(n_factors)
        self.factor_Q2 = None
        if self.n_factors:
            # Model_factors_mixin._configured, fitnoise.py:490-518
            m = self.y.shape[1]
            if controls is not None and np.any(controls):
                fdesign = np.asarray(control_design, dtype="float64")
            else:
                fdesign = np.asarray(design, dtype="float64")
            Q, R = np.linalg.qr(fdesign, mode="complete")
            Q2 = Q[:, fdesign.shape[1]:]
            m2 = Q2.shape[1]
            good = np.all(np.isfinite(self.y), axis=1)
            goody = (Q2.T @ self.y[good].T).T
            u, sv, v = np.linalg.svd(goody, full_matrices=0)
            rows = np.argsort(-np.abs(sv))[:self.n_factors]
            finit = list((v[rows] * sv[rows][:, None]).flatten() / np.sqrt(goody.shape[0]))
            initial = finit + list(initial)
            bounds = [(None, None)] * (m2 * self.n_factors) + list(bounds)
            self.factor_Q2 = Q2
        if self.is_t:                                   # Model_t_mixin._configured :431-436
            initial = [5.0] + list(initial)
            bounds = [(1e-3, 100.0)] + list(bounds)
        self.aux = aux
        self.initial = np.asarray(initial, dtype="float64")
        self.bounds = list(bounds)

    def split(self, theta):
        """Model_t_mixin._unpack :410-416: df is theta[0] for t models; then the factor packs
        (Model_factors_mixin._unpack :447-458); the rest belongs to the normal model."""
        theta = np.asarray(theta, dtype="float64")
        if self.is_t:
            df, theta = float(theta[0]), theta[1:]
        elif self.spec.name == "independent":
            df = INDEPENDENT_DF
        else:
            df = None                                    # Mvnormal
        if self.n_factors:
            theta = theta[self.factor_Q2.shape[1] * self.n_factors:]
        return df, theta

    def factors(self, theta):
        """List of the factor vectors f_i (length m)."""
        if not self.n_factors:
            return []
        theta = np.asarray(theta, dtype="float64")
        at = 1 if self.is_t else 0
        m2 = self.factor_Q2.shape[1]
        return [self.factor_Q2 @ theta[at + i * m2: at + (i + 1) * m2] for i in range(self.n_factors)]

    def model_cost(self, theta):
        """Model_factors_mixin._model_cost :470-478."""
        f = self.factors(theta)
        return sum(float(f[i] @ f[j]) ** 2 for i in range(len(f)) for j in range(i))

    def model_cost_grad(self, theta):
        theta = np.asarray(theta, dtype="float64")
        grad = np.zeros(len(theta))
        f = self.factors(theta)
        if len(f) < 2:
            return grad
        at = 1 if self.is_t else 0
        m2 = self.factor_Q2.shape[1]
        for i in range(len(f)):
            gf = np.zeros(len(f[i]))
            for j in range(len(f)):
                if j != i:
                    gf += 2.0 * float(f[i] @ f[j]) * f[j]
            grad[at + i * m2: at + (i + 1) * m2] = self.factor_Q2.T @ gf
        return grad

    def sigma2(self, theta, row):
        df, rest = self.split(theta)
        return self.spec.sigma2(rest, self.aux[row])

    def dsigma2(self, theta, row):
        df, rest = self.split(theta)
        return self.spec.dsigma2(rest, self.aux[row])

    def covariance(self, theta, row):
        """_get_dist(...).covar: diag(sigma2) (+ sum_i f_i f_i', Model_factors_mixin._get_dist :481-487)."""
        covar = np.diag(self.sigma2(theta, row))
        for f in self.factors(theta):
            covar = covar + np.outer(f, f)
        return covar

    def df_of(self, theta):
        return self.split(theta)[0]


# --------------------------------------------------------------------------------------
# Distribution algebra pieces used by the path (distributions.py), on explicit matrices.
# --------------------------------------------------------------------------------------

def good_mask(sigma2):
    """Mv*.good (distributions.py:31-36, 127-132) for mean 0 and a diagonal covariance:
    sample j is usable iff row j of diag(sigma2) is finite, i.e. sigma2[j] is finite."""
    return np.isfinite(sigma2)


def log_density(S, z, df):
    """Mvnormal.log_density (distributions.py:39-49) when df is None, else
    Mvt.log_density (:135-147); mean is zero on this path.  Explicit inv and det,
    as the reference does (distributions.py:4-7, env.py:178-185)."""
    p = S.shape[0]
    Sinv = np.linalg.inv(S)
    d2 = float(z @ Sinv @ z)
    logdet = np.log(np.linalg.det(S))
    if df is None:
        return -0.5 * (np.log(2 * np.pi) * p + logdet + d2)
    v = df
    return (scipy.special.gammaln(0.5 * (v + p)) - scipy.special.gammaln(0.5 * v)
            - (0.5 * p) * np.log(np.pi * v) - 0.5 * logdet
            - (0.5 * (v + p)) * np.log(1 + d2 / v))


def tail_p_value(S, z, df):
    """Mvnormal.p_value (:52-57, chi2 sf) / Mvt.p_value (:150-155, F sf) of offset z."""
    p = S.shape[0]
    q = float(z @ np.linalg.inv(S) @ z)
    if df is None:
        return float(scipy.stats.chi2.sf(q, df=p))
    return float(scipy.stats.f.sf(q / p, dfn=p, dfd=df))


# --------------------------------------------------------------------------------------
# _fit_noise (fitnoise.py:17-160)
# --------------------------------------------------------------------------------------

def prepare_rows(cfg, designs, theta_for_good=None, skip_rank_deficient=False):
    """fitnoise.py:23-35.  Returns {row: (retain, tQ2, z2)} for rows with k > c.

    ``skip_rank_deficient=True`` is the product's documented divergence: rows whose
    retained sub-design has k > c but less than full column rank are left out of the REML sum,
    because the reference's term for them depends on LAPACK's Householder details, not on the
    mathematics.  Default False = the reference's behaviour."""
    y = cfg.y
    n, m = y.shape
    theta = cfg.initial if theta_for_good is None else theta_for_good
    out = {}
    for row in range(n):
        design = designs[row]
        keep = np.isfinite(y[row]) & good_mask(cfg.sigma2(theta, row))
        retain = np.arange(m, dtype="int32")[keep]
        c = design.shape[1]
        if len(retain) <= c:
            continue
        if skip_rank_deficient and np.linalg.matrix_rank(design[retain]) < c:
            continue
        Q, R = np.linalg.qr(design[retain], mode="complete")     # env.py:200
        tQ2 = Q[:, c:].T
        out[row] = (retain, tQ2, tQ2 @ y[row, retain])
    return out


def score_row(cfg, theta, row, retain, tQ2, z2):
    """fitnoise.py:44-50: -log density of z2 under marginal(retain).transformed(tQ2)."""
    covar = cfg.covariance(theta, row)[np.ix_(retain, retain)]   # marginal :84-86/:186-188
    S = tQ2 @ covar @ tQ2.T                                      # transformed :66-71/:167-173
    return -log_density(S, z2, cfg.df_of(theta))


def score_row_grad(cfg, theta, row, retain, tQ2, z2):
    """Value and gradient of score_row by matrix calculus on the reference's own S
    (what theano.gradient.grad would return, fitnoise.py:93).  Not the REML identities."""
    theta = np.asarray(theta, dtype="float64")
    df = cfg.df_of(theta)
    covar = cfg.covariance(theta, row)[np.ix_(retain, retain)]
    S = tQ2 @ covar @ tQ2.T
    p = S.shape[0]
    Sinv = np.linalg.inv(S)
    u = Sinv @ z2
    d2 = float(z2 @ u)
    M = tQ2.T @ Sinv @ tQ2                  # k x k
    tu = tQ2.T @ u
    kappa = 0.5 if (df is None) else 0.5 * (df + p) / (df + d2)
    g_cov = 0.5 * M - kappa * np.outer(tu, tu)         # d(-ll)/d covar[retain, retain]
    g_s2 = np.diag(g_cov)                              # d(-ll)/d sigma2_j, j in retain
    ds2 = cfg.dsigma2(theta, row)[:, retain]
    g_rest = ds2 @ g_s2
    if cfg.n_factors:                                  # d covar = df f' + f df'  ->  2 g_cov f ; f = Q2 pack
        Q2r = cfg.factor_Q2[retain]
        g_pack = [Q2r.T @ (2.0 * g_cov @ f[retain]) for f in cfg.factors(theta)]
        g_rest = np.concatenate(g_pack + [g_rest])
    value = -log_density(S, z2, df)
    if cfg.is_t:
        v = df
        g_v = (-0.5 * scipy.special.digamma(0.5 * (v + p)) + 0.5 * scipy.special.digamma(0.5 * v)
               + 0.5 * p / v + 0.5 * np.log1p(d2 / v)
               - 0.5 * (v + p) * (d2 / (v * v)) / (1 + d2 / v))
        return value, np.concatenate([[g_v], g_rest])
    return value, g_rest


def total_score(cfg, prepared, theta):
    """fitnoise.py:63-68: model cost (:204-205, :470-478) + sum of score_row."""
    return cfg.model_cost(theta) + sum(score_row(cfg, theta, row, *item) for row, item in prepared.items())


def total_score_grad(cfg, prepared, theta):
    value = cfg.model_cost(theta)
    grad = cfg.model_cost_grad(theta)
    for row, item in prepared.items():
        v, g = score_row_grad(cfg, theta, row, *item)
        value += v
        grad += g
    return value, grad


def optimize_loose(cfg, prepared):
    """The reference's numpy path: fitnoise.py:73-76 (numeric gradient, default tolerances)."""
    return scipy.optimize.minimize(lambda t: total_score(cfg, prepared, t), cfg.initial,
                                   method="L-BFGS-B", bounds=cfg.bounds)


def optimize_tight(value_grad, initial, bounds, max_rounds=50):
    """The reference's Theano-path driver loop, fitnoise.py:143-158: L-BFGS-B with
    jac=True, ftol=gtol=1e-12, restarted from the optimum until |x - initial| < 1e-6."""
    initial = np.asarray(initial, dtype="float64")
    for _ in range(max_rounds):
        res = scipy.optimize.minimize(value_grad, initial, method="L-BFGS-B", jac=True,
                                      bounds=bounds, options=dict(ftol=1e-12, gtol=1e-12))
        if np.all(np.abs(res.x - initial) < 1e-6):
            break
        initial = res.x
    return res


def fit_noise(cfg, design, control_design=None, controls=None, force_param=None, tight=True,
              skip_rank_deficient=False):
    """Model.fit_noise, fitnoise.py:207-291.  Returns a dict of the result attributes."""
    y = cfg.y
    n, m = y.shape
    design = np.asarray(design, dtype="float64")
    if control_design is None:
        control_design = np.ones((m, 1))
    if controls is None:
        controls = np.zeros(n, dtype=bool)
    control_design = np.asarray(control_design, dtype="float64")
    controls = np.asarray(controls, dtype=bool)
    designs = [control_design if flag else design for flag in controls]      # :233
    prepared = prepare_rows(cfg, designs, skip_rank_deficient=skip_rank_deficient)

    if force_param is not None:                                               # :38-41
        x = np.asarray(force_param, dtype="float64")
        assert len(x) == len(cfg.initial)
        fun = 0.0
    elif len(cfg.initial) == 0:                                               # :52-53
        x, fun = np.zeros(0), 0.0
    elif tight:
        res = optimize_tight(lambda t: total_score_grad(cfg, prepared, t), cfg.initial, cfg.bounds)
        x, fun = res.x, float(res.fun)
    else:
        res = optimize_loose(cfg, prepared)
        x, fun = res.x, float(res.fun)

    df_t = cfg.df_of(x)
    noise_p = np.repeat(np.nan, n)
    averages = np.repeat(np.nan, n)
    for row, (retain, tQ2, z2) in prepared.items():                           # :259-267
        S = tQ2 @ cfg.covariance(x, row)[np.ix_(retain, retain)] @ tQ2.T
        noise_p[row] = tail_p_value(S, z2, df_t)
        averages[row] = np.mean(y[row, retain])
    good_p = noise_p[np.isfinite(noise_p)]                                    # :270-274
    combined = 0.0 if not len(good_p) else min(1, np.minimum.reduce(good_p) * len(good_p))
    n_residuals = sum(m - d.shape[1] for d in designs)                        # :276
    df = n_residuals - len(cfg.initial)
    return dict(x=x, fun=fun, df=df, deviance=fun * 2.0, score=fun / (np.log(2.0) * df),
                noise_p_values=noise_p, noise_combined_p_value=combined, averages=averages,
                prepared=prepared)


# --------------------------------------------------------------------------------------
# fit_coef (fitnoise.py:294-347) and test (:368-396)
# --------------------------------------------------------------------------------------

def fit_coef(cfg, theta, design):
    """Returns coef (n x c), coef covariances (n x c x c), posterior df (n); NaN rows where
    the reference leaves coef_dists[row] = None."""
    y = cfg.y
    n, m = y.shape
    design = np.asarray(design, dtype="float64")
    c = design.shape[1]
    df = cfg.df_of(theta)
    coef = np.tile(np.nan, (n, c))
    covar = np.tile(np.nan, (n, c, c))
    dfs = np.tile(np.nan, n)
    for row in range(n):
        s2 = cfg.sigma2(theta, row)
        retain = np.arange(m)[np.isfinite(y[row]) & good_mask(s2)]            # :311-313
        k = len(retain)
        if k < c:
            continue                                                          # :315
        if np.linalg.matrix_rank(design[retain]) < c:
            continue                                                          # :316
        Q, R = np.linalg.qr(design[retain], mode="complete")
        Rinv = np.linalg.inv(R[:c])
        z = Q.T @ y[row, retain]                                              # :324
        T = Q.T @ cfg.covariance(theta, row)[np.ix_(retain, retain)] @ Q     # marginal+transformed
        T11, T12, T21, T22 = T[:c, :c], T[:c, c:], T[c:, :c], T[c:, c:]
        T22inv = np.linalg.inv(T22)
        G = T12 @ T22inv
        z1, z2 = z[:c], z[c:]
        cmean = G @ z2                                                        # conditional
        ccov = T11 - G @ T21
        p2 = k - c
        if df is not None:                                                    # Mvt :191-215
            ccov = ccov * ((df + z2 @ T22inv @ z2) / (df + p2))
            dfs[row] = df + p2
        coef[row] = (-Rinv) @ (cmean - z1)                                    # shifted, transformed
        covar[row] = Rinv @ ccov @ Rinv.T
    return coef, covar, dfs


def test_contrasts(coef, covar, dfs, is_normal, coef_idx=None, contrasts=None):
    """Model.test, fitnoise.py:368-396."""
    n, c = coef.shape
    if coef_idx is not None:
        assert contrasts is None
        contrasts = np.zeros((c, len(coef_idx)))
        for i, j in enumerate(coef_idx):
            contrasts[j, i] = 1.0
    contrasts = np.asarray(contrasts, dtype="float64")
    kc = contrasts.shape[1]
    contrasted = np.tile(np.nan, (n, kc))
    ccovar = np.tile(np.nan, (n, kc, kc))
    p = np.tile(np.nan, n)
    for row in range(n):
        if np.isnan(covar[row, 0, 0]):          # coef_dists[row] is None  (:385)
            continue
        mean = contrasts.T @ coef[row]
        cov = contrasts.T @ covar[row] @ contrasts
        contrasted[row], ccovar[row] = mean, cov
        off = 0.0 - mean
        q = float(off @ np.linalg.inv(cov) @ off)
        if is_normal:
            p[row] = scipy.stats.chi2.sf(q, df=kc)
        else:
            p[row] = scipy.stats.f.sf(q / kc, dfn=kc, dfd=dfs[row])
    return contrasted, ccovar, p, fdr(p)


def fdr(p):
    """p_adjust.py:3-21 Benjamini & Hochberg; NaN -> 1.0 and counted in n."""
    p = np.array(p, dtype="float64")
    p[np.isnan(p)] = 1.0
    n = len(p)
    order = np.argsort(p)[::-1]
    back = np.argsort(order)
    ps = p[order]
    qs = np.minimum(1.0, np.minimum.accumulate(ps * n / np.arange(n, 0, -1)))
    return qs[back]


def get_weights(cfg, theta):
    """Model.get_weights, fitnoise.py:399-404: 1/diag(covar) per gene."""
    return np.array([1.0 / np.diag(cfg.covariance(theta, row)) for row in range(cfg.y.shape[0])])


def fit_and_test(model, y, design, context=None, coef_idx=None, contrasts=None,
                 control_design=None, controls=None, noise_design=None, force_param=None,
                 nan_aware=False, tight=True, skip_rank_deficient=False, n_factors=0):
    """Model.fit + .test end to end (fitnoise.py:350-365, 368-396) as one dict."""
    nd = design if noise_design is None else noise_design
    cfg = Configured(model, y, context, nan_aware=nan_aware, n_factors=n_factors, design=nd,
                     control_design=control_design if control_design is not None else np.ones((np.shape(y)[1], 1)),
                     controls=controls)
    out = fit_noise(cfg, nd, control_design, controls, force_param, tight=tight,
                    skip_rank_deficient=skip_rank_deficient)
    coef, covar, dfs = fit_coef(cfg, out["x"], design)
    out.update(coef=coef, coef_covar=covar, coef_df=dfs)
    if coef_idx is not None or contrasts is not None:
        is_normal = cfg.df_of(out["x"]) is None
        contrasted, ccov, p, q = test_contrasts(coef, covar, dfs, is_normal, coef_idx, contrasts)
        out.update(contrasts=contrasted, contrast_covar=ccov, p_values=p, q_values=q)
    out["cfg"] = cfg
    return out


# --------------------------------------------------------------------------------------
# fitnoise.transform (fittransform.py).  PARITY PARTLY PINNED: the reference builds this objective from
# Theano symbols and Theano cannot be installed here, so `Transform.fit` itself (its autodiff gradient
# and its optimiser run) cannot be executed.  What does run without Theano is pinned by
# tests/golden/transform_poly.npz (tests/test_transform.py): `_apply_transform`, `_unpack_transform`,
# `_configured` are the reference's own code on arrays, and the objective is the reference's
# expression (fittransform.py:103-121) evaluated with its own qr_complete / dot / log.  UNPINNED: the
# gradient (checked against central differences instead) and the fitted parameters.
# --------------------------------------------------------------------------------------

def transform_apply(pack, x, order):
    """Transform_polynomial._unpack_transform / _apply_transform, fittransform.py:180-194."""
    m = x.shape[1]
    poly = 1.0
    for i in range(order):
        poly = poly * x + pack[i * m:(i + 1) * m][None, :]
    return np.log(poly) * (1.0 / (order * np.log(2.0)))


def transform_cost(pack, x, design, order):
    """vcost of Transform.fit, fittransform.py:103-121."""
    q, r = np.linalg.qr(design, mode="complete")
    q_null = q[:, design.shape[1]:]
    y = transform_apply(pack, x, order)
    centered = y - y.mean(axis=0)[None, :]
    residual = centered @ q_null
    return float(np.log((residual * residual).sum()) - np.log((centered * centered).sum()))


def transform_cost_grad(pack, x, design, order):
    """The cost and its gradient by matrix calculus on the reference's own quantities (what
    theano.gradient.grad returns at fittransform.py:137)."""
    pack = np.asarray(pack, dtype="float64")
    n, m = x.shape
    q, r = np.linalg.qr(design, mode="complete")
    q_null = q[:, design.shape[1]:]
    poly = 1.0
    for i in range(order):
        poly = poly * x + pack[i * m:(i + 1) * m][None, :]
    scale = 1.0 / (order * np.log(2.0))
    y = np.log(poly) * scale
    centered = y - y.mean(axis=0)[None, :]
    residual = centered @ q_null
    v1, v0 = (residual * residual).sum(), (centered * centered).sum()
    g_centered = 2.0 * (residual @ q_null.T) / v1 - 2.0 * centered / v0
    g_y = g_centered - g_centered.mean(axis=0)[None, :]
    grad = np.zeros(order * m)
    for i in range(order):
        dy = scale * x ** (order - 1 - i) / poly
        grad[i * m:(i + 1) * m] = (g_y * dy).sum(axis=0)
    return float(np.log(v1) - np.log(v0)), grad


def transform_fit(x, design=None, order=2):
    """Transform.fit's optimiser loop, fittransform.py:140-163."""
    x = np.asarray(x, dtype="float64")
    n, m = x.shape
    if design is None:
        design = np.ones((m, 1))
    design = np.asarray(design, dtype="float64")
    a = np.asarray([0.0] * (order - 1) * m + [1.0] * m)
    bounds = [(0.0, None)] * (order - 1) * m + [(1e-10, None)] * m
    last = None
    for _ in range(100):
        fit = scipy.optimize.minimize(lambda p: transform_cost_grad(p, x, design, order), a,
                                      method="L-BFGS-B", jac=True, bounds=bounds)
        a = fit.x
        if last is not None and last - fit.fun <= 1e-6:
            break
        last = fit.fun
    y = transform_apply(a, x, order)
    return fit, y - y.mean(axis=0)
