"""Public model layer: ``Dataset``, ``Model`` and the noise-model classes, with the reference's
names, arguments, result attributes and text formats (reference: fitnoise/fitnoise.py:164-892).

What differs from the reference is where the work happens.  The reference loops over genes in
Python (fitnoise.py:23-35, 132-133, 259-267, 308-341, 384-389); here each of those loops is one
CUDA kernel over all genes (csrc/fn_kernels.cuh) reached through the C ABI
(include/fitnoise_b200.h), and only the hyper-parameter vector theta lives on the host, inside the
same scipy L-BFGS-B driver loop the reference's Theano path uses (fitnoise.py:143-158).

Noise-model plug-ins keep the reference's interface (``_unpack``, ``_describe_noise``,
``_get_dist``, composition by mix-in) and add ONE declaration, ``_family(m)``: the variance-family
descriptor the kernels consume.  A subclass that overrides ``_get_dist`` without a ``_family`` is
rejected -- there is no per-gene Python/CPU fallback.
"""
import contextlib
import sys

import numpy as np
import scipy.optimize

from . import _native
from .distributions import Mvnormal, Mvt
from .env import Withable, as_matrix, as_vector
from .session import Session

INDEPENDENT_DF = 1e-10      # reference fitnoise.py:577


class Dataset(object):
    """Expression matrix (genes x samples, NaN = missing) plus per-observation context:
    ``"weights"`` (precision weights) or ``"counts"`` (PAT-Seq read counts).  fitnoise.py:164-167"""

    def __init__(self, y, context={}):
        self.y = as_matrix(y)
        self.context = context


def as_dataset(dataset):
    return dataset if isinstance(dataset, Dataset) else Dataset(dataset)


class _LazyDists(object):
    """Sequence of per-gene distribution objects built on indexing.  The reference materialises n
    Python objects holding m x m / c x c matrices (fitnoise.py:252-255, 301, 380); at 1M genes
    that is tens of GB, so rows are constructed on demand from the result arrays."""
    __fn_lazy_sequence__ = True

    def __init__(self, n, make):
        self._n, self._make = int(n), make

    def __len__(self):
        return self._n

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(self._n))]
        i = int(i)
        if i < 0:
            i += self._n
        if not 0 <= i < self._n:
            raise IndexError(i)
        return self._make(i)

    def __iter__(self):
        return (self._make(i) for i in range(self._n))


class _Lazy(object):
    """A value computed on first use."""

    def __init__(self, compute):
        self._compute, self._value = compute, None

    def get(self):
        if self._compute is not None:
            self._value, self._compute = self._compute(), None
        return self._value


def _row_means(y):
    """Row means over the finite entries (NaN for an all-missing row)."""
    finite = np.isfinite(y)
    count = finite.sum(axis=1)
    total = np.where(finite, y, 0.0).sum(axis=1)
    return np.where(count > 0, total / np.maximum(count, 1), np.nan)


def _optimize(value_grad, initial, bounds, verbose, session=None):
    """The reference's driver loop (fitnoise.py:143-158): L-BFGS-B with analytic gradient and
    ftol = gtol = 1e-12, restarted from the optimum until it moves by less than 1e-6.  For the
    duration of the loop the session keeps its evaluation kernel resident on the GPU."""
    initial = np.asarray(initial, dtype="float64")
    with (session.resident() if session is not None else contextlib.nullcontext()):
        while True:
            result = scipy.optimize.minimize(
                value_grad, initial, method="L-BFGS-B", jac=True, bounds=bounds,
                options=dict(ftol=1e-12, gtol=1e-12))
            if verbose:
                print(result)
            if np.all(np.abs(result.x - initial) < 1e-6):
                return result
            initial = result.x


class Model(Withable):
    """
        score = mean residual log2 density
    """

    def __init__(self):
        pass

    def __repr__(self):
        if hasattr(self, "param"):
            result = self._describe_noise(self.param)
            result += "noise combined p-value = %f\n" % self.noise_combined_p_value
            result += "noise fit score = %f bits\n" % self.score
        else:
            result = "<%s>\n" % self.__class__.__name__
        return result

    @property
    def description(self):
        return repr(self)

    # ---- plug-in interface ---------------------------------------------------------------
    def _unpack(self, packed):
        return Withable()

    def _describe_noise(self, param):
        return ""

    def _model_cost(self, param):
        return 0.0

    def _family(self, m):
        """Variance-family descriptor (``_native.Family``) for m samples."""
        raise NotImplementedError(
            "%s does not declare a variance family (_family); fitnoise_b200 evaluates noise models "
            "only through its CUDA kernels and has no per-gene Python fallback" % type(self).__name__)

    _aux_key = None                     # which Dataset.context entry is the second n x m array
    _has_model_cost = False             # True for plug-ins with a _model_cost / factor vectors (_split_theta, _cost_grad)

    # theta as the optimiser sees it <-> what the kernels take: the variance parameters (and df)
    # plus, for the factor models, the m x nf matrix of factor vectors.  Composed by the mix-ins.
    def _split_theta(self, theta):
        return theta, None

    def _join_grad(self, grad_core, grad_factors):
        return grad_core

    def _cost_grad(self, theta):
        """_model_cost of the packed theta and its gradient (0 for every model without factors)."""
        return 0.0, np.zeros(len(theta))

    def _initial_bounds(self, m, row_variance_mean):
        return [], []

    def _aux_row(self, i):
        """Row i of the reference's ``_aux`` matrix (only for the lazy ``noise_dists`` views)."""
        return np.zeros(0)

    def _get_dist(self, param, aux):
        raise NotImplementedError

    # ---- configuration: upload + device-side preparation -----------------------------------
    def _configured(self):
        y = self.data.y
        n, m = y.shape
        family = self._family(m)
        aux = None
        if self._aux_key is not None:
            if self._aux_key in self.data.context:
                aux = as_matrix(self.data.context[self._aux_key])
                assert aux.shape == y.shape
            elif self._aux_key == "counts":
                raise KeyError("counts")
        session = Session(y, aux, family,
                          self.controls if np.any(self.controls) else None,
                          self.noise_design, self.control_design,
                          engine_factory=getattr(self, "_engine_factory", None))
        if family.kind == _native.KIND_PATSEQ:
            assert session.n_bad_counts == 0, \
                "Tail lengths with zero count should be NaN (or NA in R)"
        initial, bounds = self._initial_bounds(m, session.row_variance_mean)
        return self._with(_session=session, _initial=list(initial), _bounds=list(bounds))

    def _objective(self, theta):
        """REML objective (model cost + sum of score_row over all genes, fitnoise.py:63-68) and its
        gradient at the packed theta, evaluated on the GPU: (value, grad, n_used, n_bad)."""
        theta = as_vector(theta)
        if not self._has_model_cost:
            # no factor vectors, no model cost: theta goes to the kernels as it is (the split / join /
            # cost plumbing below costs several numpy calls per evaluation, which shows at ~10 us per evaluation)
            value, grad, _, used, bad = self._session.nll_grad(theta, None)
            return value, grad, used, bad
        core, factors = self._split_theta(theta)
        value, grad_core, grad_factors, used, bad = self._session.nll_grad(core, factors)
        cost, cost_grad = self._cost_grad(theta)
        return value + cost, self._join_grad(grad_core, grad_factors) + cost_grad, used, bad

    # ---- fit_noise (fitnoise.py:207-291) ---------------------------------------------------
    def fit_noise(self, data, design=None, control_design=None, controls=None, use_theano=True,
                  verbose=False, force_param=None):
        """``use_theano`` is accepted for signature compatibility and ignored."""
        data = as_dataset(data)
        design = as_matrix(design)                       # design is effectively required (:212)

        n, m = data.y.shape
        if control_design is None:
            control_design = np.ones((m, 1))
        if controls is None:
            controls = np.zeros(n, dtype="bool")          # the reference builds [False]*n (:217)
        control_design = as_matrix(control_design)
        controls = as_vector(controls, "bool")
        assert design.shape[0] == m and control_design.shape[0] == m and len(controls) == n

        result = self._with(data=data, noise_design=design, control_design=control_design,
                            controls=controls)
        result = result._configured()
        session = result._session
        initial = as_vector(result._initial)

        evaluations = [0]

        def value_grad(theta):
            evaluations[0] += 1
            if verbose:
                print(theta, end=" ")
                sys.stdout.flush()
            value, grad, used, bad = result._objective(theta)
            if bad:
                value = np.inf
            if verbose:
                print("->", value)
            return value, grad

        if force_param is not None:                                   # :38-41
            force_param = as_vector(force_param)
            assert len(force_param) == len(initial)
            fit = Withable()._with(x=force_param, fun=0.0)
        elif len(initial) == 0:                                       # :52-53
            fit = Withable()._with(x=np.zeros(0), fun=0.0)
        else:
            assert np.all(np.isfinite(initial)), "initial hyper-parameters are not finite"
            fit = _optimize(value_grad, initial, result._bounds, verbose, session)
        param = result._unpack(as_vector(fit.x))
        if verbose:
            print("Model cost:", result._model_cost(param))

        theta, factors = result._split_theta(as_vector(fit.x))
        if factors is not None:
            session.set_factors(factors)
        noise_p_values, averages = session.noise_stats(theta)         # :257-267

        # :270-274 -- min(1, min(good p) * number of good p), without the boolean-indexed copy of the
        # reference (3 ms of a 20 ms fit at a million genes)
        finite = np.isfinite(noise_p_values)
        n_good = int(np.count_nonzero(finite))
        if not n_good:
            noise_combined_p_value = 0.0
        else:
            # (p-values are NaN or in [0, 1]: fmin skips the NaN, and an infinity could never be the minimum)
            noise_combined_p_value = min(1, float(np.fmin.reduce(noise_p_values)) * n_good)

        n_controls = int(np.count_nonzero(controls))                  # :276-277
        n_residuals = (n - n_controls) * (m - design.shape[1]) + n_controls * (m - control_design.shape[1])
        df = n_residuals - len(initial)

        fitted = result._with(
            optimization_result=fit,
            param=param,
            df=df,
            deviance=fit.fun * 2.0,
            score=fit.fun / (np.log(2.0) * df),
            noise_p_values=noise_p_values,
            noise_combined_p_value=noise_combined_p_value,
            averages=averages,
            _theta=theta,
            _evaluations=evaluations[0],
        )
        # (closes over `result`, not `fitted`: no reference cycle through the fit result)
        fitted.noise_dists = _LazyDists(n, lambda i: result._get_dist(param, result._aux_row(i)))
        return fitted

    # ---- fit_coef (fitnoise.py:294-347) ----------------------------------------------------
    def fit_coef(self, design=None):
        if design is None:
            design = self.noise_design
        design = as_matrix(design)
        token = object()
        session, theta = self._session, self._theta
        coef = session.coef(theta, design, token)
        is_t = self._family(self.data.y.shape[1]).dist != _native.DIST_NORMAL
        # the c x c posterior covariances (and df) are fetched from the GPU only if somebody looks
        posterior = _Lazy(lambda: session.coef_covar(theta, design, token))

        def make(i):
            covar, dfpost = posterior.get()
            if covar[i, 0, 0] != covar[i, 0, 0]:
                return None
            if is_t:
                return Mvt(coef[i], covar[i], dfpost[i])
            return Mvnormal(coef[i], covar[i])

        return self._with(design=design, coef=coef, coef_dists=_LazyDists(len(coef), make),
                          _coef_posterior=posterior, _coef_token=token)

    @property
    def _coef_covar(self):
        return self._coef_posterior.get()[0]

    @property
    def _coef_df(self):
        return self._coef_posterior.get()[1]

    # ---- fit (fitnoise.py:350-365) ---------------------------------------------------------
    def fit(self, data, design=None, noise_design=None, control_design=None, controls=None,
            use_theano=True, verbose=False, force_param=None):
        return (self.fit_noise(data=data,
                               design=noise_design if noise_design is not None else design,
                               control_design=control_design, controls=controls,
                               use_theano=use_theano, verbose=verbose, force_param=force_param)
                .fit_coef(design))

    # ---- test (fitnoise.py:368-396) --------------------------------------------------------
    def test(self, coef=None, contrasts=None):
        names = None
        if coef is not None:
            assert contrasts is None
            contrasts = np.zeros((self.design.shape[1], len(coef)))
            for i in range(len(coef)):
                contrasts[coef[i], i] = 1.0
            names = ["coef%d" % (int(i) + 1) for i in coef]       # old_fitnoise.R:430, 522-531
        assert contrasts is not None, "No coefficients or contrasts to test given."
        contrasts = as_matrix(contrasts)
        assert contrasts.shape[0] == self.design.shape[1]

        session = self._session
        if session.coef_token is not self._coef_token:        # another fit_coef ran on this session since
            session.coef(self._theta, self.design, self._coef_token)
        contrasted, p_values, q_values, order = session.test(contrasts)
        if names is None:
            names = ["contrast%d" % (i + 1) for i in range(contrasts.shape[1])]   # old_fitnoise.R:39-47

        posterior, mean = self._coef_posterior, self.coef
        is_t = self._family(self.data.y.shape[1]).dist != _native.DIST_NORMAL

        def make(i):
            covar, dfpost = posterior.get()
            if covar[i, 0, 0] != covar[i, 0, 0]:
                return None
            base = Mvt(mean[i], covar[i], dfpost[i]) if is_t else Mvnormal(mean[i], covar[i])
            return base.transformed(contrasts.T)

        return self._with(contrasts=contrasted, contrast_dists=_LazyDists(len(mean), make),
                          p_values=p_values, q_values=q_values,
                          _order=order, _contrast_names=names)       # (uint32 from the device sort; widened on use)

    # ---- top table (legacy R test.fit, backyard/old_fitnoise.R:783-826) ---------------------
    def top_table(self, sort=True, genes=None):
        """Table of the tested contrasts in the layout of the legacy R ``test.fit`` (intended there
        to match limma's topTableF; on a fit that has not been tested it is the table of the NOISE
        fit, ``P.Value`` = ``noise.P.Value``, as ``test.fit(fit)`` without coefs or contrasts):
        one column per contrast, ``AveExpr``, ``P.Value``,
        ``adj.P.Val``, ``noise.P.Value`` and, when control genes were given, ``control``; rows
        ordered by p (ascending, ties in input order, missing p last -- R's ``order()``).  The order
        comes from the device radix sort that also produced the q-values.

        Returns a pandas DataFrame when pandas is importable, else a dict of columns plus "row".
        ``genes``: optional dict of per-gene annotation columns appended to the table."""
        if hasattr(self, "p_values"):
            p_values, q_values, order = self.p_values, self.q_values, self._order
            names, contrasts = self._contrast_names, self.contrasts
        else:
            # neither coefficients nor contrasts were tested: "this tests the fit of the noise"
            # (old_fitnoise.R:776-795): the table of the noise p-values, adjusted and ordered on the GPU
            assert hasattr(self, "noise_p_values"), "top_table() needs a fitted model"
            p_values = self.noise_p_values
            q_values, order = self._session.bh_ranked(p_values)
            names, contrasts = [], None
        n = len(p_values)
        rows = np.asarray(order, dtype=np.int64) if sort else np.arange(n)
        columns = {}
        for i, name in enumerate(names):
            columns[name] = contrasts[rows, i]
        with np.errstate(invalid="ignore"):
            columns["AveExpr"] = _row_means(self.data.y)[rows]          # rowMeans(fit$data, na.rm=TRUE)
        columns["P.Value"] = p_values[rows]
        columns["adj.P.Val"] = q_values[rows]
        columns["noise.P.Value"] = self.noise_p_values[rows]
        if np.any(self.controls):
            columns["control"] = np.asarray(self.controls)[rows]
        for key, value in (genes or {}).items():
            columns[key] = np.asarray(value)[rows]
        try:
            import pandas
        except ImportError:
            columns["row"] = rows
            return columns
        return pandas.DataFrame(columns, index=rows)

    # ---- get_weights (fitnoise.py:399-404) -------------------------------------------------
    def get_weights(self):
        """ Weight matrix for limma. """
        return self._session.weights(self._theta)


# ------------------------------------------------------------------------------------------------
# mix-ins and noise models
# ------------------------------------------------------------------------------------------------

class Model_t_mixin(Model):
    """ This mix-in converts an Mv_normal model to an Mvt model (fitnoise.py:407-436). """

    def _unpack(self, pack):
        df = pack[0]
        return super(Model_t_mixin, self)._unpack(pack[1:])._with(df=df)

    def _describe_noise(self, param):
        return "df = %f\n" % param.df + super(Model_t_mixin, self)._describe_noise(param)

    def _get_dist(self, param, aux):
        inner = super(Model_t_mixin, self)._get_dist(param, aux)
        assert type(inner) is Mvnormal
        return Mvt(inner.mean, inner.covar, param.df)

    def _family(self, m):
        family = super(Model_t_mixin, self)._family(m)
        assert family.dist == _native.DIST_NORMAL
        family.dist = _native.DIST_T
        return family

    def _initial_bounds(self, m, row_variance_mean):
        initial, bounds = super(Model_t_mixin, self)._initial_bounds(m, row_variance_mean)
        return [5.0] + list(initial), [(1e-3, 100.0)] + list(bounds)

    def _split_theta(self, theta):
        core, factors = super(Model_t_mixin, self)._split_theta(theta[1:])
        return np.concatenate([theta[:1], core]), factors

    def _join_grad(self, grad_core, grad_factors):
        return np.concatenate([grad_core[:1], super(Model_t_mixin, self)._join_grad(grad_core[1:], grad_factors)])

    def _cost_grad(self, theta):
        cost, grad = super(Model_t_mixin, self)._cost_grad(theta[1:])
        return cost, np.concatenate([[0.0], grad])


class Model_factors_mixin(Model):
    """ This mix-in adds unknown batch effect(s) to the model (fitnoise.py:439-518): the covariance
        of every gene gains sum_i f_i f_i' with f_i = Q2 . pack_i in the orthogonal complement of
        the (control) design.  The kernels fit it as the diagonal model with the design augmented by
        the factor vectors and a unit ridge on them (csrc/fn_gene.cuh, eval_factors); the penalty
        sum_{i<j} (f_i . f_j)^2 and the chain rule through Q2 are host-side (m x nf numbers). """

    _has_model_cost = True

    def __init__(self, n_factors, *etc):
        self.n_factors = n_factors
        super(Model_factors_mixin, self).__init__(*etc)

    def _unpack(self, pack):
        m2 = self._factor_Q2.shape[1]
        factors = []
        for i in range(self.n_factors):
            factors.append(self._factor_Q2 @ pack[:m2])
            pack = pack[m2:]
        return super(Model_factors_mixin, self)._unpack(pack)._with(factors=factors)

    def _describe_noise(self, param):
        result = ""
        for i in range(self.n_factors):
            result += "factor: [%s]\n" % ", ".join("%f" % item for item in param.factors[i])
        return result + super(Model_factors_mixin, self)._describe_noise(param)

    def _model_cost(self, param):
        """ It should always be possible to reduce this cost to zero. """
        cost = super(Model_factors_mixin, self)._model_cost(param)
        for i in range(self.n_factors):
            for j in range(i):
                cost = cost + float(np.dot(param.factors[i], param.factors[j])) ** 2
        return cost

    def _get_dist(self, param, aux):
        covar = np.zeros((len(aux), len(aux)))
        for i in range(self.n_factors):
            covar = covar + np.outer(param.factors[i], param.factors[i])
        return super(Model_factors_mixin, self)._get_dist(param, aux).plus_covar(covar)

    def _family(self, m):
        family = super(Model_factors_mixin, self)._family(m)
        assert family.kind == _native.KIND_SCALE1, "factors combine with Model_normal/_t/_independent only"
        family.nf = int(self.n_factors)
        return family

    def _configured(self):
        # fitnoise.py:490-518
        design = self.control_design if np.any(self.controls) else self.noise_design
        Q, R = np.linalg.qr(design, mode="complete")
        return super(Model_factors_mixin, self._with(_factor_Q2=Q[:, design.shape[1]:]))._configured()

    def _initial_bounds(self, m, row_variance_mean):
        initial, bounds = super(Model_factors_mixin, self)._initial_bounds(m, row_variance_mean)
        Q2 = self._factor_Q2
        m2 = Q2.shape[1]
        y = self.data.y
        good = np.all(np.isfinite(y), axis=1)
        goody = y[good] @ Q2                                          # (Q2' y')' of :507
        if goody.size <= 20000000:
            u, sv, v = np.linalg.svd(goody, full_matrices=0)
        else:       # same right singular vectors (up to sign) from the m2 x m2 Gram matrix
            evals, evecs = np.linalg.eigh(goody.T @ goody)
            sv, v = np.sqrt(np.maximum(evals, 0.0)), evecs.T
        rows = np.argsort(-np.abs(sv))[:self.n_factors]
        start = list((v[rows] * sv[rows][:, None]).flatten() / np.sqrt(goody.shape[0]))
        return start + list(initial), [(None, None)] * (m2 * self.n_factors) + list(bounds)

    def _split_theta(self, theta):
        m2 = self._factor_Q2.shape[1]
        npack = m2 * self.n_factors
        core, nothing = super(Model_factors_mixin, self)._split_theta(theta[npack:])
        packs = np.asarray(theta[:npack]).reshape(self.n_factors, m2)
        return core, np.ascontiguousarray(self._factor_Q2 @ packs.T)            # m x nf

    def _join_grad(self, grad_core, grad_factors):
        packs = (self._factor_Q2.T @ grad_factors).T.reshape(-1)               # chain rule through Q2
        return np.concatenate([packs, super(Model_factors_mixin, self)._join_grad(grad_core, None)])

    def _cost_grad(self, theta):
        m2 = self._factor_Q2.shape[1]
        npack = m2 * self.n_factors
        cost, grad_rest = super(Model_factors_mixin, self)._cost_grad(theta[npack:])
        F = self._factor_Q2 @ np.asarray(theta[:npack]).reshape(self.n_factors, m2).T
        gram = F.T @ F
        grad_f = np.zeros_like(F)
        for i in range(self.n_factors):
            for j in range(self.n_factors):
                if i != j:
                    grad_f[:, i] += 2.0 * gram[i, j] * F[:, j]
                    if j < i:
                        cost = cost + gram[i, j] ** 2
        return cost, np.concatenate([(self._factor_Q2.T @ grad_f).T.reshape(-1), grad_rest])


class _Weighted(Model):
    """aux = 1/weights or ones (fitnoise.py:539-544); the device gets the precision weights."""
    _aux_key = "weights"

    def _aux_row(self, i):
        if "weights" in self.data.context:
            with np.errstate(divide="ignore"):
                return 1.0 / as_matrix(self.data.context["weights"])[i]
        return np.ones(self.data.y.shape[1])


class Model_normal(_Weighted):
    def _unpack(self, pack):
        return Withable()._with(variance=pack[0])

    def _describe_noise(self, param):
        return "sd = %f\n" % np.sqrt(param.variance)

    def _get_dist(self, param, aux):
        return Mvnormal(np.zeros(len(aux)), np.diag(param.variance * aux))

    def _family(self, m):
        return _native.Family(_native.KIND_SCALE1, 1, 0, 0, 0, _native.DIST_NORMAL, 0.0)

    def _initial_bounds(self, m, var):
        # reference: var = mean(var(y, axis=1)) (fitnoise.py:546); NaN-aware here (divergences)
        return [var], [(var * 1e-6, var * 1e6)]


class Model_t(Model_t_mixin, Model_normal):
    pass


class Model_independent(_Weighted):
    def _describe_noise(self, param):
        return ""

    def _get_dist(self, param, aux):
        return Mvt(np.zeros(len(aux)), np.diag(aux), INDEPENDENT_DF)

    def _family(self, m):
        return _native.Family(_native.KIND_SCALE1, 0, 0, 0, 0, _native.DIST_T_FIXED, INDEPENDENT_DF)


class Model_normal_per_sample(_Weighted):
    def _unpack(self, pack):
        return Withable()._with(variances=pack)

    def _describe_noise(self, param):
        return "sd = [%s]\n" % ", ".join("%f" % item for item in np.sqrt(param.variances))

    def _get_dist(self, param, aux):
        return Mvnormal(np.zeros(len(aux)), np.diag(param.variances * aux))

    def _family(self, m):
        return _native.Family(_native.KIND_SCALE_PS, m, 0, 0, 0, _native.DIST_NORMAL, 0.0)

    def _initial_bounds(self, m, var):
        return [var] * m, [(var * 1e-6, var * 1e6)] * m


class Model_t_per_sample(Model_t_mixin, Model_normal_per_sample):
    pass


class _Patseq(Model):
    """ Differential tail length detection in PAT-Seq: variance = read term / reads + sample term.
        The device gets the raw counts and derives 1/max(count,1) and the average tail itself. """
    _aux_key = "counts"
    _per_sample_a = False
    _per_sample_b = False
    _tail_own = False
    _v3 = False
    _start = (100.0, 0.01)

    def _family(self, m):
        return _native.Family(_native.KIND_PATSEQ, m if self._per_sample_a else 1,
                              m if self._per_sample_b else 1, int(self._tail_own), int(self._v3),
                              _native.DIST_NORMAL, 0.0)

    def _initial_bounds(self, m, var):
        na = m if self._per_sample_a else 1
        nb = m if self._per_sample_b else 1
        initial = [self._start[0]] * na + [self._start[1]] * nb
        return initial, [(1e-12, 1e12)] * (na + nb)

    def _unpack(self, pack):
        na = len(pack) // 2 if self._per_sample_a else 1
        read = pack[:na] if self._per_sample_a else pack[0]
        sample = pack[na:] if self._per_sample_b else pack[na]
        return Withable()._with(read_variance=read, sample_variance=sample)

    def _aux_row(self, i):
        y = self.data.y[i]
        counts = as_matrix(self.data.context["counts"])[i]
        y0 = np.where(np.isfinite(y), y, 0.0)
        if self._tail_own:
            return np.concatenate([np.maximum(counts, 1.0), y0])                 # :648-652
        num = (y0 * counts).sum()
        if not self._v3:
            num = max(1.0, num)                                                   # :747 vs :853
        avg = num / max(1.0, counts.sum())
        return np.concatenate([np.maximum(counts, 1.0), [avg]])                   # :749-750

    def _get_dist(self, param, aux):
        if self._tail_own:
            m = len(aux) // 2
            count, tail = aux[:m], aux[m:]
        else:
            m = len(aux) - 1
            count, tail = aux[:m], aux[m]
        if self._v3:
            s2 = (param.read_variance / count + param.sample_variance) * (tail * tail)
        else:
            s2 = param.read_variance / count + param.sample_variance * (tail * tail)
        return Mvnormal(np.zeros(m), np.diag(s2 * np.ones(m)))


class Model_normal_patseq_v1(_Patseq):
    _tail_own = True
    _start = (250.0, 0.01)

    def _describe_noise(self, param):
        return "variance = %f^2 / reads + (%f * tail)^2\n" % (
            np.sqrt(param.read_variance), np.sqrt(param.sample_variance))


class Model_t_v1_patseq(Model_t_mixin, Model_normal_patseq_v1):
    pass


Model_t_patseq_v1 = Model_t_v1_patseq      # the reference's class name is transposed (fitnoise.py:680)


class Model_normal_patseq_v1_per_sample(_Patseq):
    _tail_own = True
    _per_sample_b = True

    def _describe_noise(self, param):
        return "variance = %f^2 / reads + (sample_cv * tail)^2\nsample_cv = [%s]\n" % (
            np.sqrt(param.read_variance),
            ", ".join("%f" % item for item in np.sqrt(param.sample_variance)))


class Model_t_patseq_v1_per_sample(Model_t_mixin, Model_normal_patseq_v1_per_sample):
    pass


class Model_normal_patseq_v2(_Patseq):
    def _describe_noise(self, param):
        return "variance = %f^2 / reads + (%f * avgtail)^2\n" % (
            np.sqrt(param.read_variance), np.sqrt(param.sample_variance))


class Model_t_patseq_v2(Model_t_mixin, Model_normal_patseq_v2):
    pass


class Model_normal_patseq_v2_per_sample(_Patseq):
    _per_sample_a = True
    _per_sample_b = True

    def _describe_noise(self, param):
        return ("variance = read_sd^2 / reads + (sample_cv * avgtail)^2\n"
                "read_sd = [%s]\n"
                "sample_cv = [%s]\n" % (
                    ", ".join("%f" % item for item in np.sqrt(param.read_variance)),
                    ", ".join("%f" % item for item in np.sqrt(param.sample_variance))))


class Model_t_patseq_v2_per_sample(Model_t_mixin, Model_normal_patseq_v2_per_sample):
    pass


class Model_normal_patseq_v3(_Patseq):
    _v3 = True
    _start = (1.0, 0.01)

    def _describe_noise(self, param):
        return "variance = (%f^2 / reads + %f^2) * tail^2\n" % (
            np.sqrt(param.read_variance), np.sqrt(param.sample_variance))


class Model_t_patseq_v3(Model_t_mixin, Model_normal_patseq_v3):
    pass


# Bless PATSEQ v2 (fitnoise.py:889-892)
Model_normal_patseq = Model_normal_patseq_v2
Model_t_patseq = Model_t_patseq_v2
Model_normal_patseq_per_sample = Model_normal_patseq_v2_per_sample
Model_t_patseq_per_sample = Model_t_patseq_v2_per_sample

# Names used by BASELINE.json / the legacy R package (backyard/old_fitnoise.R:599-614)
Model_t_standard = Model_t
Model_t_independent = Model_independent


class Model_normal_factors(Model_factors_mixin, Model_normal):
    pass


class Model_t_factors(Model_t_mixin, Model_factors_mixin, Model_normal):
    pass


class Model_independent_factors(Model_factors_mixin, Model_independent):
    pass
