"""Host-side value objects for ONE gene's distribution: what ``noise_dists[i]``, ``coef_dists[i]``
and ``contrast_dists[i]`` hand back to the user (reference: fitnoise/distributions.py).

The batched work -- log-densities, marginals, transforms and conditionals of every gene for every
theta -- is done by the CUDA kernels (csrc/fn_gene.cuh); it never goes through these classes.
They exist so that code written against the reference's result attributes (``.mean``, ``.covar``,
``.df``, ``.transformed(A)``, ``.p_value(x)``, ...) keeps working on single rows pulled out of a
fit.  Same public methods as the reference's Mvnormal / Mvt.
"""
import numpy as np
import scipy.special
import scipy.stats


class _Elliptical(object):
    """Shared algebra of the two families; ``df`` is None for the normal."""
    df = None

    def __init__(self, mean, covar):
        self.mean = np.asarray(mean, dtype="float64").reshape(-1)
        self.covar = np.asarray(covar, dtype="float64")
        assert self.covar.ndim == 2

    def _like(self, mean, covar, df=None):
        raise NotImplementedError

    def __bool__(self):
        return True

    __nonzero__ = __bool__

    @property
    def good(self):
        """Per-coordinate usability: finite mean and a finite covariance row (reference :31-36)."""
        return np.isfinite(self.mean) & np.all(np.isfinite(self.covar), axis=1)

    def _distance2(self, x):
        off = np.asarray(x, dtype="float64") - self.mean
        return float(off @ np.linalg.solve(self.covar, off)) if len(off) else 0.0

    def _logdet(self):
        if self.covar.shape[0] == 0:
            return 0.0
        sign, logdet = np.linalg.slogdet(self.covar)
        return logdet if sign > 0 else np.nan

    def transformed(self, A):
        A = np.asarray(A, dtype="float64")
        return self._like(A @ self.mean, A @ self.covar @ A.T)

    def shifted(self, x):
        return self._like(self.mean + np.asarray(x, dtype="float64"), self.covar)

    def plus_covar(self, A):
        return self._like(self.mean, self.covar + np.asarray(A, dtype="float64"))

    def marginal(self, i):
        i = np.asarray(i, dtype="int64")
        return self._like(self.mean[i], self.covar[np.ix_(i, i)])

    def _conditional_parts(self, i1, i2, x2):
        i1 = np.asarray(i1, dtype="int64")
        i2 = np.asarray(i2, dtype="int64")
        off2 = np.asarray(x2, dtype="float64") - self.mean[i2]
        c12 = self.covar[np.ix_(i1, i2)]
        c22 = self.covar[np.ix_(i2, i2)]
        gain = np.linalg.solve(c22.T, c12.T).T if len(i2) else np.zeros((len(i1), 0))
        mean = self.mean[i1] + gain @ off2
        covar = self.covar[np.ix_(i1, i1)] - gain @ c12.T
        d2 = float(off2 @ np.linalg.solve(c22, off2)) if len(i2) else 0.0
        return mean, covar, d2, len(i2)

    def random(self):
        L = np.linalg.cholesky(self.covar)
        z = L @ np.random.normal(size=len(self.mean))
        if self.df is not None:
            z = z * np.sqrt(self.df / np.random.chisquare(self.df))
        return self.mean + z


class Mvnormal(_Elliptical):
    def _like(self, mean, covar, df=None):
        return Mvnormal(mean, covar)

    def __repr__(self):
        return "Mvnormal(\n   mean=%r,\n  covar=%r\n)" % (self.mean, self.covar)

    def log_density(self, x):
        n = len(self.mean)
        return -0.5 * (np.log(2 * np.pi) * n + self._logdet() + self._distance2(x))

    def p_value(self, x):
        return float(scipy.stats.chi2.sf(self._distance2(x), df=len(self.mean)))

    def conditional(self, i1, i2, x2):
        mean, covar, _, _ = self._conditional_parts(i1, i2, x2)
        return Mvnormal(mean, covar)


class Mvt(_Elliptical):
    def __init__(self, mean, covar, df):
        _Elliptical.__init__(self, mean, covar)
        self.df = float(df)

    def _like(self, mean, covar, df=None):
        return Mvt(mean, covar, self.df if df is None else df)

    def __repr__(self):
        return "Mvt(\n     df=%r,\n   mean=%r,\n  covar=%r\n)" % (self.df, self.mean, self.covar)

    def log_density(self, x):
        p, v = len(self.mean), self.df
        return (scipy.special.gammaln(0.5 * (v + p)) - scipy.special.gammaln(0.5 * v)
                - 0.5 * p * np.log(np.pi * v) - 0.5 * self._logdet()
                - 0.5 * (v + p) * np.log1p(self._distance2(x) / v))

    def p_value(self, x):
        p = len(self.mean)
        return float(scipy.stats.f.sf(self._distance2(x) / p, dfn=p, dfd=self.df))

    def conditional(self, i1, i2, x2):
        mean, covar, d2, p2 = self._conditional_parts(i1, i2, x2)
        return Mvt(mean, covar * ((self.df + d2) / (self.df + p2)), self.df + p2)
