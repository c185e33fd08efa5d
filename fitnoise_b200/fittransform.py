"""Optimal moderated log transformation of count data (reference: fitnoise/fittransform.py).

    y = log2(P_j(x)) / order,   P_j(x) = x^order + a_0j x^(order-1) + ... + a_(order-1)j

with one monic polynomial per sample, chosen to minimise ln(var_h1) - ln(var_h0): the variance of
the column-centred transformed data left after projecting out the design, relative to its total
variance (fittransform.py:103-121).  The reference builds this objective in Theano and differentiates
it automatically; here one CUDA kernel (csrc/fn_transform.cuh) streams the count matrix once per
evaluation and returns the handful of sums from which objective and gradient follow in closed form.
The L-BFGS-B restart loop (fittransform.py:140-156) stays on the host.  No CPU fallback.
"""
import os

import numpy as np
import scipy.optimize

from . import _native, parallel
from .env import Withable, as_matrix, as_vector


def _combine(sums, n, m, order, projector):
    """Objective and gradient from the kernel's sums (fn_transform.cuh), summed over all ranks."""
    syy, stt = sums[0], sums[1]
    s = sums[2:2 + m]
    blocks = sums[2 + m:].reshape(3, order, m)
    A, B, D = blocks[0], blocks[1], blocks[2]
    mu = s / n
    pmu = projector @ mu                        # Q Q' mu
    v0 = syy - n * float(mu @ mu)
    v1 = v0 - (stt - n * float(mu @ pmu))
    cost = np.log(v1) - np.log(v0)
    grad = (2.0 / v1 - 2.0 / v0) * (A - mu[None, :] * B) - (2.0 / v1) * (D - pmu[None, :] * B)
    return cost, grad.reshape(-1), mu


class Transform(Withable):
    def __repr__(self):
        result = self.__class__.__name__
        if hasattr(self, "y"):
            result += "\n" + self.optimization_result.message
            result += "\nTransform parameters: " + repr(self.transform)
        return result

    def _configured(self):
        return self

    def fit(self, x, design=None, verbose=False):
        x = as_matrix(x)
        n, m = x.shape
        if design is None:
            design = np.ones((m, 1))
        design = as_matrix(design)
        q, r = np.linalg.qr(design, mode="complete")
        q_design = q[:, :design.shape[1]]
        q_null = q[:, design.shape[1]:]
        result = self._with(x=x, design=design, null=q_null)._configured()
        if verbose:
            print("Initial guess:", result._param_initial)
        initial = as_vector(result._param_initial)
        bounds = result._param_bounds
        order = result.order
        projector = q_design @ q_design.T

        rank, world = parallel.world()
        lo, hi = (int(b) for b in parallel.shard_bounds(n, world)[rank:rank + 2])
        factory = getattr(self, "_engine_factory", None)
        engine = factory() if factory is not None else _native.acquire_engine(int(os.environ.get("LOCAL_RANK", "0")))
        try:
            engine.transform_upload(x[lo:hi], design)

            def score(pack):
                sums = parallel.allreduce_sum(engine.transform_sums(order, pack))
                cost, grad, _ = _combine(sums, n, m, order, projector)
                return cost, grad

            a = initial
            last = None
            for i in range(100):                                   # fittransform.py:140-156
                fit = scipy.optimize.minimize(score, a, method="L-BFGS-B", jac=True, bounds=bounds)
                a = fit.x
                if verbose:
                    print("Score:", fit.fun)
                if last is not None and last - fit.fun <= 1e-6:
                    break
                last = fit.fun
            if verbose:
                print("Optimized:", list(fit.x))
                print(fit.message)
            pack = fit.x.copy()
            sums = parallel.allreduce_sum(engine.transform_sums(order, pack))
            _, _, col_mean = _combine(sums, n, m, order, projector)
            y = parallel.allgather_rows(engine.transform_apply(order, pack, col_mean), n)
        finally:
            if factory is None:
                _native.release_engine(engine)
        transform = result._unpack_transform(m, pack)
        return result._with(optimization_result=fit, transform=transform, y=y)


class Transform_polynomial(Transform):
    def __init__(self, order):
        assert order > 0
        self.order = order

    def _unpack_transform(self, m, pack):
        vecs = []
        for i in range(self.order):
            vec, pack = pack[:m], pack[m:]
            vecs.append(vec)
        return vecs

    def _configured(self):
        result = super(Transform_polynomial, self)._configured()
        m = result.x.shape[1]
        param_initial = [0.0] * (self.order - 1) * m + [1.0] * m
        param_bounds = [(0.0, None)] * (self.order - 1) * m + [(1e-10, None)] * m
        assert len(param_initial) == len(param_bounds)
        return result._with(_param_initial=param_initial, _param_bounds=param_bounds)


def transform(x, design=None, order=2, verbose=False):
    return Transform_polynomial(order).fit(x, design=design, verbose=verbose)
