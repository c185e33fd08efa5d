"""TEST-ONLY: host (g++) build of the kernels' per-gene mathematics and a stand-in evaluator.

``HostEvaluator`` has the interface of ``fitnoise_b200._native.Engine`` and runs the SAME
arithmetic as the CUDA kernels (fitnoise_b200/csrc/fn_gene.cuh, fn_special.cuh compiled for the
host, one gene at a time).  The CPU test-suite uses it to check (a) that arithmetic against the
oracle without a GPU and (b) the host-side logic of the package -- model classes, optimiser loop,
sharding and collectives over gloo.  It is injected through a test seam
(``model._engine_factory``); the product never constructs one and has no CPU fallback.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
BUILD = os.path.join(HERE, "_build")
_LIBS = {}

D = ctypes.POINTER(ctypes.c_double)
U8 = ctypes.POINTER(ctypes.c_uint8)
I32 = ctypes.POINTER(ctypes.c_int)


def _load(name):
    if name in _LIBS:
        return _LIBS[name]
    os.makedirs(BUILD, exist_ok=True)
    src = os.path.join(HERE, name + ".cpp")
    out = os.path.join(BUILD, "lib%s.so" % name)
    deps = [src] + [os.path.join(ROOT, "fitnoise_b200", "csrc", f)
                    for f in ("fn_gene.cuh", "fn_special.cuh", "fn_host.hpp")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wno-unknown-pragmas", "-shared", "-fPIC",
                               "-o", out + ".tmp%d" % os.getpid(), src])
        os.replace(out + ".tmp%d" % os.getpid(), out)
    _LIBS[name] = ctypes.CDLL(out)
    return _LIBS[name]


def special():
    return _load("special_host")


def gene():
    return _load("gene_host")


def _dp(a):
    return a.ctypes.data_as(D) if a is not None else None


def _vec(fn, *arrays):
    arrays = [np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float64), np.shape(arrays[0])))
              for a in arrays]
    out = np.empty_like(arrays[0])
    getattr(special(), fn)(out.size, *[_dp(a) for a in arrays], _dp(out))
    return out


def digamma(x):
    return _vec("hc_digamma", x)


def lgamma_digamma(x):
    return _vec("hc_lgamma2", x), _vec("hc_digamma2", x)


def f_sf(f, dfn, dfd):
    return _vec("hc_f_sf", f, dfn, dfd)


def chi2_sf(q, df):
    return _vec("hc_chi2_sf", q, df)


class _Info(object):
    pass


class HostEvaluator(object):
    """Same methods as _native.Engine, computed by the host build of the kernel mathematics."""

    def __init__(self):
        self.lib = gene()

    def close(self):
        pass

    def upload(self, y, aux, family, controls, design_noise, design_ctrl=None):
        self.y = np.ascontiguousarray(y, dtype=np.float64)
        self.n, self.m = self.y.shape
        self.aux = None if aux is None else np.ascontiguousarray(aux, dtype=np.float64)
        self.family = family
        self.controls = None if controls is None else np.ascontiguousarray(controls, dtype=np.uint8)
        self.dn = np.ascontiguousarray(design_noise, dtype=np.float64)
        self.dc = np.ascontiguousarray(np.ones((self.m, 1)) if design_ctrl is None else design_ctrl,
                                       dtype=np.float64)
        self.F = np.zeros((self.m, max(family.nf, 1)))
        theta0 = np.ones(max(family.theta_length(), 1))
        r = self._eval(theta0)
        info = _Info()
        info.row_variance_mean = r["rowvar"]
        info.n_eligible = int(r["eligible"].sum())
        info.n_bad_counts = r["bad"]
        info.n_variance_rows = int(np.isfinite(self.y).any(axis=1).sum())
        info.cst_sum = 0.0
        return info

    def _eval(self, theta):
        theta = np.ascontiguousarray(theta, dtype=np.float64).reshape(-1)
        n, P = self.n, self.family.theta_length()
        score, d2, avg = np.empty(n), np.empty(n), np.empty(n)
        pdf, elig = np.empty(n, np.int32), np.empty(n, np.uint8)
        val, rv, bad = ctypes.c_double(), ctypes.c_double(), ctypes.c_long()
        grad = np.zeros(max(P, 1))
        F = np.ascontiguousarray(self.F, dtype=np.float64)
        grad_f = np.zeros_like(F)
        rc = self.lib.hc_eval(
            ctypes.byref(self.family), ctypes.c_long(n), self.m, _dp(self.y), _dp(self.aux),
            self.controls.ctypes.data_as(U8) if self.controls is not None else None,
            self.dn.shape[1], _dp(self.dn), self.dc.shape[1], _dp(self.dc), _dp(theta),
            _dp(score), _dp(d2), pdf.ctypes.data_as(I32), _dp(avg), elig.ctypes.data_as(U8),
            ctypes.byref(val), _dp(grad), ctypes.byref(rv), ctypes.byref(bad), _dp(F), _dp(grad_f))
        if rc != 0:
            raise RuntimeError("hc_eval failed with %d" % rc)
        return dict(score=score, d2=d2, p=pdf, averages=avg, eligible=elig, value=val.value,
                    grad=grad[:P], rowvar=rv.value, bad=bad.value, grad_f=grad_f)

    def set_factors(self, F):
        self.F = np.ascontiguousarray(F, dtype=np.float64).reshape(self.m, -1)

    def nll_grad_factors(self, theta, F):
        self.set_factors(F)
        r = self._eval(theta)
        used = int(np.isfinite(r["score"]).sum())
        return r["value"], r["grad"], r["grad_f"], used, 0

    def nll_grad(self, theta):
        r = self._eval(theta)
        used = int(np.isfinite(r["score"]).sum())
        return r["value"], r["grad"], used, 0

    def per_gene(self, theta):
        r = self._eval(theta)
        return r["score"], r["d2"], r["p"]

    def noise_stats(self, theta):
        r = self._eval(theta)
        fam = self.family
        ok = np.isfinite(r["d2"]) & (r["p"] > 0)
        out = np.full(self.n, np.nan)
        p = r["p"][ok].astype(float)
        if fam.dist == 0:
            out[ok] = chi2_sf(r["d2"][ok], p) if ok.any() else []
        else:
            v = float(np.asarray(theta).reshape(-1)[0]) if fam.dist == 1 else fam.fixed_df
            out[ok] = f_sf(r["d2"][ok] / p, p, np.full(len(p), v)) if ok.any() else []
        return out, r["averages"]

    def coef(self, theta, design, want_covar=True):
        theta = np.ascontiguousarray(theta, dtype=np.float64).reshape(-1)
        design = np.ascontiguousarray(design, dtype=np.float64)
        c = design.shape[1]
        coef, covar, dfp = np.empty((self.n, c)), np.empty((self.n, c, c)), np.empty(self.n)
        rc = self.lib.hc_coef(ctypes.byref(self.family), ctypes.c_long(self.n), self.m, _dp(self.y),
                              _dp(self.aux), c, _dp(design), _dp(theta), _dp(coef), _dp(covar), _dp(dfp),
                              _dp(np.ascontiguousarray(self.F, dtype=np.float64)))
        if rc != 0:
            raise RuntimeError("hc_coef failed with %d" % rc)
        self._coef = (coef, covar, dfp, int(self.family.dist != 0))
        if not want_covar:
            return coef, None, None
        return coef, covar, dfp

    def coef_fetch(self):
        return self._coef[1], self._coef[2]

    def test_bh(self, contrasts):
        con, _, p = self.test(contrasts)
        q, self.last_order = self.bh_ranked(p)
        return con, p, q

    def bh_ranked(self, p):
        return self.bh(p), np.argsort(np.asarray(p, dtype=np.float64), kind="stable")

    def test(self, contrasts, want_covar=False):
        coef, covar, dfp, is_t = self._coef
        contrasts = np.ascontiguousarray(contrasts, dtype=np.float64)
        c, kc = contrasts.shape
        con, ccov, p = np.empty((self.n, kc)), np.empty((self.n, kc, kc)), np.empty(self.n)
        self.lib.hc_test(ctypes.c_long(self.n), c, kc, _dp(coef), _dp(covar), _dp(dfp), is_t,
                         _dp(contrasts), _dp(con), _dp(ccov), _dp(p))
        return con, ccov, p

    def bh(self, p):
        # plain numpy statement of p_adjust.py:3-21 (the device radix-sort path is covered by -m gpu)
        p = np.array(p, dtype="float64")
        p[np.isnan(p)] = 1.0
        n = len(p)
        order = np.argsort(p)[::-1]
        qs = np.minimum(1.0, np.minimum.accumulate(p[order] * n / np.arange(n, 0, -1)))
        q = np.empty(n)
        q[order] = qs
        return q

    # fitnoise.transform stand-in: the kernel's sums (csrc/fn_transform.cuh) in plain numpy
    def transform_upload(self, x, design):
        self.tx = np.asarray(x, dtype=np.float64)
        q, _ = np.linalg.qr(np.asarray(design, dtype=np.float64))
        self.tq = q

    def _transform_parts(self, order, a):
        a = np.asarray(a, dtype=np.float64).reshape(order, -1)
        poly = 1.0
        for i in range(order):
            poly = poly * self.tx + a[i][None, :]
        scale = 1.0 / (order * np.log(2.0))
        return a, poly, np.log(poly) * scale, scale

    def transform_sums(self, order, a):
        a, poly, y, scale = self._transform_parts(order, a)
        t = y @ self.tq
        proj = t @ self.tq.T
        A, B, D = [], [], []
        for i in range(order):
            w = scale * self.tx ** (order - 1 - i) / poly
            A.append((y * w).sum(axis=0)); B.append(w.sum(axis=0)); D.append((proj * w).sum(axis=0))
        return np.concatenate([[(y * y).sum(), (t * t).sum()], y.sum(axis=0)] + A + B + D)

    def transform_apply(self, order, a, col_mean):
        a, poly, y, scale = self._transform_parts(order, a)
        return y - np.asarray(col_mean)[None, :]

    def weights(self, theta):
        theta = np.asarray(theta, dtype=np.float64).reshape(-1)
        fam = self.family
        off = 1 if fam.dist == 1 else 0
        a, b = theta[off:off + fam.na], theta[off + fam.na:off + fam.na + fam.nb]
        with np.errstate(divide="ignore"):
            if fam.kind == 2:
                rc = 1.0 / np.maximum(self.aux, 1.0)
                y0 = np.where(np.isfinite(self.y), self.y, 0.0)
                num = (y0 * self.aux).sum(axis=1)
                if not fam.v3:
                    num = np.maximum(1.0, num)
                gs = (num / np.maximum(1.0, self.aux.sum(axis=1)))[:, None] ** 2
                U = rc * gs if fam.v3 else rc
                T = y0 * y0 if fam.tail_own else np.broadcast_to(gs, self.y.shape)
                return 1.0 / (a * U + b * T)
            w = np.ones_like(self.y) if self.aux is None else self.aux
            if fam.nf:
                return 1.0 / ((a if fam.na else 1.0) / w + (self.F ** 2).sum(axis=1)[None, :])
            return w / (a if fam.na else 1.0)
