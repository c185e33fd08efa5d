"""ctypes binding of libfitnoise_b200.so (include/fitnoise_b200.h).

This is the only way the package computes anything: there is no CPU fallback.  Importing the
package works without a GPU (so that host-side logic can be tested), but creating an Engine
raises unless the CUDA library loads and a CUDA device is present.
"""
import ctypes
import os

import numpy as np

from . import build as _build

_LIB = None

c_double_p = ctypes.POINTER(ctypes.c_double)
c_uint8_p = ctypes.POINTER(ctypes.c_uint8)
c_int32_p = ctypes.POINTER(ctypes.c_int32)

KIND_SCALE1, KIND_SCALE_PS, KIND_PATSEQ = 0, 1, 2
DIST_NORMAL, DIST_T, DIST_T_FIXED = 0, 1, 2

EXPORTS = [
    "fn_last_error", "fn_device_count", "fn_version", "fn_create", "fn_destroy", "fn_synchronize",
    "fn_upload", "fn_upload_dev", "fn_nll_grad", "fn_nll_grad_dev", "fn_per_gene", "fn_noise_stats",
    "fn_coef", "fn_test", "fn_bh", "fn_bh_dev", "fn_weights", "fn_set_tuning", "fn_launch_count",
    "fn_set_timing", "fn_last_kernel_ms", "fn_comm_mailbox", "fn_comm_connect", "fn_comm_disconnect",
    "fn_test_bh", "fn_coef_fetch", "fn_bh_ranked", "fn_nll_grad_factors", "fn_set_factors", "fn_time_nll_grad",
    "fn_transform_upload", "fn_transform_sums", "fn_transform_apply",
    "fn_resident_begin", "fn_resident_end", "fn_resident_stats", "fn_source_hash", "fn_l2_flush",
    "fn_result_dev", "fn_fetch", "fn_bh_ranked_dev", "fn_pinned_alloc", "fn_pinned_free",
]

# fn_result_dev selectors (include/fitnoise_b200.h)
(RES_SCORE, RES_D2, RES_PDF, RES_NOISE_P, RES_AVERAGES, RES_COEF, RES_COVAR, RES_DFPOST, RES_CONTRASTED,
 RES_P, RES_Q, RES_ORDER) = range(12)
RES_DTYPE = {RES_PDF: np.int32, RES_ORDER: np.uint32}


class Family(ctypes.Structure):
    """fn_family (include/fitnoise_b200.h)."""
    _fields_ = [("kind", ctypes.c_int32), ("na", ctypes.c_int32), ("nb", ctypes.c_int32),
                ("tail_own", ctypes.c_int32), ("v3", ctypes.c_int32), ("dist", ctypes.c_int32),
                ("nf", ctypes.c_int32), ("fixed_df", ctypes.c_double)]

    def __init__(self, kind=0, na=0, nb=0, tail_own=0, v3=0, dist=0, fixed_df=0.0, nf=0):
        super(Family, self).__init__(kind, na, nb, tail_own, v3, dist, nf, fixed_df)

    def theta_length(self):
        return (1 if self.dist == DIST_T else 0) + self.na + self.nb


class PrepInfo(ctypes.Structure):
    _fields_ = [("row_variance_mean", ctypes.c_double), ("cst_sum", ctypes.c_double),
                ("n_eligible", ctypes.c_int64), ("n_bad_counts", ctypes.c_int64),
                ("n_variance_rows", ctypes.c_int64)]


class NativeError(RuntimeError):
    pass


def library_path():
    return _build.LIB_PATH


_PYCALL = [False, None]


def _pycall():
    """fn_pycall.nll_grad, or None when the optional module is not there (FN_NO_PYCALL=1 disables it)."""
    if _PYCALL[0]:
        return _PYCALL[1]
    _PYCALL[0] = True
    path = _build.pycall_path()
    if os.environ.get("FN_NO_PYCALL") or not os.path.exists(path):
        return None
    try:
        import importlib.machinery
        import importlib.util
        loader = importlib.machinery.ExtensionFileLoader("_fn_pycall", path)
        spec = importlib.util.spec_from_loader("_fn_pycall", loader)
        module = importlib.util.module_from_spec(spec)
        loader.exec_module(module)
        _PYCALL[1] = module.nll_grad
    except Exception:        # built for another interpreter: ctypes does the same job
        _PYCALL[1] = None
    return _PYCALL[1]


def load_library():
    """Load the CUDA library, building it first when it is missing or was built from other sources
    than the ones in this tree (content hash, fitnoise_b200.build.source_hash).  Raises when it
    cannot be built -- never substitutes another implementation, never loads a stale binary."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if _build.needs_build():
        try:
            _build.build()
        except Exception as exc:  # no nvcc on the box
            raise NativeError("libfitnoise_b200.so is %s and cannot be built here (%s). "
                              "Run `python __graft_entry__.py` / fitnoise_b200.build.build(); "
                              "there is no CPU fallback." % (
                                  "out of date" if os.path.exists(path) else "not built", exc))
    lib = ctypes.CDLL(path)
    lib.fn_last_error.restype = ctypes.c_char_p
    lib.fn_version.restype = ctypes.c_char_p
    lib.fn_source_hash.restype = ctypes.c_char_p
    if lib.fn_source_hash().decode() != _build.source_hash() and not os.environ.get("FN_SKIP_SOURCE_HASH"):
        raise NativeError("libfitnoise_b200.so was built from different sources than fitnoise_b200/csrc "
                          "(ABI may not match the Python declarations); rebuild it")
    lib.fn_resident_begin.argtypes = [ctypes.c_void_p]
    lib.fn_resident_end.argtypes = [ctypes.c_void_p]
    lib.fn_l2_flush.argtypes = [ctypes.c_void_p]
    lib.fn_result_dev.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.POINTER(ctypes.c_void_p),
                                  ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int32)]
    lib.fn_fetch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
    lib.fn_bh_ranked_dev.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
                                     ctypes.POINTER(ctypes.c_void_p)]
    lib.fn_pinned_alloc.argtypes = [ctypes.c_int64, ctypes.POINTER(ctypes.c_void_p)]
    lib.fn_pinned_free.argtypes = [ctypes.c_void_p]
    lib.fn_resident_stats.restype = ctypes.c_int64
    lib.fn_resident_stats.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64)]
    lib.fn_launch_count.restype = ctypes.c_int64
    lib.fn_launch_count.argtypes = [ctypes.c_void_p]
    lib.fn_last_kernel_ms.restype = ctypes.c_double
    lib.fn_last_kernel_ms.argtypes = [ctypes.c_void_p]
    lib.fn_create.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]
    lib.fn_destroy.argtypes = [ctypes.c_void_p]
    lib.fn_synchronize.argtypes = [ctypes.c_void_p]
    upload_args = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p,
                   ctypes.POINTER(Family), ctypes.c_void_p, ctypes.c_int32, c_double_p,
                   ctypes.c_int32, c_double_p, ctypes.POINTER(PrepInfo)]
    lib.fn_upload.argtypes = upload_args
    lib.fn_upload_dev.argtypes = upload_args
    lib.fn_nll_grad.argtypes = [ctypes.c_void_p, c_double_p, ctypes.c_int32, c_double_p, c_double_p,
                                ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64)]
    lib.fn_nll_grad_factors.argtypes = [ctypes.c_void_p, c_double_p, ctypes.c_int32, c_double_p, c_double_p,
                                        c_double_p, c_double_p, ctypes.POINTER(ctypes.c_int64),
                                        ctypes.POINTER(ctypes.c_int64)]
    lib.fn_set_factors.argtypes = [ctypes.c_void_p, c_double_p]
    lib.fn_transform_upload.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p,
                                        ctypes.c_int32, c_double_p]
    lib.fn_transform_sums.argtypes = [ctypes.c_void_p, ctypes.c_int32, c_double_p, c_double_p]
    lib.fn_transform_apply.argtypes = [ctypes.c_void_p, ctypes.c_int32, c_double_p, c_double_p, c_double_p]
    lib.fn_time_nll_grad.argtypes = [ctypes.c_void_p, c_double_p, ctypes.c_int32, ctypes.c_int32, c_double_p]
    lib.fn_nll_grad_dev.argtypes = [ctypes.c_void_p, c_double_p, ctypes.c_int32, ctypes.c_void_p]
    lib.fn_per_gene.argtypes = [ctypes.c_void_p, c_double_p, ctypes.c_int32, c_double_p, c_double_p, c_int32_p]
    lib.fn_noise_stats.argtypes = [ctypes.c_void_p, c_double_p, ctypes.c_int32, c_double_p, c_double_p]
    lib.fn_coef.argtypes = [ctypes.c_void_p, c_double_p, ctypes.c_int32, ctypes.c_int32, c_double_p,
                            c_double_p, c_double_p, c_double_p]
    lib.fn_test.argtypes = [ctypes.c_void_p, ctypes.c_int32, c_double_p, c_double_p, c_double_p, c_double_p]
    lib.fn_test_bh.argtypes = [ctypes.c_void_p, ctypes.c_int32, c_double_p, c_double_p, c_double_p, c_double_p,
                               ctypes.c_void_p]
    lib.fn_bh_ranked.argtypes = [ctypes.c_void_p, ctypes.c_int64, c_double_p, c_double_p, ctypes.c_void_p]
    lib.fn_coef_fetch.argtypes = [ctypes.c_void_p, c_double_p, c_double_p, c_double_p]
    lib.fn_bh.argtypes = [ctypes.c_void_p, ctypes.c_int64, c_double_p, c_double_p]
    lib.fn_bh_dev.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
    lib.fn_weights.argtypes = [ctypes.c_void_p, c_double_p, ctypes.c_int32, c_double_p]
    lib.fn_set_tuning.argtypes = [ctypes.c_void_p] + [ctypes.c_int32] * 4
    lib.fn_set_timing.argtypes = [ctypes.c_void_p, ctypes.c_int32]
    lib.fn_comm_mailbox.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p]
    lib.fn_comm_connect.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p]
    lib.fn_comm_disconnect.argtypes = [ctypes.c_void_p]
    _LIB = lib
    return lib


def device_count():
    return int(load_library().fn_device_count())


def _dp(a):
    return a.ctypes.data_as(c_double_p) if a is not None else None


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        assert a.shape == tuple(shape), "expected shape %s, got %s" % (tuple(shape), a.shape)
    return a


class _PinnedBlock(object):
    """Owner of one block of pooled pinned host memory; numpy arrays made from it keep it alive and
    the block goes back to the library's pool when the last of them dies."""

    def __init__(self, lib, nbytes):
        ptr = ctypes.c_void_p()
        if lib.fn_pinned_alloc(int(nbytes), ctypes.byref(ptr)) != 0:
            raise NativeError(lib.fn_last_error().decode())
        self._lib, self.ptr, self.nbytes = lib, ptr.value, int(nbytes)

    @property
    def __array_interface__(self):
        return {"shape": (self.nbytes,), "typestr": "|u1", "data": (self.ptr, False), "version": 3}

    def __del__(self):
        try:
            self._lib.fn_pinned_free(ctypes.c_void_p(self.ptr))
        except Exception:
            pass


def pinned_empty(shape, dtype=np.float64):
    """numpy.empty in pinned host memory (a DMA target: device-to-host copies run at PCIe rate and
    touch no fresh pages).  Falls back to plain memory for empty arrays."""
    shape = tuple(int(x) for x in (shape if isinstance(shape, (tuple, list)) else (shape,)))
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
    if nbytes == 0:
        return np.empty(shape, dtype)
    block = _PinnedBlock(load_library(), nbytes)
    return np.asarray(block).view(dtype).reshape(shape)


class Engine(object):
    """One GPU's shard of the genes (an fn_handle)."""

    def __init__(self, device=0, stream=None):
        self._lib = load_library()
        self._h = ctypes.c_void_p()
        rc = self._lib.fn_create(int(device), ctypes.c_void_p(stream) if stream else None, ctypes.byref(self._h))
        if rc != 0:
            self._h = None
            raise NativeError(self._lib.fn_last_error().decode())
        self.device = int(device)
        self.n = self.m = 0
        self.family = None
        self.info = None
        # plain addresses for the CPython fast call (csrc/fn_pycall.c)
        self._h_addr = int(self._h.value)
        self._nll_addr = int(ctypes.cast(self._lib.fn_nll_grad, ctypes.c_void_p).value)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.fn_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise NativeError(self._lib.fn_last_error().decode())

    # ---- data -----------------------------------------------------------------------------
    def upload(self, y, aux, family, controls, design_noise, design_ctrl=None):
        y = _f64(y)
        n, m = y.shape
        aux = None if aux is None else _f64(aux, (n, m))
        controls = None if controls is None else np.ascontiguousarray(controls, dtype=np.uint8)
        dn = _f64(design_noise)
        assert dn.ndim == 2 and dn.shape[0] == m, "design must have one row per sample"
        dc = None if design_ctrl is None else _f64(design_ctrl)
        info = PrepInfo()
        self._keep = (y, aux, controls)
        self._check(self._lib.fn_upload(
            self._h, n, m, y.ctypes.data, aux.ctypes.data if aux is not None else None,
            ctypes.byref(family), controls.ctypes.data if controls is not None else None,
            dn.shape[1], _dp(dn), dc.shape[1] if dc is not None else 0, _dp(dc), ctypes.byref(info)))
        self._keep = None
        self.n, self.m, self.family, self.info = n, m, family, info
        return info

    def upload_dev(self, n, m, y_ptr, aux_ptr, family, controls_ptr, design_noise, design_ctrl=None):
        """Device-resident inputs (raw CUDA pointers, e.g. torch tensor .data_ptr())."""
        dn = _f64(design_noise)
        dc = None if design_ctrl is None else _f64(design_ctrl)
        info = PrepInfo()
        self._check(self._lib.fn_upload_dev(
            self._h, n, m, y_ptr, aux_ptr or None, ctypes.byref(family), controls_ptr or None,
            dn.shape[1], _dp(dn), dc.shape[1] if dc is not None else 0, _dp(dc), ctypes.byref(info)))
        self.n, self.m, self.family, self.info = n, m, family, info
        return info

    # ---- likelihood -----------------------------------------------------------------------
    def nll_grad(self, theta):
        """One REML objective + gradient evaluation over this engine's genes (all ranks' genes when
        the fused exchange is connected): (value, grad, n_used, n_bad).  Called once per optimiser
        step: through the thin CPython entry (csrc/fn_pycall.c: the same fn_nll_grad, reached by its
        address with buffer-protocol arguments, ~0.3 us of host time) when that module was built, else
        through ctypes with scaffolding that is built once and reused (2-3 us)."""
        fast = _pycall() if self._h else None
        if fast is not None and type(theta) is np.ndarray and theta.dtype == np.float64 and theta.flags.c_contiguous:
            grad = np.empty(theta.shape[0])
            rc, value, used, bad = fast(self._nll_addr, self._h_addr, theta, grad)
            if rc != 0:
                self._check(rc)
            return value, grad, used, bad
        call = self.__dict__.get("_nll_call")
        P = len(theta)
        if call is None or call[0] != P:
            value, used, bad = ctypes.c_double(), ctypes.c_int64(), ctypes.c_int64()
            tbuf, grad = np.zeros(max(P, 1)), np.zeros(max(P, 1))
            # pointers to persistent buffers: no per-call ctypes object construction
            call = self._nll_call = (P, value, used, bad, grad, tbuf, ctypes.cast(tbuf.ctypes.data, c_double_p),
                                     ctypes.c_int32(P), ctypes.byref(value), ctypes.cast(grad.ctypes.data, c_double_p),
                                     ctypes.byref(used), ctypes.byref(bad), self._lib.fn_nll_grad)
        _, value, used, bad, grad, tbuf, ptheta, cP, pvalue, pgrad, pused, pbad, fn = call
        if P:
            tbuf[:] = theta
        rc = fn(self._h, ptheta, cP, pvalue, pgrad, pused, pbad)
        if rc != 0:
            self._check(rc)
        return value.value, grad[:P].copy(), used.value, bad.value

    # ---- resident evaluation kernel -----------------------------------------------------------
    def resident_begin(self):
        """Keep ONE evaluation kernel on the GPU until resident_end(): nll_grad then costs no kernel
        launch (include/fitnoise_b200.h, fn_resident_begin)."""
        self._check(self._lib.fn_resident_begin(self._h))

    def resident_end(self):
        if getattr(self, "_h", None):
            self._check(self._lib.fn_resident_end(self._h))

    def l2_flush(self):
        """Evict the inputs from the L2 (benchmarking aid); works in the resident phase too."""
        self._check(self._lib.fn_l2_flush(self._h))

    def resident_stats(self):
        """(resident now?, resident kernels launched, evaluations they executed)"""
        launches, evals = ctypes.c_int64(), ctypes.c_int64()
        on = self._lib.fn_resident_stats(self._h, ctypes.byref(launches), ctypes.byref(evals))
        return bool(on), launches.value, evals.value

    def nll_grad_factors(self, theta, factors):
        """Factor models: (value, grad of [df, variance], d/dF (m x nf), n_used, n_bad)."""
        theta = _f64(theta).reshape(-1)
        factors = _f64(factors).reshape(self.m, -1)
        P = len(theta)
        value = ctypes.c_double()
        grad, grad_f = np.zeros(max(P, 1)), np.zeros_like(factors)
        used, bad = ctypes.c_int64(), ctypes.c_int64()
        self._check(self._lib.fn_nll_grad_factors(self._h, _dp(theta), P, _dp(factors), ctypes.byref(value),
                                                  _dp(grad), _dp(grad_f), ctypes.byref(used), ctypes.byref(bad)))
        return value.value, grad[:P], grad_f, used.value, bad.value

    def set_factors(self, factors):
        factors = _f64(factors).reshape(self.m, -1)
        self._check(self._lib.fn_set_factors(self._h, _dp(factors)))

    def nll_grad_dev(self, theta, out_ptr):
        theta = _f64(theta).reshape(-1)
        self._check(self._lib.fn_nll_grad_dev(self._h, _dp(theta), len(theta), out_ptr))

    def per_gene(self, theta):
        theta = _f64(theta).reshape(-1)
        score, d2, p = pinned_empty(self.n), pinned_empty(self.n), pinned_empty(self.n, np.int32)
        self._check(self._lib.fn_per_gene(self._h, _dp(theta), len(theta), _dp(score), _dp(d2),
                                          p.ctypes.data_as(c_int32_p)))
        return score, d2, p

    def noise_stats(self, theta):
        theta = _f64(theta).reshape(-1)
        noise_p, averages = pinned_empty(self.n), pinned_empty(self.n)
        self._check(self._lib.fn_noise_stats(self._h, _dp(theta), len(theta), _dp(noise_p), _dp(averages)))
        return noise_p, averages

    def noise_stats_dev(self, theta):
        """noise_stats with the results left on the device (result_dev(RES_NOISE_P / RES_AVERAGES))."""
        theta = _f64(theta).reshape(-1)
        self._check(self._lib.fn_noise_stats(self._h, _dp(theta), len(theta), None, None))

    # ---- coefficients and tests -----------------------------------------------------------
    def coef(self, theta, design, want_covar=True):
        """GLS coefficients.  With want_covar=False the covariances and posterior df stay on the
        device (fetch them later with coef_fetch); returns (coef, covar|None, dfpost|None)."""
        theta = _f64(theta).reshape(-1)
        design = _f64(design)
        c = design.shape[1]
        coef = pinned_empty((self.n, c))
        covar = pinned_empty((self.n, c, c)) if want_covar else None
        dfpost = pinned_empty(self.n) if want_covar else None
        self._check(self._lib.fn_coef(self._h, _dp(theta), len(theta), c, _dp(design), _dp(coef),
                                      _dp(covar), _dp(dfpost)))
        self._coef_c = c
        return coef, covar, dfpost

    def coef_dev(self, theta, design):
        """coef() with every result left on the device (result_dev(RES_COEF / RES_COVAR / RES_DFPOST))."""
        theta = _f64(theta).reshape(-1)
        design = _f64(design)
        self._check(self._lib.fn_coef(self._h, _dp(theta), len(theta), design.shape[1], _dp(design), None, None, None))
        self._coef_c = design.shape[1]

    def test_dev(self, contrasts):
        """test() with the results left on the device (result_dev(RES_CONTRASTED / RES_P))."""
        contrasts = _f64(contrasts)
        self._check(self._lib.fn_test(self._h, contrasts.shape[1], _dp(contrasts), None, None, None))

    def result_dev(self, which):
        """(device pointer, rows, columns, dtype) of a result of the last call that computed it."""
        ptr, rows, cols = ctypes.c_void_p(), ctypes.c_int64(), ctypes.c_int32()
        self._check(self._lib.fn_result_dev(self._h, which, ctypes.byref(ptr), ctypes.byref(rows), ctypes.byref(cols)))
        return ptr.value, rows.value, cols.value, np.dtype(RES_DTYPE.get(which, np.float64))

    def fetch(self, out, dev_ptr):
        """Copy out.nbytes from device memory at dev_ptr into the (contiguous) host array `out`."""
        assert out.flags["C_CONTIGUOUS"]
        if out.nbytes:
            self._check(self._lib.fn_fetch(self._h, out.ctypes.data, dev_ptr, out.nbytes))
        return out

    def bh_ranked_dev(self, n, p_ptr, q_ptr):
        """BH over n device-resident p-values; returns the device pointer of the order (uint32 x n)."""
        order = ctypes.c_void_p()
        self._check(self._lib.fn_bh_ranked_dev(self._h, n, p_ptr, q_ptr, ctypes.byref(order)))
        return order.value

    def coef_fetch(self):
        """Covariances (n x c x c) and posterior df (n) of the last coef()."""
        c = self._coef_c
        covar, dfpost = pinned_empty((self.n, c, c)), pinned_empty(self.n)
        self._check(self._lib.fn_coef_fetch(self._h, None, _dp(covar), _dp(dfpost)))
        return covar, dfpost

    def test_bh(self, contrasts):
        """test() and the BH adjustment of its p-values in one call (p stays on the device)."""
        contrasts = _f64(contrasts)
        kc = contrasts.shape[1]
        contrasted, p, q = pinned_empty((self.n, kc)), pinned_empty(self.n), pinned_empty(self.n)
        order = pinned_empty(self.n, np.uint32)
        self._check(self._lib.fn_test_bh(self._h, kc, _dp(contrasts), _dp(contrasted), _dp(p), _dp(q),
                                         order.ctypes.data))
        self.last_order = order
        return contrasted, p, q

    def test(self, contrasts, want_covar=False):
        contrasts = _f64(contrasts)
        kc = contrasts.shape[1]
        contrasted = pinned_empty((self.n, kc))
        ccov = pinned_empty((self.n, kc, kc)) if want_covar else None
        p = pinned_empty(self.n)
        self._check(self._lib.fn_test(self._h, kc, _dp(contrasts), _dp(contrasted), _dp(ccov), _dp(p)))
        return contrasted, ccov, p

    def bh(self, p):
        p = _f64(p).reshape(-1)
        q = pinned_empty(len(p))
        self._check(self._lib.fn_bh(self._h, len(p), _dp(p), _dp(q)))
        return q

    def bh_ranked(self, p):
        """(q, order): BH q-values and the genes by ascending p (stable, NaN last)."""
        p = _f64(p).reshape(-1)
        q, order = pinned_empty(len(p)), pinned_empty(len(p), np.uint32)
        if len(p):
            self._check(self._lib.fn_bh_ranked(self._h, len(p), _dp(p), _dp(q), order.ctypes.data))
        return q, order

    def bh_dev(self, n, p_ptr, q_ptr):
        self._check(self._lib.fn_bh_dev(self._h, n, p_ptr, q_ptr))

    def weights(self, theta):
        theta = _f64(theta).reshape(-1)
        out = pinned_empty((self.n, self.m))
        self._check(self._lib.fn_weights(self._h, _dp(theta), len(theta), _dp(out)))
        return out

    # ---- fitnoise.transform -----------------------------------------------------------------
    def transform_upload(self, x, design):
        x, design = _f64(x), _f64(design)
        self._tr_shape = x.shape
        self._check(self._lib.fn_transform_upload(self._h, x.shape[0], x.shape[1], x.ctypes.data,
                                                  design.shape[1], _dp(design)))

    def transform_sums(self, order, a):
        """Raw sums of one pass over x at polynomial coefficients a (order x m); see fn_transform.cuh."""
        a = _f64(a).reshape(order, -1)
        m = a.shape[1]
        sums = np.empty(2 + m * (1 + 3 * order))
        self._check(self._lib.fn_transform_sums(self._h, order, _dp(a), _dp(sums)))
        return sums

    def transform_apply(self, order, a, col_mean):
        a, col_mean = _f64(a).reshape(order, -1), _f64(col_mean).reshape(-1)
        y = np.empty(self._tr_shape)
        self._check(self._lib.fn_transform_apply(self._h, order, _dp(a), _dp(col_mean), _dp(y)))
        return y

    # ---- fused multi-GPU exchange ---------------------------------------------------------
    def comm_mailbox(self, world, max_values):
        """Allocate this rank's mailbox; returns its 64-byte CUDA IPC handle."""
        buf = ctypes.create_string_buffer(64)
        self._check(self._lib.fn_comm_mailbox(self._h, world, max_values, buf))
        return buf.raw

    def comm_connect(self, rank, world, handles):
        """handles: the ranks' IPC handles concatenated in rank order (world x 64 bytes)."""
        assert len(handles) == 64 * world
        self._check(self._lib.fn_comm_connect(self._h, rank, world, ctypes.c_char_p(handles)))

    def comm_disconnect(self):
        self._check(self._lib.fn_comm_disconnect(self._h))

    # ---- instrumentation ------------------------------------------------------------------
    def synchronize(self):
        self._check(self._lib.fn_synchronize(self._h))

    def set_tuning(self, lpg=0, threads=0, stages=0, ctas_per_sm=0):
        self._check(self._lib.fn_set_tuning(self._h, lpg, threads, stages, ctas_per_sm))

    def set_timing(self, on=True):
        self._check(self._lib.fn_set_timing(self._h, 1 if on else 0))

    def time_nll_grad(self, theta, reps=50):
        """Kernel-only ms per nll+grad launch: `reps` launches queued back to back between events."""
        theta = _f64(theta).reshape(-1)
        ms = ctypes.c_double()
        self._check(self._lib.fn_time_nll_grad(self._h, _dp(theta), len(theta), int(reps), ctypes.byref(ms)))
        return ms.value

    def last_kernel_ms(self):
        return float(self._lib.fn_last_kernel_ms(self._h))

    def launch_count(self):
        return int(self._lib.fn_launch_count(self._h))


_POOL = {}


def acquire_engine(device):
    """An idle engine of this device from the process-wide pool, or a new one.  Creating a handle
    (stream, events, pinned buffers) and destroying it costs tens of milliseconds; fits reuse them."""
    free = _POOL.setdefault(int(device), [])
    if free:
        return free.pop()
    return Engine(int(device))


def release_engine(engine):
    if engine is not None and getattr(engine, "_h", None):
        _POOL.setdefault(engine.device, []).append(engine)


_DEFAULT = {}


def default_engine(device=None):
    """Process-wide engine for the given (or the rank's) device."""
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    if device not in _DEFAULT:
        _DEFAULT[device] = Engine(device)
    return _DEFAULT[device]
