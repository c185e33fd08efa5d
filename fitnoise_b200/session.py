"""One fit's device-resident state: this rank's shard of the genes on its GPU plus the collectives
that make the per-rank results global.  The numerical work is all in the CUDA library
(``_native.Engine``); this module only shards, reduces and gathers."""
import os
import weakref

import numpy as np

from . import _native, parallel


def _give_back(engine, pooled):
    if pooled == "rank":
        _RANK_STATE["busy"] = False           # the rank's engine stays alive, connected, with its buffers
    elif pooled:
        _native.release_engine(engine)
    elif hasattr(engine, "close"):
        engine.close()


# Multi-GPU: what a rank sets up ONCE per process and every later fit reuses -- its engine (device
# buffers, pinned result pool), the torch stream that carries kernels and collectives, and the fused
# exchange (mailboxes mapped into every rank with CUDA IPC: milliseconds of handle exchange and
# cudaIpcOpenMemHandle that a 20,000-gene fit cannot afford to repeat).
_RANK_STATE = {}
_MAILBOX_VALUES = 8 + 2 * 256      # room for every model the kernels accept (per-sample blocks: m <= 256)


def _shard_min_bytes():
    """Matrices smaller than this (bytes of y) are fitted whole by every rank (FN_SHARD_MIN_MB, default 8)."""
    try:
        return float(os.environ.get("FN_SHARD_MIN_MB", "8")) * 1048576.0
    except ValueError:
        return 8.0 * 1048576.0


class _Resident(object):
    def __init__(self, session):
        self.session, self.on = session, False

    def __enter__(self):
        s = self.session
        usable = (hasattr(s.local, "resident_begin") and s.family.nf == 0 and s.P > 0
                  and (s.world == 1 or s._fused) and not os.environ.get("FN_NO_RESIDENT"))
        if usable:
            s.local.resident_begin()
            self.on = True
        return self

    def __exit__(self, *exc):
        if self.on and self.session.local is not None:
            self.session.local.resident_end()
        self.on = False
        return False


class Session(object):
    """``local`` is this rank's evaluator.  By default a ``_native.Engine`` (CUDA; raises without
    a GPU).  Tests of the sharding logic on a CPU-only box pass their own stand-in through
    ``engine_factory`` -- the product never does."""

    def __init__(self, y, aux, family, controls, design_noise, design_ctrl, engine_factory=None):
        self.rank, self.world = parallel.world()
        y = np.asarray(y, dtype=np.float64)
        self.n, self.m = y.shape
        # Small problems are not worth sharding: at 20,000 x 6 an evaluation is ~10 us on one GPU and the
        # exchanges, gathers and rendezvous of a sharded fit cost more than the work they spread (3.1 ms on
        # 8 GPUs against 1.8 ms on one).  Below the threshold every rank fits ALL genes on its own GPU --
        # same kernels, same order of summation, identical results on every rank, no collective.
        # (with a stand-in evaluator -- the CPU tests -- only when the threshold is set explicitly)
        if (self.world > 1 and (engine_factory is None or "FN_SHARD_MIN_MB" in os.environ)
                and y.nbytes < _shard_min_bytes()):
            self.rank, self.world = 0, 1
        self.bounds = parallel.shard_bounds(self.n, self.world)
        self.lo, self.hi = int(self.bounds[self.rank]), int(self.bounds[self.rank + 1])
        self.family = family
        self.P = family.theta_length()
        self._cuda_reduce = None
        if engine_factory is None:
            engine_factory = self._cuda_engine
        self._pooled = False
        self.local = engine_factory()
        # Return the engine when this session goes away.  weakref.finalize keeps the engine
        # reachable until then, so that a garbage-collected cycle containing the session can never
        # run the engine's own finaliser first.
        self._finalizer = weakref.finalize(self, _give_back, self.local, self._pooled)
        sl = slice(self.lo, self.hi)
        info = self.local.upload(
            y[sl], None if aux is None else np.asarray(aux, dtype=np.float64)[sl], family,
            None if controls is None else np.asarray(controls)[sl], design_noise, design_ctrl)
        # global preparation summary
        if self.world == 1:
            self.n_eligible, self.n_bad_counts = int(info.n_eligible), int(info.n_bad_counts)
            self.row_variance_mean = float(info.row_variance_mean)
        else:
            rows = float(info.n_variance_rows)
            sums = parallel.allreduce_sum([float(info.n_eligible), float(info.n_bad_counts),
                                           self._nan0(info.row_variance_mean * rows), rows])
            self.n_eligible, self.n_bad_counts = int(sums[0]), int(sums[1])
            self.row_variance_mean = sums[2] / sums[3] if sums[3] > 0 else np.nan
        self.coef_token = None
        self._fused = False
        if self._cuda_reduce is not None and os.environ.get("FN_REDUCE", "fused") == "fused":
            state = _RANK_STATE
            if state.get("engine") is self.local and "fused" in state:
                self._fused = state["fused"]              # connected by an earlier fit of this process
            else:
                self._connect_fused()
                if state.get("engine") is self.local:
                    state["fused"] = self._fused

    def _connect_fused(self):
        """Exchange the ranks' mailbox handles (CUDA IPC) so that the nll+grad kernel itself adds the
        sums of all GPUs over NVLink (csrc/fn_kernels.cuh grid_reduce); FN_REDUCE=nccl keeps the
        kernel + ncclAllReduce path instead."""
        import torch.distributed as dist
        ok, handle = 1, None
        try:
            handle = self.local.comm_mailbox(self.world, max(_MAILBOX_VALUES, 8 + max(2, self.family.nf) * self.m))
        except _native.NativeError:               # CUDA IPC not available on this box
            ok = 0
        handles = [None] * self.world
        dist.all_gather_object(handles, handle)   # every rank takes part, whatever happened
        if ok and all(h is not None for h in handles):
            try:
                self.local.comm_connect(self.rank, self.world, b"".join(handles))
            except _native.NativeError:           # peer access not available
                ok = 0
        else:
            ok = 0
        with self._torch.cuda.stream(self._stream):
            flag = self._torch.tensor([ok], dtype=self._torch.int64, device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            agreed = int(flag.item())
        if not agreed:                            # every rank falls back together: kernel + ncclAllReduce
            try:
                self.local.comm_disconnect()
            except _native.NativeError:
                pass
        dist.barrier()
        self._fused = bool(agreed)

    @staticmethod
    def _nan0(x):
        return 0.0 if x != x else float(x)

    def _cuda_engine(self):
        if self.world > 1:
            # One torch stream carries the kernels AND the collectives, so that the all-reduce is
            # ordered after the kernel that produced its input (the handle would otherwise launch on
            # a private non-blocking stream that the NCCL stream knows nothing about).
            import torch
            self._torch = torch
            state = _RANK_STATE
            if state.get("engine") is not None and not state.get("busy") and state.get("world") == self.world:
                self._stream = state["stream"]
                engine = state["engine"]
            else:
                torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", self.rank)))
                self._stream = torch.cuda.Stream()
                engine = _native.Engine(torch.cuda.current_device(), self._stream.cuda_stream)
                if state.get("engine") is None or state.get("world") != self.world:
                    state.clear()
                    state.update(engine=engine, stream=self._stream, world=self.world)
            if state.get("engine") is engine:
                state["busy"] = True
                self._pooled = "rank"
            with torch.cuda.stream(self._stream):
                self._cuda_reduce = torch.zeros(self.P + 3, dtype=torch.float64, device="cuda")
            self._stream.synchronize()
            return engine
        self._pooled = True
        return _native.acquire_engine(int(os.environ.get("LOCAL_RANK", "0")))

    def close(self):
        """Give the GPU state back: pooled engines are kept (with their buffers) for the next fit."""
        self.local = None
        self._finalizer()

    # ---- likelihood -----------------------------------------------------------------------
    def resident(self):
        """Context manager for a run of nll_grad calls (the optimiser loop): keeps ONE evaluation
        kernel resident on the GPU so that an evaluation costs no kernel launch (fn_resident_begin).
        Used when this rank's kernel produces the global totals by itself (one GPU, or the fused
        exchange); a no-op for factor models and for the kernel + ncclAllReduce path -- a collective
        could not run beside a kernel that occupies every SM."""
        return _Resident(self)

    def nll_grad(self, theta, factors=None):
        """Global (all ranks) value and gradient of the REML objective at theta:
        (value, grad, grad_factors | None, n_used, n_bad).  ``factors``: m x nf matrix of the factor
        models; its gradient comes back in the same shape."""
        theta = np.asarray(theta, dtype=np.float64)
        if factors is not None:
            value, grad, gf, used, bad = self.local.nll_grad_factors(theta, factors)
            if self.world > 1 and not self._fused:
                out = parallel.allreduce_sum(np.concatenate([[value], grad, gf.reshape(-1), [used, bad]]))
                P, F = len(grad), gf.size
                value, grad, gf = float(out[0]), out[1:1 + P], out[1 + P:1 + P + F].reshape(gf.shape)
                used, bad = int(out[-2]), int(out[-1])
            return value, grad, gf, used, bad
        if self._fused:
            value, grad, used, bad = self.local.nll_grad(theta)   # the kernel epilogue already added all ranks
            return value, grad, None, used, bad
        if self._cuda_reduce is not None:
            buf = self._cuda_reduce
            with self._torch.cuda.stream(self._stream):
                self.local.nll_grad_dev(theta, buf.data_ptr())
                parallel.allreduce_sum_tensor(buf)
                out = buf.cpu().numpy()
            return float(out[0]), out[1:1 + self.P].copy(), None, int(out[self.P + 1]), int(out[self.P + 2])
        value, grad, used, bad = self.local.nll_grad(theta)
        if self.world == 1:
            return value, grad, None, used, bad
        out = parallel.allreduce_sum(np.concatenate([[value], grad, [used, bad]]))
        return float(out[0]), out[1:1 + self.P], None, int(out[self.P + 1]), int(out[self.P + 2])

    def set_factors(self, factors):
        """Factor vectors (m x nf) used by the per-gene statistics, coefficients and weights."""
        self.local.set_factors(factors)

    # ---- per-gene results: local at world == 1, gathered device-to-device otherwise ---------
    def _gather(self, which, keep_device=False):
        """Result `which` (``_native.RES_*``) of ALL genes on every rank.  The shards are gathered
        device-to-device (NCCL all-gather on the stream the kernels ran on) and come to the host once,
        into pinned memory.  ``keep_device``: also return the gathered device tensor."""
        ptr, rows, cols, dtype = self.local.result_dev(which)
        full = self._gather_dev(ptr, rows, cols, dtype)
        shape = (self.n,) if cols == 1 and which not in (_native.RES_COEF, _native.RES_CONTRASTED) else (self.n, cols)
        out = _native.pinned_empty(shape, dtype)
        self.local.fetch(out, full.data_ptr())
        return (out, full) if keep_device else out

    def _gather_dev(self, ptr, rows, cols, dtype):
        """All ranks' row blocks (device memory at ptr: rows x cols) as one contiguous n x cols device
        tensor on this rank."""
        torch = self._torch
        import torch.distributed as dist
        tdtype = {np.dtype(np.float64): torch.float64, np.dtype(np.int32): torch.int32,
                  np.dtype(np.uint32): torch.int32}[np.dtype(dtype)]
        sizes = np.diff(self.bounds)
        width = int(sizes.max())
        with torch.cuda.stream(self._stream):
            local = _device_tensor(torch, ptr, (rows, cols), tdtype)
            if rows != width:                                   # ragged shards: pad to the widest
                padded = torch.zeros((width, cols), dtype=tdtype, device=local.device)
                padded[:rows] = local
                local = padded
            full = torch.empty((self.world * width, cols), dtype=tdtype, device=local.device)
            dist.all_gather_into_tensor(full, local.contiguous())
            if int(sizes.min()) != width:                       # drop the padding rows
                full = torch.cat([full[r * width:r * width + int(sizes[r])] for r in range(self.world)])
        return full

    @property
    def _device_gather(self):
        # several ranks on CUDA engines; a stand-in evaluator (CPU tests, gloo) gathers through the host
        return self.world > 1 and self._cuda_reduce is not None

    def per_gene(self, theta):
        score, d2, p = self.local.per_gene(theta)
        if self.world == 1:
            return score, d2, p
        if not self._device_gather:
            return (parallel.allgather_rows(score, self.n), parallel.allgather_rows(d2, self.n),
                    parallel.allgather_rows(p, self.n))
        return (self._gather(_native.RES_SCORE), self._gather(_native.RES_D2), self._gather(_native.RES_PDF))

    def noise_stats(self, theta):
        if self.world == 1:                          # one GPU, or a small problem every rank fits whole
            return self.local.noise_stats(theta)
        if not self._device_gather:
            noise_p, averages = self.local.noise_stats(theta)
            return parallel.allgather_rows(noise_p, self.n), parallel.allgather_rows(averages, self.n)
        self.local.noise_stats_dev(theta)
        return self._gather(_native.RES_NOISE_P), self._gather(_native.RES_AVERAGES)

    # ---- coefficients, tests --------------------------------------------------------------
    def coef(self, theta, design, token):
        """Coefficients now; their covariances and posterior df stay on the GPU until asked for.
        With several ranks the covariances of ALL genes are gathered to this rank's GPU right away
        (NVLink), so that fetching them later is a local copy and not a collective that every
        rank would have to enter."""
        self.coef_token = token
        if self.world == 1:
            coef, _, _ = self.local.coef(theta, design, want_covar=False)
            return coef
        if not self._device_gather:
            # the covariances are gathered right away as well: a later, lazy gather would be a
            # collective that only some ranks might enter
            coef, covar, dfpost = self.local.coef(theta, design, want_covar=True)
            self._covar_host = (parallel.allgather_rows(covar, self.n), parallel.allgather_rows(dfpost, self.n))
            return parallel.allgather_rows(coef, self.n)
        self.local.coef_dev(theta, design)
        coef = self._gather(_native.RES_COEF)
        c = coef.shape[1]
        ptr, rows, cols, dtype = self.local.result_dev(_native.RES_COVAR)
        self._covar_dev = self._gather_dev(ptr, rows, cols, dtype)
        ptr, rows, cols, dtype = self.local.result_dev(_native.RES_DFPOST)
        self._dfpost_dev = self._gather_dev(ptr, rows, cols, dtype)
        self._coef_c = c
        return coef

    def coef_covar(self, theta, design, token):
        """(covar n x c x c, dfpost n); never a collective."""
        if self.world == 1:
            if self.coef_token is not token:          # another fit_coef ran on this session since
                self.coef(theta, design, token)
            return self.local.coef_fetch()
        if self.coef_token is token and not self._device_gather:
            return self._covar_host
        if self.coef_token is not token:
            raise RuntimeError("the coefficient covariances of this fit are no longer on the GPU: another "
                               "fit_coef ran on the session since (with several ranks they cannot be "
                               "recomputed lazily: every rank would have to take part)")
        c = self._coef_c
        covar, dfpost = _native.pinned_empty((self.n, c, c)), _native.pinned_empty(self.n)
        self.local.fetch(covar, self._covar_dev.data_ptr())
        self.local.fetch(dfpost, self._dfpost_dev.data_ptr())
        return covar, dfpost

    def test(self, contrasts):
        """contrasted, p_values, q_values (BH over ALL genes), order (genes by ascending p, stable,
        NaN last: the top-table order)."""
        if self.world == 1:
            contrasted, p, q = self.local.test_bh(contrasts)   # p never leaves the device before the adjustment
            return contrasted, p, q, self.local.last_order
        if not self._device_gather:
            contrasted, _, p = self.local.test(contrasts)
            contrasted, p = parallel.allgather_rows(contrasted, self.n), parallel.allgather_rows(p, self.n)
            q, order = self.local.bh_ranked(p)
            return contrasted, p, q, order
        torch = self._torch
        self.local.test_dev(contrasts)
        contrasted = self._gather(_native.RES_CONTRASTED)
        p, p_dev = self._gather(_native.RES_P, keep_device=True)
        # every rank adjusts the full, device-resident p vector on its own GPU
        with torch.cuda.stream(self._stream):
            q_dev = torch.empty(self.n, dtype=torch.float64, device=p_dev.device)
        order_ptr = self.local.bh_ranked_dev(self.n, p_dev.data_ptr(), q_dev.data_ptr())
        q, order = _native.pinned_empty(self.n), _native.pinned_empty(self.n, np.uint32)
        self.local.fetch(q, q_dev.data_ptr())
        self.local.fetch(order, order_ptr)
        return contrasted, p, q, order

    def bh_ranked(self, p):
        """(q, order) of a full-length p vector every rank holds."""
        return self.local.bh_ranked(p)

    def bh(self, p):
        # every rank holds the full p vector after the gather and adjusts it on its own GPU
        return self.local.bh(p)

    def weights(self, theta):
        if self.world == 1:
            return self.local.weights(theta)
        return parallel.allgather_rows(self.local.weights(theta), self.n)


class _DevicePointer(object):
    """Zero-copy view of raw device memory for torch (``__cuda_array_interface__``)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def _device_tensor(torch, ptr, shape, tdtype):
    typestr = {torch.float64: "<f8", torch.int32: "<i4"}[tdtype]
    if shape[0] == 0:
        return torch.empty(shape, dtype=tdtype, device="cuda")
    return torch.as_tensor(_DevicePointer(ptr, shape, typestr), device="cuda")
