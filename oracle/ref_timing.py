"""Time the reference's own numpy scorer (``score`` in fitnoise/fitnoise.py:63-71) on a gene sample.

BENCHMARK INFRASTRUCTURE: used by ``bench.py`` (``--impl reference`` and ``cpu_baseline``) only.
The scorer is a closure inside the reference's ``_fit_noise``; it is reached through the reference's
public call ``Model_t().fit_noise(y, design, use_theano=False)`` with ``scipy.optimize.minimize``
replaced, for the duration of the call, by a stub that calls the objective it is handed a fixed
number of times at a fixed theta and reports the time -- so what runs and what is timed is the
reference's unmodified per-gene loop (``score_row``: marginal -> transformed -> log_density,
distributions.py), after its own per-gene QR preparation (not timed: once per fit).

One call of ``score`` is the objective VALUE only; the reference's numpy path gets its gradient by
finite differences, i.e. 1 + P such calls per optimiser step (fitnoise.py:73-76).
"""
import time
import warnings

import numpy as np

from . import make_ref


class _Stop(Exception):
    pass


def score_seconds(y, design, theta, calls=1, model="Model_t"):
    """Seconds per call of the reference's ``score(theta)`` over the genes of y, and the value."""
    ref = make_ref.load()
    if ref is None:
        raise RuntimeError("oracle/_ref is missing: run __graft_entry__.build() where /root/reference is mounted")
    module = ref.fitnoise
    box = {}

    def stub(fun, x0, *args, **kwargs):
        fun(np.asarray(theta, dtype="float64"))            # warm (first touch of the prepared items)
        t0 = time.perf_counter()
        for _ in range(calls):
            value = fun(np.asarray(theta, dtype="float64"))
        box["seconds"] = (time.perf_counter() - t0) / calls
        box["value"] = float(value)
        raise _Stop()

    real = module.scipy.optimize.minimize
    module.scipy.optimize.minimize = stub
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            try:
                getattr(ref, model)().fit_noise(y, design, use_theano=False)
            except _Stop:
                pass
    finally:
        module.scipy.optimize.minimize = real
    return box["seconds"], box["value"]
