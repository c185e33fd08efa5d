"""Benjamini & Hochberg FDR q-values (reference: fitnoise/p_adjust.py:3-21), computed on the GPU:
radix sort + suffix-min scan (csrc/fn_bh.cuh).  Bit-exact with the reference given p."""
import numpy as np

from . import _native


def fdr(p):
    """ Benjamini & Hochberg FDR q-values.  NaN p-values count as 1.0 (and count in n). """
    p = np.array(p, dtype="float64").reshape(-1)
    if len(p) == 0:
        return p
    return _native.default_engine().bh(p)
