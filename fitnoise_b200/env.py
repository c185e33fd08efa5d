"""Small host-side helpers of the public API: immutable-record base class, array coercion and the
JSON export used by the R wrapper (reference: fitnoise/env.py:22-51,57-83)."""
import copy

import numpy as np


class Withable(object):
    """Record whose ``_with(**fields)`` returns a modified shallow copy; nothing is mutated in
    place (reference env.py:46-51).  Every model / fit result is one of these."""

    def _with(self, **fields):
        clone = copy.copy(self)
        clone.__dict__.update(fields)
        return clone


def as_tensor(x, dtype="float64"):
    return np.asarray(x, dtype=dtype)


def as_scalar(x, dtype="float64"):
    x = as_tensor(x, dtype)
    assert x.ndim == 0
    return x


def as_vector(x, dtype="float64"):
    x = as_tensor(x, dtype)
    assert x.ndim == 1
    return x


def as_matrix(x, dtype="float64"):
    x = as_tensor(x, dtype)
    assert x.ndim == 2
    return x


def as_jsonic(item):
    """Plain lists / dicts / numbers for json.dumps, NaN -> None, attributes starting with an
    underscore dropped (reference env.py:22-43; what R's jsonlite reads back)."""
    if isinstance(item, Withable):
        return {key: as_jsonic(value) for key, value in vars(item).items() if not key.startswith("_")}
    if isinstance(item, dict):
        return {key: as_jsonic(value) for key, value in item.items()}
    if isinstance(item, (list, tuple)) or hasattr(item, "__fn_lazy_sequence__"):
        return [as_jsonic(value) for value in item]
    if isinstance(item, np.ndarray):
        return as_jsonic(item.tolist())
    if isinstance(item, np.generic):
        item = item.item()
    if hasattr(item, "__fn_jsonic__"):
        return item.__fn_jsonic__()
    if isinstance(item, float) and item != item:
        return None
    return item
