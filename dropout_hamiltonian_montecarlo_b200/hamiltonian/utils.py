"""Mirror of reference ``hamiltonian/utils.py:4-26`` (host-side data helpers; not on the hot path)."""
from collections.abc import Iterable

import numpy as np


def one_hot(y, num_classes):
    """utils.py:4-8."""
    y = np.asarray(y)
    enc = np.zeros((len(y), num_classes))
    enc[np.arange(len(y)), y.astype(np.int64)] = 1.0
    return enc


def scaler_fit(X):
    """utils.py:10-14."""
    lo, hi = np.amin(X, 0), np.amax(X, 0)
    return (X - lo) / (hi - lo), lo, hi


def scaler_scale(X, min_col, max_col):
    """utils.py:16-18."""
    return (X - min_col) / (max_col - min_col)


def flatten(items):
    """utils.py:20-26."""
    for x in items:
        if isinstance(x, Iterable) and not isinstance(x, (str, bytes)):
            yield from flatten(x)
        else:
            yield x
