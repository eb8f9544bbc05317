"""Host-side data preparation with the reference's semantics (one-off, outside the hot path)."""
from __future__ import annotations

import numpy as np


def gaussian_normalise(x):
    """`gaussian_normalise!` (src/datatypes/gaussian_cluster.jl:85-94): per column, subtract the
    median and divide by 0.5*(median - 5 % quantile) + eps.  Returns a new array."""
    x = np.array(x, dtype=np.float64, copy=True)
    med = np.median(x, axis=0)
    q05 = np.quantile(x, 0.05, axis=0)  # Julia's default quantile is the same linear rule
    sig = 0.5 * (med - q05) + np.finfo(np.float64).eps
    return (x - med) / sig


def coerce_categorical(data):
    """`coerce_categorical` (src/datatypes/categorical_cluster.jl:81-92): recode each column to
    1..L in order of first appearance."""
    data = np.asarray(data)
    out = np.empty(data.shape, dtype=np.int64)
    for j in range(data.shape[1]):
        seen = {}
        for i, v in enumerate(data[:, j].tolist()):
            out[i, j] = seen.setdefault(v, len(seen) + 1)
    return out
