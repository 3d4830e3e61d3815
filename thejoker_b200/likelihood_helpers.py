"""Host-side pieces of thejoker/likelihood_helpers.py: the design-matrix builders
(tiny, per star) and the in-memory drivers, which here hand device-resident work to
the CUDA library instead of looping on the CPU."""
from __future__ import annotations

import numpy as np

__all__ = ["get_constant_term_design_matrix", "get_trend_design_matrix", "ln_normal"]


def get_constant_term_design_matrix(data, ids=None):
    """Constant-term columns of the design matrix: a column of ones plus one
    indicator column per additional survey id (likelihood_helpers.py:8-25)."""
    if ids is None:
        ids = np.zeros(len(data), dtype=int)
    ids = np.array(ids)
    unq_ids = np.unique(ids)
    constant_part = np.zeros((len(data), len(unq_ids)))
    constant_part[:, 0] = 1.0
    for j, id_ in enumerate(unq_ids[1:]):
        constant_part[ids == id_, j + 1] = 1.0
    return constant_part


def get_trend_design_matrix(data, ids, poly_trend):
    """Design matrix for the linear parameters without the K column:
    [1 | 1{id==k}.. | dt | dt^2 ..], dt = t - t_ref (likelihood_helpers.py:28-37)."""
    const_M = get_constant_term_design_matrix(data, ids)
    dt = data._t_bmjd - data._t_ref_bmjd
    trend_M = np.vander(dt, N=poly_trend, increasing=True)[:, 1:]
    return np.ascontiguousarray(np.hstack((const_M, trend_M)))


def ln_normal(x, mu, var):
    """likelihood_helpers.py:232-233."""
    return -0.5 * (np.log(2 * np.pi * var) + (x - mu) ** 2 / var)
