"""Design-matrix builders for the linear parameters other than K (tiny, per star, host
side).  Same names and results as thejoker/likelihood_helpers.py:8-37 and :232-233; the
in-memory drivers of that module live in thejoker.py / sharding.py here because they
hand device-resident work to the CUDA library instead of looping on the CPU."""
from __future__ import annotations

import numpy as np

__all__ = ["get_constant_term_design_matrix", "get_trend_design_matrix", "ln_normal"]


def get_constant_term_design_matrix(data, ids=None):
    """Columns for the systemic velocity and the per-survey offsets: a column of ones,
    then one 0/1 indicator column for every survey id after the first (in sorted id
    order).  Shape (n_times, n_surveys)."""
    n = len(data)
    survey = np.zeros(n, dtype=int) if ids is None else np.asarray(ids)
    extra = np.unique(survey)[1:]
    cols = [np.ones(n)] + [np.where(survey == sid, 1.0, 0.0) for sid in extra]
    return np.stack(cols, axis=1)


def get_trend_design_matrix(data, ids, poly_trend):
    """All columns of the design matrix except the Keplerian one:
    [1 | survey indicators | dt | dt^2 | ...] with dt = t - t_ref in days and
    ``poly_trend - 1`` powers of dt.  Shape (n_times, n_linear - 1), C-contiguous."""
    dt = np.asarray(data._t_bmjd, dtype=float) - data._t_ref_bmjd
    blocks = [get_constant_term_design_matrix(data, ids)]
    if poly_trend > 1:
        # running products (dt, dt*dt, (dt*dt)*dt, ...): bit-identical to the reference's
        # np.vander(dt, N=poly_trend, increasing=True)[:, 1:], unlike dt**k for k >= 3
        powers = [dt]
        for _ in range(2, poly_trend):
            powers.append(powers[-1] * dt)
        blocks.append(np.stack(powers, axis=1))
    return np.ascontiguousarray(np.concatenate(blocks, axis=1))


def ln_normal(x, mu, var):
    """Log of the normal density N(x | mu, var)."""
    resid = np.asarray(x) - mu
    return -0.5 * (resid * resid / var + np.log(2 * np.pi * var))
