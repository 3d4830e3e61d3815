"""RVData: the slice of thejoker/data.py the hot path reads.

Kept: constructor signature ``RVData(t, rv, rv_err, t_ref=None, clean=True)``, NaN
cleaning, the time sort, ``t_ref`` = earliest time, and the attributes CJokerHelper
extracts (``_t_bmjd``, ``_t_ref_bmjd``, ``rv``, ``rv_err``, ``ivar``, ``__len__``;
thejoker/src/fast_likelihood.pyx:155-166).  Plotting, table guessing and timeseries
I/O (data.py:157-505) are outside the hot path and not provided.
"""
from __future__ import annotations

import logging

import numpy as np

from . import units as u

logger = logging.getLogger("thejoker_b200")

__all__ = ["RVData"]


def _time_to_bmjd(t):
    """BMJD float array from an array, or an astropy Time if astropy is present
    (data.py:41-45 uses ``t.tcb.mjd``)."""
    if hasattr(t, "tcb") and hasattr(t.tcb, "mjd"):
        return np.atleast_1d(np.asarray(t.tcb.mjd, dtype=float))
    return np.atleast_1d(np.asarray(t, dtype=float))


class RVData:
    """Time-domain radial velocity measurements for a single target.

    Parameters
    ----------
    t : array_like
        Barycentric MJD of each measurement (an ``astropy.time.Time`` is accepted
        when astropy is installed).
    rv : Quantity [speed] or array_like
        Radial velocities; bare arrays are taken to be in km/s.
    rv_err : Quantity [speed] or array_like
        1-D standard deviations.  A 2-D covariance is stored but cannot be used on
        the hot path (the reference has the same limit: fast_likelihood.pyx:83,
        data_helpers.py:102-107).
    t_ref : float (optional)
        Reference time in BMJD; default the earliest time.  ``False`` disables it.
    clean : bool
        Drop non-finite points (data.py:75-95).
    """

    def __init__(self, t, rv, rv_err, t_ref=None, clean=True):
        _t = _time_to_bmjd(t)
        rv = rv if isinstance(rv, u.Quantity) else u.Quantity(rv, u.km / u.s)
        rv_err = rv_err if isinstance(rv_err, u.Quantity) else u.Quantity(rv_err, rv.unit)
        if not rv.unit.is_equivalent(u.km / u.s):
            raise u.UnitsError("rv must have units of speed")
        self.rv = u.Quantity(np.atleast_1d(rv.value), rv.unit)
        self.rv_err = u.Quantity(np.atleast_1d(rv_err.value), rv_err.unit)
        self._t_bmjd = _t

        if self.rv_err.ndim == 1:
            self._has_cov = False
            if not self.rv_err.unit.is_equivalent(u.km / u.s):
                raise u.UnitsError("1-D rv_err must have units of speed")
        elif self.rv_err.ndim == 2:
            self._has_cov = True
        else:
            raise ValueError("rv_err must be 1- or 2-dimensional")

        n = self.rv.size
        if self.rv_err.shape != (n, n) and self.rv_err.shape != (n,):
            raise ValueError(f"Invalid shape for input RV error {self.rv_err.shape}. Should either "
                             f"be ({n},) or ({n}, {n})")
        if self._t_bmjd.shape != self.rv.shape:
            raise ValueError(f"Shape of input times and RVs must be consistent "
                             f"({self._t_bmjd.shape} vs {self.rv.shape})")

        if clean:
            idx = np.isfinite(self._t_bmjd) & np.isfinite(self.rv.value)
            if self._has_cov:
                idx &= np.isfinite(self.rv_err.value).all(axis=0)
            else:
                idx &= np.isfinite(self.rv_err.value)
            n_filter = len(idx) - idx.sum()
            if n_filter > 0:
                logger.info(f"Filtering {n_filter} NaN/Inf data points")
            self._apply_index(idx)

        # sort on times (data.py:97-105)
        self._apply_index(self._t_bmjd.argsort())

        if t_ref is False:
            self.t_ref = None
            self._t_ref_bmjd = 0.0
        else:
            if t_ref is None:
                t_ref = self._t_bmjd.min() if len(self._t_bmjd) else 0.0
            elif hasattr(t_ref, "tcb"):
                t_ref = float(t_ref.tcb.mjd)
            self.t_ref = float(t_ref)
            self._t_ref_bmjd = float(t_ref)

    def _apply_index(self, idx):
        self._t_bmjd = self._t_bmjd[idx]
        self.rv = self.rv[idx]
        if self._has_cov:
            self.rv_err = self.rv_err[idx][:, idx]
        else:
            self.rv_err = self.rv_err[idx]

    @property
    def t(self):
        """Observation times, BMJD."""
        return self._t_bmjd

    @property
    def cov(self):
        if self._has_cov:
            return self.rv_err
        return u.Quantity(np.diag(self.rv_err.value**2), self.rv_err.unit**2)

    @property
    def ivar(self):
        """Inverse variance (data.py:147-153)."""
        if self._has_cov:
            return u.Quantity(np.linalg.inv(self.rv_err.value), u.one / self.rv_err.unit)
        return u.Quantity(1.0 / self.rv_err.value**2, u.one / self.rv_err.unit**2)

    def phase(self, P, t_ref=None, t0=None):
        """Orbital phase of each observation for period P [day], relative to ``t_ref``
        [BMJD] (default: the data's reference epoch).  data.py:365-392; ``t0`` is the
        reference's deprecated name for ``t_ref`` and is still accepted."""
        if t_ref is None:
            t_ref = t0
        t_ref = self._t_ref_bmjd if t_ref is None else t_ref
        P = u.to_value(P, u.day, u.day)
        return ((self._t_bmjd - t_ref) / P) % 1.0

    def __len__(self):
        return len(self._t_bmjd)

    def __getitem__(self, slc):
        err = self.rv_err[slc][:, slc] if self._has_cov else self.rv_err[slc]
        return self.__class__(t=self._t_bmjd[slc], rv=self.rv[slc], rv_err=err,
                              t_ref=self.t_ref if self.t_ref is not None else False, clean=False)

    def copy(self):
        return self[slice(None)]
