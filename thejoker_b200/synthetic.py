"""Synthetic radial-velocity data in the shape of the reference's own test fixture
(thejoker/tests/test_sampler.py:17-47) and a synthetic default-prior sampler, used by
tests/, bench.py and the smoke test.

``rv_curve`` solves Kepler's equation with a plain numpy Newton iteration.  It only
*generates data* (a few hundred epochs); it is not a likelihood path and the product
never evaluates likelihoods with it.
"""
from __future__ import annotations

import numpy as np

from . import units as u
from .data import RVData

J2000_BMJD = 51544.5


def rv_curve(t, t_ref, P, e, omega, M0, K):
    """K (cos(f + omega) + e cos omega) with M = 2 pi (t - t_ref)/P - M0
    (conventions: thejoker/samples.py:228-229, _keplerian_orbit.py:642-656)."""
    t = np.asarray(t, dtype=float)
    M = 2 * np.pi * (t - t_ref) / P - M0
    M = M - 2 * np.pi * np.round(M / (2 * np.pi))
    E = M + e * np.sin(M)
    for _ in range(100):
        dE = (E - e * np.sin(E) - M) / (1 - e * np.cos(E))
        E = E - dE
        if np.all(np.abs(dE) < 1e-15):
            break
    f = 2 * np.arctan2(np.sqrt(1 + e) * np.sin(E / 2), np.sqrt(1 - e) * np.cos(E / 2))
    return K * (np.cos(f + omega) + e * np.cos(omega))


def make_data(n_times=8, rng=None, v1=None, K=None, sigma=0.5, t_span_periods=3.0):
    """The reference fixture: P=51.8239 d, K=54.2473 km/s, v0=31.48502 km/s, e=0.3,
    omega=0.283, M0=2.592, t = J2000 + P sort(U(0,3)), err 0.5 km/s
    (test_sampler.py:17-47).  As there, the RVs are noise-free."""
    rng = np.random.default_rng() if rng is None else rng
    P = 51.8239
    K = 54.2473 if K is None else float(u.to_value(K, u.km / u.s, u.km / u.s))
    v0 = 31.48502
    t = J2000_BMJD + P * np.sort(rng.uniform(0, t_span_periods, n_times))
    truth = dict(P=P, K=K, e=0.3, omega=0.283, M0=2.592, v0=v0, t0=J2000_BMJD)
    rv = rv_curve(t, J2000_BMJD, P, 0.3, 0.283, 2.592, K) + v0
    if v1 is not None:
        rv = rv + float(v1) * (t - J2000_BMJD)
    err = np.full_like(rv, sigma)
    return RVData(t, rv * u.km / u.s, rv_err=err * u.km / u.s), truth


def make_noisy_data(n_times=64, seed=42, K=None, sigma=0.5, v1=None, t_span_periods=3.0):
    """BASELINE.md section 5 data: the fixture above plus Gaussian noise of sigma."""
    rng = np.random.default_rng(seed)
    data, truth = make_data(n_times, rng=rng, K=K, sigma=sigma, v1=v1,
                            t_span_periods=t_span_periods)
    noisy = data.rv.value + rng.normal(0, sigma, size=len(data))
    return RVData(data._t_bmjd, noisy * u.km / u.s, rv_err=data.rv_err), truth


def default_prior_columns(n, seed=123, s_lognormal=None, P_min=2.0, P_max=1024.0):
    """Host draw of the default prior in packed units: P [d] log-uniform, e ~ Beta(0.867,
    3.03), omega, M0 ~ U(-pi, pi), s = 0 or LogNormal(mu, sigma) [rv unit]."""
    rng = np.random.default_rng(seed)
    P = np.exp(rng.uniform(np.log(P_min), np.log(P_max), n))
    e = rng.beta(0.867, 3.03, n)
    omega = rng.uniform(-np.pi, np.pi, n)
    M0 = rng.uniform(-np.pi, np.pi, n)
    if s_lognormal is None:
        s = np.zeros(n)
    else:
        s = np.exp(rng.normal(s_lognormal[0], s_lognormal[1], n))
    return P, e, omega, M0, s
