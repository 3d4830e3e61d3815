"""ctypes front-end of the CPU oracle (oracle/joker_oracle.c, oracle/joker_truth.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(thejoker_b200/) never imports this module.

PARITY STATUS: pinned -- bit for bit to the reference's own compiled Cython
(oracle/ref_cython.py, tests/test_ref_pinning.py, tests/golden/ref_*.npz) for everything
the reference implements, and to twobody's own stored outputs (the reference's noiseless
docs/examples/*.ecsv, tests/golden/ref_examples.npz; 4e-11 K, the resolution of their
time stamps) for the Kepler function it takes from that absent third party.  See the
header of joker_oracle.c.

`OracleHelper` mirrors the method surface of the reference's CJokerHelper
(thejoker/src/fast_likelihood.pyx:70-576) on plain arrays.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libjoker_oracle.so")
_lib = None

_dp = ctypes.POINTER(ctypes.c_double)


class OrcSpec(ctypes.Structure):
    _fields_ = [
        ("n_times", ctypes.c_int),
        ("n_linear", ctypes.c_int),
        ("t0", ctypes.c_double),
        ("t", _dp),
        ("rv", _dp),
        ("ivar", _dp),
        ("trend_M", _dp),
        ("mu", _dp),
        ("Lambda", _dp),
        ("K_prior_kind", ctypes.c_int),
        ("sigma_K0", ctypes.c_double),
        ("P0", ctypes.c_double),
        ("max_K", ctypes.c_double),
        ("jitter_mode", ctypes.c_int),
        ("kepler_tol", ctypes.c_double),
        ("kepler_maxiter", ctypes.c_int),
        ("kepler_variant", ctypes.c_int),
    ]


def build(force: bool = False) -> str:
    """Compile oracle/libjoker_oracle.so with the recipe in oracle/Makefile."""
    srcs = [os.path.join(_HERE, f) for f in ("joker_oracle.c", "joker_truth.c", "joker_oracle.h")]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs
    )
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "libjoker_oracle.so"], check=True,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return _LIB_PATH


def _lapack_pointers():
    """Addresses of the LAPACK routines the reference cimports
    (fast_likelihood.pyx:19: scipy.linalg.cython_lapack)."""
    import scipy.linalg.cython_lapack as cl

    api = ctypes.pythonapi
    api.PyCapsule_GetName.restype = ctypes.c_char_p
    api.PyCapsule_GetName.argtypes = [ctypes.py_object]
    api.PyCapsule_GetPointer.restype = ctypes.c_void_p
    api.PyCapsule_GetPointer.argtypes = [ctypes.py_object, ctypes.c_char_p]
    out = []
    for name in ("dgetrf", "dgetri", "dsysv"):
        cap = cl.__pyx_capi__[name]
        out.append(api.PyCapsule_GetPointer(cap, api.PyCapsule_GetName(cap)))
    return out


_blas_limited = False


def _limit_blas_threads():
    """The oracle parallelises over prior samples (OpenMP); the L x L / N x N LAPACK
    calls inside each thread must stay single-threaded (scipy's OpenBLAS is a pthreads
    build and warns / oversubscribes when entered from an OpenMP region)."""
    global _blas_limited
    if _blas_limited:
        return
    try:
        from threadpoolctl import threadpool_limits

        threadpool_limits(limits=1, user_api="blas")
    except Exception:
        pass
    _blas_limited = True


def load(use_lapack: bool = True):
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_LIB_PATH)
        lib.orc_set_lapack.argtypes = [ctypes.c_void_p] * 3
        lib.orc_eccentric_anomaly.restype = ctypes.c_double
        lib.orc_eccentric_anomaly.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                              ctypes.c_int, ctypes.c_int]
        lib.orc_design_column.argtypes = [ctypes.POINTER(OrcSpec), _dp, _dp]
        lib.orc_batch_marginal_ln_likelihood.argtypes = [ctypes.POINTER(OrcSpec), _dp,
                                                         ctypes.c_long, _dp, ctypes.c_int]
        lib.orc_likelihood_worker_full.restype = ctypes.c_double
        lib.orc_likelihood_worker_full.argtypes = [ctypes.POINTER(OrcSpec), _dp, ctypes.c_int] + [_dp] * 6
        lib.orc_batch_posterior_aAinv.argtypes = [ctypes.POINTER(OrcSpec), _dp, ctypes.c_long,
                                                  ctypes.c_int, _dp, _dp, _dp]
        lib.orc_truth_marginal_ln_likelihood.argtypes = [ctypes.POINTER(OrcSpec), _dp,
                                                         ctypes.c_long, _dp, _dp, ctypes.c_int]
        lib.orc_truth_posterior_aA.argtypes = [ctypes.POINTER(OrcSpec), _dp, ctypes.c_int, _dp, _dp]
        _lib = lib
    if use_lapack:
        _limit_blas_threads()
        _lib.orc_set_lapack(*_lapack_pointers())
    else:
        _lib.orc_set_lapack(None, None, None)
    return _lib


def _p(a):
    return a.ctypes.data_as(_dp)


def _f8(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class OracleHelper:
    """CPU stand-in for the reference's CJokerHelper on plain float64 arrays.

    Parameters mirror what CJokerHelper.__init__ extracts (pyx:125-253).
    ``jitter_mode=0`` reproduces the reference as written (the jitter ``s`` is
    ignored, pyx:458 is a dead store); ``jitter_mode=1`` applies it.
    """

    packed_order = ["P", "e", "omega", "M0", "s"]

    def __init__(self, t, rv, ivar, t0, trend_M, mu, Lambda, K_prior_kind=0, sigma_K0=0.0,
                 P0=1.0, max_K=np.inf, jitter_mode=0, kepler_tol=1e-10, kepler_maxiter=128,
                 kepler_variant=0, use_lapack=True):
        self._lib = load(use_lapack)
        self.t, self.rv, self.ivar = _f8(t), _f8(rv), _f8(ivar)
        self.trend_M = _f8(trend_M).reshape(len(self.t), -1)
        self.n_times = len(self.t)
        self.n_linear = 1 + self.trend_M.shape[1]
        self.mu, self.Lambda = _f8(mu)[: self.n_linear].copy(), _f8(Lambda)[: self.n_linear].copy()
        if len(self.mu) != self.n_linear or len(self.Lambda) != self.n_linear:
            raise ValueError("mu / Lambda must have n_linear entries")
        self.spec = OrcSpec(self.n_times, self.n_linear, float(t0), _p(self.t), _p(self.rv),
                            _p(self.ivar), _p(self.trend_M), _p(self.mu), _p(self.Lambda),
                            int(K_prior_kind), float(sigma_K0), float(P0), float(max_K),
                            int(jitter_mode), float(kepler_tol), int(kepler_maxiter),
                            int(kepler_variant))
        self.a = self.A = self.Ainv = self.b = self.B = self.Binv = None

    @classmethod
    def from_spec(cls, spec: dict, **overrides):
        """Build from the plain-array dict a product CJokerHelper exposes as ``.spec``."""
        keys = ("t", "rv", "ivar", "t0", "trend_M", "mu", "Lambda", "K_prior_kind", "sigma_K0",
                "P0", "max_K", "jitter_mode")
        kw = {k: spec[k] for k in keys if k in spec}
        kw.update(overrides)
        return cls(**kw)

    # pyx:428-469
    def batch_marginal_ln_likelihood(self, chunk, n_threads=1):
        chunk = _f8(chunk)
        if chunk.ndim != 2 or chunk.shape[1] != 5:
            raise ValueError("chunk must have shape (n, 5): P, e, omega, M0, s")
        ll = np.full(chunk.shape[0], np.nan)
        self._lib.orc_batch_marginal_ln_likelihood(ctypes.byref(self.spec), _p(chunk),
                                                   chunk.shape[0], _p(ll), int(n_threads))
        return ll

    # pyx:547-576
    def test_likelihood_worker(self, chunk_row, clamp=None):
        row = _f8(chunk_row)
        N, L = self.n_times, self.n_linear
        self.a, self.A, self.Ainv = np.zeros(L), np.zeros((L, L)), np.zeros((L, L))
        self.b, self.B, self.Binv = np.zeros(N), np.zeros((N, N)), np.zeros((N, N))
        co = -1 if clamp is None else int(bool(clamp))
        return self._lib.orc_likelihood_worker_full(ctypes.byref(self.spec), _p(row), co,
                                                    _p(self.a), _p(self.A), _p(self.Ainv),
                                                    _p(self.b), _p(self.B), _p(self.Binv))

    def posterior_aAinv(self, chunk, clamp=None):
        chunk = _f8(chunk).reshape(-1, 5)
        n, L = chunk.shape[0], self.n_linear
        ll, a, Ainv = np.zeros(n), np.zeros((n, L)), np.zeros((n, L, L))
        co = -1 if clamp is None else int(bool(clamp))
        self._lib.orc_batch_posterior_aAinv(ctypes.byref(self.spec), _p(chunk), n, co, _p(ll),
                                            _p(a), _p(Ainv))
        return ll, a, Ainv

    # pyx:471-545
    def batch_get_posterior_samples(self, chunk, n_linear_samples_per, rng, clamp=None):
        chunk = _f8(chunk).reshape(-1, 5)
        n, L = chunk.shape[0], self.n_linear
        lls, a, Ainv = self.posterior_aAinv(chunk, clamp)
        samples = np.zeros((n, n_linear_samples_per, 5 + L))
        ll = np.zeros((n, n_linear_samples_per))
        for i in range(n):
            lin = rng.multivariate_normal(a[i], np.linalg.inv(Ainv[i]), size=n_linear_samples_per)
            samples[i, :, :5] = chunk[i]
            samples[i, :, 5:] = lin
            ll[i] = lls[i]
        return samples.reshape(n * n_linear_samples_per, -1), ll.reshape(-1)

    def design_column(self, chunk_row):
        z = np.zeros(self.n_times)
        self._lib.orc_design_column(ctypes.byref(self.spec), _p(_f8(chunk_row)), _p(z))
        return z

    def truth_ll(self, chunk, n_threads=0):
        """Quad-precision ll and the cancellation ratio kappa = d^T C^-1 d / chi2."""
        chunk = _f8(chunk).reshape(-1, 5)
        ll, kappa = np.zeros(chunk.shape[0]), np.zeros(chunk.shape[0])
        rc = self._lib.orc_truth_marginal_ln_likelihood(ctypes.byref(self.spec), _p(chunk),
                                                        chunk.shape[0], _p(ll), _p(kappa),
                                                        int(n_threads))
        if rc != 0:
            raise RuntimeError("truth oracle failed")
        return ll, kappa

    def truth_aA(self, chunk_row, clamp=False):
        L = self.n_linear
        a, A = np.zeros(L), np.zeros((L, L))
        self._lib.orc_truth_posterior_aA(ctypes.byref(self.spec), _p(_f8(chunk_row)),
                                         int(bool(clamp)), _p(a), _p(A))
        return a, A


# ---------------------------------------------------------------------------
# host-side restatements of the accept / iteration logic


def rejection_accept(lls, uu, max_posterior_samples=None):
    """likelihood_helpers.py:107-109 / multiproc_helpers.py:256-258."""
    lls = np.asarray(lls)
    good = np.where(np.exp(lls - lls.max()) > uu)[0]
    if max_posterior_samples is not None:
        good = good[:max_posterior_samples]
    return good


def near_threshold_count(lls, uu, tol=1e-12):
    lls = np.asarray(lls)
    return int(np.sum(np.abs(np.exp(lls - lls.max()) - uu) <= tol))


def batch_tasks_ranges(n_tasks, n_batches, start_idx=0):
    """utils.py:22-72, index form: contiguous split, first n%G batches get +1."""
    out = []
    if n_batches > 0 and n_tasks >= n_batches:
        base, rmdr = divmod(n_tasks, n_batches)
        i1 = start_idx
        for i in range(n_batches):
            i2 = i1 + base + (1 if i < rmdr else 0)
            out.append((i1, i2))
            i1 = i2
    else:
        out.append((start_idx, n_tasks + start_idx))
    return out


def iterative_rejection_indices(ll_fn, n_total, rng, n_requested_samples, init_batch_size=None,
                                growth_factor=128, safety_factor=1, maxiter=128):
    """likelihood_helpers.py:130-229 (safety_factor=1) / multiproc_helpers.py:289-427
    (safety_factor=4): returns (good_idx[:n_requested], all_lls)."""
    n_process = growth_factor * n_requested_samples if init_batch_size is None else init_batch_size
    if n_process > n_total:
        raise ValueError("Prior sample library not big enough!")
    all_lls = np.array([])
    start = 0
    for _ in range(maxiter):
        all_lls = np.concatenate((all_lls, ll_fn(start, start + n_process)))
        uu = rng.uniform(size=len(all_lls))
        good = np.where(np.exp(all_lls - all_lls.max()) > uu)[0]
        if len(good) == 0:
            raise RuntimeError("Failed to find any good samples!")
        if len(good) >= n_requested_samples:
            break
        start += n_process
        n_process = int(safety_factor * (n_requested_samples - len(good)) / len(good) * len(all_lls))
        if start + n_process > n_total:
            n_process = n_total - start
        if n_process <= 0:
            break
    else:
        raise RuntimeError("Hit maximum number of iterations!")
    return good[:n_requested_samples], all_lls
