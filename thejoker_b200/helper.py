"""CJokerHelper: the reference's operator boundary
(``cdef class CJokerHelper``, thejoker/src/fast_likelihood.pyx:70-576) on top of
libthejoker_b200.so.

Same constructor ``(data, prior, trend_M)``, same attributes (``internal_units``,
``packed_order``, ``data``, ``prior``, ``a``, ``A``, ``b``) and methods
(``batch_marginal_ln_likelihood``, ``batch_get_posterior_samples``,
``test_likelihood_worker``, ``__reduce__``).  The N x N workspaces ``B`` / ``Binv``
are never formed on the GPU path; only the CPU oracle exposes
them.  In addition the helper has a device-resident API (torch tensors used as raw
device buffers) that the drivers in thejoker.py use to keep prior shards, ll and
accepted indices on the GPU.
"""
from __future__ import annotations

import ctypes
from collections import OrderedDict

import numpy as np

from . import _lib
from . import units as u
from .prior import FixedCompanionMass
from .samples import _nonlinear_internal_units, _nonlinear_packed_order

__all__ = ["CJokerHelper", "extract_spec"]

JITTER_MODES = {"apply": 1, "reference": 0, 1: 1, 0: 0, True: 1, False: 0}


def extract_spec(data, prior, trend_M, jitter_mode="apply"):
    """What CJokerHelper.__init__ pulls out of (data, prior, trend_M)
    (fast_likelihood.pyx:125-253), as plain float64 arrays / scalars.

    Differences from the reference, all on paths its tests never exercise
    (SURVEY.md section 0.6): ``P0`` is always converted to days (pyx:240-241 converts
    to the prior's P unit while the chunk P is in days); a custom Normal K prior
    together with ``n_offsets > 0`` stores its variance in ``Lambda[0]`` (pyx:249-252
    writes it to ``Lambda[n_offsets]``).
    """
    rv_unit = data.rv.unit
    internal_units = OrderedDict()
    for k in ("P", "e", "omega", "M0"):
        internal_units[k] = _nonlinear_internal_units[k]
    internal_units["s"] = rv_unit
    internal_units["K"] = rv_unit
    internal_units["v0"] = rv_unit
    for offset in prior.v0_offsets:  # offsets sit between v0 and v1 (pyx:143-145)
        internal_units[offset.name] = rv_unit
    for i, name in enumerate(prior._v_trend_names):
        internal_units[name] = rv_unit / u.day**i

    n_times = len(data)
    n_poly = prior.poly_trend
    n_offsets = prior.n_offsets
    n_linear = 1 + n_poly + n_offsets
    if data._has_cov:
        raise NotImplementedError("a full covariance matrix cannot be used on the hot path "
                                  "(the reference requires a 1-D ivar, pyx:83, 165-166)")
    trend_M = np.ascontiguousarray(trend_M, dtype=np.float64)
    if trend_M.ndim != 2 or trend_M.shape[0] != n_times or trend_M.shape[1] != n_linear - 1:
        raise ValueError("Invalid design matrix shape: {}, expected: {}".format(
            trend_M.shape, (n_times, n_linear - 1)))  # pyx:174-179
    if n_linear > _lib.TJB_MAX_LINEAR:
        raise ValueError(f"n_linear = {n_linear} exceeds the supported maximum "
                         f"{_lib.TJB_MAX_LINEAR}")

    mu = np.zeros(n_linear)
    Lambda = np.zeros(n_linear)
    for i in range(n_offsets):  # pyx:209-220
        name = prior.v0_offsets[i].name
        m, s = prior.pars[name].mean_std(internal_units[name])
        mu[2 + i] = m
        Lambda[2 + i] = s**2

    Kdist = prior.pars["K"]
    fixed_K_prior = 0 if isinstance(Kdist, FixedCompanionMass) else 1  # pyx:225-228
    sigma_K0 = P0 = max_K = 0.0
    for i, name in enumerate(prior._linear_equiv_units.keys()):  # pyx:230-252
        dist = prior.pars[name]
        to_unit = internal_units[name]
        m, s = dist.mean_std(to_unit)
        if name == "K" and fixed_K_prior == 0:
            sigma_K0 = float(dist._sigma_K0.to_value(to_unit))
            P0 = float(dist._P0.to_value(u.day))
            max_K = float(dist._max_K.to_value(to_unit))
            mu[i] = m
        elif name == "K":
            Lambda[0] = s**2
            mu[0] = m
        elif name == "v0":
            Lambda[i] = s**2
            mu[i] = m
        else:  # v1, v2, ...
            j = i + n_offsets
            Lambda[j] = s**2
            mu[j] = m
    if fixed_K_prior == 1:
        P0, max_K = 1.0, np.inf

    return dict(
        n_times=n_times, n_poly=n_poly, n_offsets=n_offsets, n_linear=n_linear,
        n_pars=len(prior.par_names),
        t0=float(data._t_ref_bmjd),
        t=np.ascontiguousarray(data._t_bmjd, dtype="f8"),
        rv=np.ascontiguousarray(data.rv.value, dtype="f8"),
        ivar=np.ascontiguousarray(data.ivar.to_value(u.one / rv_unit**2), dtype="f8"),
        trend_M=trend_M, mu=mu, Lambda=Lambda, K_prior_kind=fixed_K_prior,
        sigma_K0=sigma_K0, P0=P0, max_K=max_K,
        jitter_mode=JITTER_MODES[jitter_mode], internal_units=internal_units,
    )


def _vp(arr):
    return ctypes.c_void_p(arr.ctypes.data)


def _pcg_struct(rng):
    """numpy Generator(PCG64) state -> TjbPcg64, or None for other bit generators."""
    bg = rng.bit_generator
    if type(bg).__name__ != "PCG64":
        return None
    st = bg.state["state"]
    m = (1 << 64) - 1
    return _lib.TjbPcg64(st["state"] >> 64, st["state"] & m, st["inc"] >> 64, st["inc"] & m)


def _rebuild_from_spec(spec, device):
    return CJokerHelper.from_spec(spec, device)


class CJokerHelper:
    """GPU-backed stand-in for thejoker.src.fast_likelihood.CJokerHelper.

    Parameters
    ----------
    data : RVData
    prior : JokerPrior
    trend_M : float64 array (n_times, n_linear - 1), C-contiguous
    device : int, CUDA device index (default 0, or LOCAL_RANK under torchrun)
    jitter_mode : "apply" (default) -- the per-sample jitter enters the covariance,
        ivar/(1 + s^2 ivar); "reference" -- ignore s, exactly as the reference's
        Cython does (fast_likelihood.pyx:458 stores the result in a buffer nothing
        reads).  The two coincide for s = 0.
    """

    def __init__(self, data, prior, trend_M, device=None, jitter_mode="apply"):
        self.prior = prior
        self.data = data
        self._trend_M = np.ascontiguousarray(trend_M, dtype=np.float64)
        self._jitter_mode = jitter_mode
        self._init_from_spec(extract_spec(data, prior, self._trend_M, jitter_mode), device)

    @classmethod
    def from_spec(cls, spec, device=None):
        """Build directly from the plain arrays ``extract_spec`` returns (keys t, rv,
        ivar, t0, trend_M, mu, Lambda, K_prior_kind, sigma_K0, P0, max_K, jitter_mode),
        without RVData / JokerPrior objects -- the entry point for callers that already
        hold a star table (tests, the multi-star driver)."""
        self = cls.__new__(cls)
        self.prior = self.data = None
        spec = dict(spec)
        spec["t"], spec["rv"], spec["ivar"] = (np.ascontiguousarray(spec[k], dtype="f8")
                                               for k in ("t", "rv", "ivar"))
        n = len(spec["t"])
        spec["trend_M"] = np.ascontiguousarray(spec["trend_M"], dtype="f8").reshape(n, -1)
        spec["mu"] = np.ascontiguousarray(spec["mu"], dtype="f8")
        spec["Lambda"] = np.ascontiguousarray(spec["Lambda"], dtype="f8")
        spec.setdefault("n_times", n)
        spec.setdefault("n_linear", 1 + spec["trend_M"].shape[1])
        spec.setdefault("n_pars", 5 + spec["n_linear"])
        spec.setdefault("internal_units", OrderedDict())
        spec["jitter_mode"] = JITTER_MODES[spec.get("jitter_mode", 1)]
        self._trend_M = spec["trend_M"]
        self._jitter_mode = spec["jitter_mode"]
        self._init_from_spec(spec, device)
        return self

    def _init_from_spec(self, spec, device):
        self.spec = spec
        self.internal_units = self.spec["internal_units"]
        self.packed_order = list(_nonlinear_packed_order)
        self.n_times, self.n_linear = self.spec["n_times"], self.spec["n_linear"]
        self.n_pars = self.spec["n_pars"]
        if self.n_linear > _lib.TJB_MAX_LINEAR or len(spec["mu"]) < self.n_linear:
            raise ValueError("bad n_linear / mu length")
        self.a = self.A = self.Ainv = self.b = None
        self.last_nonfinite = 0  # NaN / inf lls seen by the last accept call
        if device is None:
            import os
            device = int(os.environ.get("LOCAL_RANK", "0"))
        self.device = int(device)

        lib = self._lib = _lib.load()
        h = ctypes.c_void_p()
        _lib.check(lib.tjb_create(ctypes.byref(self._c_spec(self.spec)), self.device, ctypes.byref(h)))
        self._h = h

    @staticmethod
    def _c_spec(sp):
        cs = _lib.TjbSpec()
        cs.n_times, cs.n_linear, cs.t_ref = sp["n_times"], sp["n_linear"], float(sp["t0"])
        dp = ctypes.POINTER(ctypes.c_double)
        cs.t, cs.rv, cs.ivar = (sp[k].ctypes.data_as(dp) for k in ("t", "rv", "ivar"))
        cs.trend_M = sp["trend_M"].ctypes.data_as(dp)
        for i in range(sp["n_linear"]):
            cs.mu[i], cs.Lambda[i] = sp["mu"][i], sp["Lambda"][i]
        cs.K_prior_kind = int(sp["K_prior_kind"])
        cs.sigma_K0, cs.P0 = float(sp["sigma_K0"]), float(sp["P0"])
        cs.max_K = float(sp["max_K"]) if np.isfinite(sp["max_K"]) else 1e300
        cs.jitter_mode = sp["jitter_mode"]
        return cs

    def update_star(self, data, prior, trend_M):
        """Point this helper (and its device buffers) at another star: the multi-star
        replacement for constructing a new CJokerHelper per star (thejoker.py:87-91)."""
        self.prior, self.data = prior, data
        self._trend_M = np.ascontiguousarray(trend_M, dtype=np.float64)
        spec = extract_spec(data, prior, self._trend_M, self._jitter_mode)
        self.spec = spec
        self.internal_units = spec["internal_units"]
        self.n_times, self.n_linear, self.n_pars = spec["n_times"], spec["n_linear"], spec["n_pars"]
        self.a = self.A = self.Ainv = self.b = None
        _lib.check(self._lib.tjb_update_star(self._h, ctypes.byref(self._c_spec(spec))))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                self._lib.tjb_destroy(h)
            except Exception:
                pass
            self._h = None

    def __reduce__(self):  # pyx:122-123
        if self.data is None:
            return (_rebuild_from_spec, (self.spec, self.device))
        return (CJokerHelper, (self.data, self.prior, np.array(self._trend_M), self.device,
                               self._jitter_mode))

    # ------------------------------------------------------------------
    # reference method surface (host buffers in, host buffers out)

    def batch_marginal_ln_likelihood(self, chunk):
        """ll for a (n, 5) chunk [P, e, omega, M0, s] in internal units (pyx:428-469).
        Host array in, host array out; copies and kernel are pipelined inside the
        library (tjb_marginal_ll_host)."""
        chunk = np.asarray(chunk)
        if chunk.ndim != 2 or chunk.shape[1] != 5:
            raise ValueError("chunk must have shape (n_samples, 5)")
        if chunk.dtype != np.float64 or not chunk.flags.c_contiguous:
            raise ValueError("Buffer dtype mismatch / not C-contiguous: expected float64[:, ::1]")
        n = chunk.shape[0]
        # pyx:443 pre-fills with NaN; every element is written on success and a failure
        # raises, so the fill pass over a multi-GB array is skipped
        ll = np.empty(n)
        _lib.check(self._lib.tjb_marginal_ll_host(self._h, _vp(chunk), n, _vp(ll)))
        return ll

    def marginal_ln_likelihood_columns(self, P, e, omega, M0, s=None, s_const=0.0, out=None):
        """ll for prior samples held as separate host columns in internal units (the
        layout of a JokerSamples) -- no (n, 5) packing, and only the columns that vary
        cross PCIe.  Host arrays in, host array out."""
        cols = [np.ascontiguousarray(c, dtype=np.float64) for c in (P, e, omega, M0)]
        n = len(cols[0])
        if any(len(c) != n for c in cols):
            raise ValueError("prior columns differ in length")
        if s is not None:
            s = np.ascontiguousarray(s, dtype=np.float64)
            if len(s) != n:
                raise ValueError("prior columns differ in length")
        ll = np.empty(n) if out is None else out
        _lib.check(self._lib.tjb_marginal_ll_host_soa(
            self._h, *[_vp(c) for c in cols], _vp(s) if s is not None else None, float(s_const),
            n, _vp(ll)))
        return ll

    def posterior_aA(self, chunk, clamp_K=False):
        """(ll, a, A) per row: posterior mean / covariance of the linear parameters
        (pyx:394-423, 530).  clamp_K=False reproduces the reference, which does not
        clamp Lambda_K in these entry points (pyx:519-521, 569-571)."""
        chunk = np.ascontiguousarray(chunk, dtype=np.float64).reshape(-1, 5)
        k, L = chunk.shape[0], self.n_linear
        ll, a, A = np.zeros(k), np.zeros((k, L)), np.zeros((k, L, L))
        _lib.check(self._lib.tjb_posterior_aA(self._h, _vp(chunk), k, int(bool(clamp_K)), _vp(ll),
                                              _vp(a), _vp(A)))
        return ll, a, A

    def batch_get_posterior_samples(self, chunk, n_linear_samples_per, rng, draw="auto",
                                    clamp_K=False):
        """pyx:471-545.  Returns (samples[n*k, 5+L], ll[n*k]).

        draw="auto" (default): "numpy" for up to 4096 rows -- the reference's numbers --
        and "device" beyond, where a Python call per row would dominate.
        draw="device": standard normals are taken from ``rng`` in the order the
        reference's ``rng.multivariate_normal`` consumes them ((k, L) per row) and the
        GPU forms x = a + F z with F F^T = A.  Same distribution and same stream
        consumption as the reference, different numbers (numpy factors A by SVD).
        draw="numpy": (a, A) from the GPU, then exactly the reference's call
        ``rng.multivariate_normal(a, A, size=k)`` per row (pyx:529-530).
        """
        chunk = np.ascontiguousarray(chunk, dtype=np.float64).reshape(-1, 5)
        n, L, k = chunk.shape[0], self.n_linear, int(n_linear_samples_per)
        if draw == "auto":
            draw = "numpy" if n <= 4096 else "device"
        if draw == "numpy":
            lls, a, A = self.posterior_aA(chunk, clamp_K)
            out = np.zeros((n, k, 5 + L))
            for i in range(n):
                out[i, :, :5] = chunk[i]
                # same numbers as the reference's call (pyx:529-530); only numpy's
                # positive-semidefiniteness warning pass (an allclose per row) is skipped
                out[i, :, 5:] = rng.multivariate_normal(a[i], A[i], size=k, check_valid="ignore")
            return out.reshape(n * k, -1), np.repeat(lls, k)
        if draw != "device":
            raise ValueError("draw must be 'device' or 'numpy'")
        normals = rng.standard_normal((n, k, L))
        out = np.zeros((n * k, 5 + L))
        lls = np.zeros(n)
        if n:
            _lib.check(self._lib.tjb_posterior_draw(self._h, _vp(chunk), n, k, int(bool(clamp_K)),
                                                    _vp(normals), _vp(out), _vp(lls)))
        return out, np.repeat(lls, k)

    def ln_unmarginalized_likelihood(self, rows):
        """ln of the un-marginalised likelihood of full posterior samples, on the GPU
        (samples.py:611-632): ``rows`` is the packed (k, 5 + n_linear) array
        ``batch_get_posterior_samples`` returns, [P, e, omega, M0, s, K, v0, (offsets),
        v1, ...] in internal units.  Survey offsets are part of the model here (the
        reference's version builds the orbit from K and the polynomial trend only)."""
        rows = np.ascontiguousarray(rows, dtype=np.float64)
        if rows.ndim != 2 or rows.shape[1] != 5 + self.n_linear:
            raise ValueError(f"rows must have shape (k, {5 + self.n_linear})")
        ll = np.full(rows.shape[0], np.nan)
        self._sync_stream()
        _lib.check(self._lib.tjb_unmarginalized_ll(self._h, _vp(rows), rows.shape[0], _vp(ll)))
        return ll

    def test_likelihood_worker(self, chunk_row):
        """pyx:547-576: ll for one row; leaves a, A, Ainv, b readable (the N x N B and
        Binv of the reference are never formed here)."""
        row = np.ascontiguousarray(chunk_row, dtype=np.float64).reshape(5)
        ll, a, A = self.posterior_aA(row[None, :], clamp_K=False)
        self.a, self.A = a[0], A[0]
        self.Ainv = np.linalg.inv(self.A)  # L x L; the kernel holds it only in factored form
        M = np.hstack((self.design_column(row)[:, None], self.spec["trend_M"]))
        self.b = M @ self.spec["mu"]  # pyx:306-309
        return float(ll[0])

    def design_column(self, chunk_row, return_stats=False):
        """Row 0 of M_T for one sample: z_n = cos(f_n + omega) + e cos(omega) (pyx:453-455)."""
        row = np.ascontiguousarray(chunk_row, dtype=np.float64).reshape(5)
        z = np.zeros(self.n_times)
        st = np.zeros(3, dtype=np.int32)
        _lib.check(self._lib.tjb_design_column(self._h, _vp(row), _vp(z), _vp(st)))
        return (z, st) if return_stats else z

    # ------------------------------------------------------------------
    # device-resident API (torch tensors are raw device buffers)

    def _sync_stream(self):
        import torch

        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self._lib.tjb_set_stream(self._h, ctypes.c_void_p(st)))

    @staticmethod
    def _ptr(t):
        return ctypes.c_void_p(t.data_ptr()) if t is not None else None

    def _check_dev(self, t, dtype, name):
        import torch

        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.device.index == self.device
                and t.dtype == dtype and t.is_contiguous()):
            raise ValueError(f"{name} must be a contiguous {dtype} CUDA tensor on device "
                             f"{self.device}")

    def new_llmax_key(self):
        """int64[1] device scalar holding the order-preserving key of the running max."""
        import torch

        key = torch.empty(1, dtype=torch.int64, device=f"cuda:{self.device}")
        self._sync_stream()
        _lib.check(self._lib.tjb_llmax_reset(self._h, self._ptr(key)))
        return key

    def set_peer_keys(self, keys):
        """Max-keys on other GPUs (int64[1] CUDA tensors) that every likelihood launch of
        this helper also updates, through NVLink peer stores (tjb_set_peer_keys)."""
        n = len(keys)
        ptrs = (ctypes.c_void_p * max(n, 1))(*[k.data_ptr() for k in keys])
        devs = (ctypes.c_int * max(n, 1))(*[k.device.index for k in keys])
        _lib.check(self._lib.tjb_set_peer_keys(self._h, ptrs, devs, n))
        self._peer_keys = list(keys)  # keep the tensors alive

    def llmax_value(self, key):
        out = ctypes.c_double()
        self._sync_stream()
        _lib.check(self._lib.tjb_llmax_get(self._h, self._ptr(key), ctypes.byref(out)))
        return out.value

    def marginal_ll_soa(self, P, e, omega, M0, s=None, s_const=0.0, out=None, llmax_key=None):
        """ll for SoA device columns.  ``s=None`` means every sample has jitter
        ``s_const``.  Asynchronous on torch's current stream."""
        import torch

        n = P.numel()
        for name, t in (("P", P), ("e", e), ("omega", omega), ("M0", M0)):
            self._check_dev(t, torch.float64, name)
            if t.numel() != n:
                raise ValueError("prior columns differ in length")
        if s is not None:
            self._check_dev(s, torch.float64, "s")
        if out is None:
            out = torch.empty(n, dtype=torch.float64, device=P.device)
        self._check_dev(out, torch.float64, "out")
        self._sync_stream()
        _lib.check(self._lib.tjb_marginal_ll_soa(self._h, self._ptr(P), self._ptr(e),
                                                 self._ptr(omega), self._ptr(M0), self._ptr(s),
                                                 float(s_const), n, self._ptr(out),
                                                 self._ptr(llmax_key)))
        return out

    def marginal_ll_generated(self, gen, index0, n, out=None, llmax_key=None):
        """ll of the prior samples with global indices [index0, index0 + n) of the
        counter-based generator ``gen`` (JokerPrior.device_generator): the likelihood
        kernel draws each sample in registers, no prior array exists.  Asynchronous on
        torch's current stream."""
        import torch

        if out is None:
            out = torch.empty(int(n), dtype=torch.float64, device=f"cuda:{self.device}")
        self._check_dev(out, torch.float64, "out")
        if out.numel() != int(n):
            raise ValueError("out has the wrong length")
        self._sync_stream()
        _lib.check(self._lib.tjb_marginal_ll_generated(self._h, ctypes.byref(gen), int(index0),
                                                       int(n), self._ptr(out),
                                                       self._ptr(llmax_key)))
        return out

    def prior_rows(self, gen, idx):
        """Packed (k, 5) host rows [P, e, omega, M0, s] (internal units) of the generated
        prior samples with the given global indices."""
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        rows = np.empty((len(idx), 5))
        if len(idx):
            self._sync_stream()
            _lib.check(self._lib.tjb_prior_rows(self._h, ctypes.byref(gen), _vp(idx), len(idx),
                                                _vp(rows)))
        return rows

    def marginal_ll_host_columns(self, P, e, omega, M0, s=None, s_const=0.0, out=None,
                                 llmax_key=None):
        """ll for HOST columns (numpy / memory-mapped, internal units), left on the device
        in ``out`` with the running max in ``llmax_key``: the columns are streamed through
        the GPU in slices (pageable memory via a page-locked ring) and never become
        resident.  Blocks until done (the GIL is released)."""
        import torch

        cols = [np.ascontiguousarray(c, dtype=np.float64) for c in (P, e, omega, M0)]
        n = len(cols[0])
        if any(len(c) != n for c in cols):
            raise ValueError("prior columns differ in length")
        if s is not None:
            s = np.ascontiguousarray(s, dtype=np.float64)
            if len(s) != n:
                raise ValueError("prior columns differ in length")
        if out is None:
            out = torch.empty(n, dtype=torch.float64, device=f"cuda:{self.device}")
        self._check_dev(out, torch.float64, "out")
        if out.numel() != n:
            raise ValueError("out has the wrong length")
        self._sync_stream()
        _lib.check(self._lib.tjb_marginal_ll_host_soa_resident(
            self._h, *[_vp(c) for c in cols], _vp(s) if s is not None else None, float(s_const),
            n, self._ptr(out), self._ptr(llmax_key)))
        return out

    def marginal_ll_aos(self, chunk, uniform_s=False, out=None, llmax_key=None):
        import torch

        self._check_dev(chunk, torch.float64, "chunk")
        if chunk.ndim != 2 or chunk.shape[1] != 5:
            raise ValueError("chunk must have shape (n, 5)")
        n = chunk.shape[0]
        if out is None:
            out = torch.empty(n, dtype=torch.float64, device=chunk.device)
        self._sync_stream()
        _lib.check(self._lib.tjb_marginal_ll_aos(self._h, self._ptr(chunk), int(bool(uniform_s)), n,
                                                 self._ptr(out), self._ptr(llmax_key)))
        return out

    def llmax_update(self, ll, key):
        self._sync_stream()
        _lib.check(self._lib.tjb_llmax_update(self._h, self._ptr(ll), ll.numel(), self._ptr(key)))

    def accept(self, ll, llmax_key, uniforms=None, rng=None, rng_offset=0, index_base=0,
               max_keep=None, near_tol=1e-12):
        """good = where(exp(ll - max) > u)[0][:max_keep] on the device
        (likelihood_helpers.py:107-109).

        Either ``uniforms`` (device float64[n]) or ``rng`` (numpy Generator on PCG64,
        *not advanced here*: sample i uses the (rng_offset+i)-th double the generator
        would return).  Returns (idx int64 device tensor, n_accepted_total, n_near).
        """
        import torch

        n = ll.numel()
        max_keep = n if max_keep is None else int(max_keep)
        idx = torch.empty(max(min(max_keep, n), 1), dtype=torch.int64, device=ll.device)
        counts = (ctypes.c_int64 * 3)()
        pcg = None
        if uniforms is None:
            pcg = _pcg_struct(rng)
            if pcg is None:
                raise ValueError("device-side uniforms need a numpy Generator on PCG64")
        self._sync_stream()
        _lib.check(self._lib.tjb_accept(self._h, self._ptr(ll), n, self._ptr(llmax_key),
                                        self._ptr(uniforms),
                                        ctypes.byref(pcg) if pcg is not None else None,
                                        int(rng_offset), int(index_base), min(max_keep, n),
                                        float(near_tol), self._ptr(idx), counts))
        self.last_nonfinite = self._accept_nonfinite()
        return idx[: counts[1]], int(counts[0]), int(counts[2])

    def _accept_nonfinite(self):
        """NaN / +-inf lls seen by the last accept call (tjb_accept_nonfinite)."""
        c = ctypes.c_int64()
        _lib.check(self._lib.tjb_accept_nonfinite(self._h, ctypes.byref(c)))
        return int(c.value)

    def accept_dist(self, comm, ll, llmax_key, global_offset, uniforms=None, rng=None,
                    max_keep=None, n_global=None, near_tol=1e-12):
        """The accept step over shards owned by different ranks (tjb_accept_dist): NCCL MAX
        all-reduce of the key, local flag / scan / scatter with globally addressed
        uniforms, all-gather of counts and indices -- all inside the library.  Collective:
        every rank of ``comm`` (a ``sharding.LibComm``) calls it.  ``ll`` is this rank's
        shard (global samples global_offset ...).  Returns (global idx int64 device tensor,
        n_accepted_total, n_near), identical on every rank."""
        import torch

        n = ll.numel()
        n_global = n if n_global is None else int(n_global)
        max_keep = n_global if max_keep is None else min(int(max_keep), n_global)
        idx = torch.empty(max(max_keep, 1), dtype=torch.int64, device=ll.device)
        counts = (ctypes.c_int64 * 3)()
        pcg = None
        if uniforms is None:
            pcg = _pcg_struct(rng)
            if pcg is None:
                raise ValueError("device-side uniforms need a numpy Generator on PCG64")
        self._sync_stream()
        _lib.check(self._lib.tjb_accept_dist(self._h, comm.handle, self._ptr(ll), n,
                                             self._ptr(llmax_key), self._ptr(uniforms),
                                             ctypes.byref(pcg) if pcg is not None else None,
                                             int(global_offset), max_keep, float(near_tol),
                                             self._ptr(idx), counts))
        self.last_nonfinite = self._accept_nonfinite()  # summed over the ranks
        return idx[: counts[1]], int(counts[0]), int(counts[2])

    def allreduce_max_key(self, comm, llmax_key):
        """Integer MAX all-reduce of the key over the ranks of ``comm`` (library NCCL)."""
        self._sync_stream()
        _lib.check(self._lib.tjb_comm_allreduce_max_key(self._h, comm.handle,
                                                        self._ptr(llmax_key)))
        return llmax_key

    def pcg64_uniform(self, rng, n, offset=0):
        """The doubles ``rng.random(n)`` would return (without advancing rng), on device."""
        import torch

        out = torch.empty(n, dtype=torch.float64, device=f"cuda:{self.device}")
        pcg = _pcg_struct(rng)
        self._sync_stream()
        _lib.check(self._lib.tjb_pcg64_uniform(self._h, ctypes.byref(pcg), int(offset), n,
                                               self._ptr(out)))
        return out

    def solver_stats(self, reset=False):
        """Run-wide counters of the Kepler solver's rare path since the last reset:
        extra FP64 passes (lane-epochs) and non-converged epochs."""
        out = (ctypes.c_uint64 * 4)()
        _lib.check(self._lib.tjb_get_stats(self._h, out, int(bool(reset))))
        return dict(extra_fp64_passes=int(out[0]), not_converged=int(out[1]))

    def fp64_peak(self, iters=20000):
        tf, ms = ctypes.c_double(), ctypes.c_double()
        self._sync_stream()
        _lib.check(self._lib.tjb_fp64_peak(self._h, int(iters), ctypes.byref(tf), ctypes.byref(ms)))
        return tf.value, ms.value

    def device_info(self):
        v = [ctypes.c_int() for _ in range(4)]
        _lib.check(self._lib.tjb_device_info(self._h, *[ctypes.byref(x) for x in v]))
        return dict(n_sm=v[0].value, ctas_per_sm=v[1].value, cc=(v[2].value, v[3].value))


def prior_sample_device(gen, index0, n, device, with_s=True):
    """Columns [P, e, omega, M0(, s)] of the generated prior samples with global indices
    [index0, index0 + n) as float64 CUDA tensors on ``device`` (tjb_prior_sample; no star
    handle needed).  Asynchronous on torch's current stream."""
    import torch

    lib = _lib.load()
    with torch.cuda.device(device):
        cols = [torch.empty(int(n), dtype=torch.float64, device=f"cuda:{device}")
                for _ in range(5 if with_s else 4)]
        st = torch.cuda.current_stream(device).cuda_stream
        ptrs = [ctypes.c_void_p(c.data_ptr()) for c in cols] + ([] if with_s else [None])
        _lib.check(lib.tjb_prior_sample(int(device), ctypes.c_void_p(st), ctypes.byref(gen),
                                        int(index0), int(n), *ptrs))
    return cols
