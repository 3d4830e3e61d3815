"""Shared builders for the test-suite: synthetic stars in the shape of the reference's
fixtures, prior chunks, and the host-emulation loader (tools/host_emulation.cpp)."""
import ctypes
import os
import subprocess

import numpy as np

import thejoker_b200 as tj
from thejoker_b200 import units as u
from thejoker_b200.data_helpers import validate_prepare_data
from thejoker_b200.helper import extract_spec
from thejoker_b200.prior import Normal
from thejoker_b200.synthetic import default_prior_columns, make_data, make_noisy_data

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def default_prior(poly_trend=1, sigma_K0=30.0, P_min=2.0, P_max=1024.0, v0_offsets=None, s=None,
                  pars=None):
    sv = [100 * u.km / u.s] + [0.5 * 50.0 ** (1 - k) * u.km / u.s / u.day**k for k in range(1, 8)]
    sv = sv[:poly_trend]
    if poly_trend == 1:
        sv = sv[0]
    elif poly_trend == 0:
        sv = None
    return tj.JokerPrior.default(P_min=P_min * u.day, P_max=P_max * u.day,
                                 sigma_K0=sigma_K0 * u.km / u.s if sigma_K0 is not None else None,
                                 sigma_v=sv, poly_trend=poly_trend, v0_offsets=v0_offsets, s=s,
                                 pars=pars)


def star_spec(n_times=16, poly_trend=1, seed=42, K=None, sigma=0.5, jitter_mode="apply",
              normal_K=None, n_surveys=1, v1=None, t_span_periods=3.0):
    """(spec dict, data, prior) for a synthetic star."""
    pars = None
    if normal_K is not None:
        pars = {"K": Normal("K", 0.0, normal_K, u.km / u.s)}
    offsets = [Normal(f"dv0_{i}", 0.0, 5.0, u.km / u.s) for i in range(1, n_surveys)]
    prior = default_prior(poly_trend, sigma_K0=None if normal_K is not None else 30.0,
                          v0_offsets=offsets or None, pars=pars)
    if n_surveys == 1:
        data, _ = make_noisy_data(n_times, seed=seed, K=K, sigma=sigma, v1=v1,
                                  t_span_periods=t_span_periods)
    else:
        rng = np.random.default_rng(seed)
        full, _ = make_noisy_data(n_times, seed=seed, K=K, sigma=sigma, v1=v1,
                                  t_span_periods=t_span_periods)
        cuts = np.sort(rng.choice(np.arange(2, n_times - 1), size=n_surveys - 1, replace=False))
        data, lo = [], 0
        for k, hi in enumerate(list(cuts) + [n_times]):
            off = 0.0 if k == 0 else rng.normal(0, 5.0)
            data.append(tj.RVData(full._t_bmjd[lo:hi], (full.rv.value[lo:hi] + off) * u.km / u.s,
                                  full.rv_err[lo:hi]))
            lo = hi
    all_data, ids, trend_M = validate_prepare_data(data, prior.poly_trend, prior.n_offsets)
    return extract_spec(all_data, prior, trend_M, jitter_mode), data, prior


def prior_chunk(n, seed=123, s_lognormal=None, s_const=None):
    cols = list(default_prior_columns(n, seed=seed, s_lognormal=s_lognormal))
    if s_const is not None:
        cols[4] = np.full(n, float(s_const))
    return np.ascontiguousarray(np.stack(cols, axis=1))


_emu = {}


def host_emulation(variant=""):
    """CPU build of the device math (tools/host_emulation.cpp); debug tooling only.
    `variant` is a compile-time flag set of the kernel headers ("" = the shipped
    configuration, "TJB_TRIM=1" = the instruction-trimmed epoch loop, ...)."""
    if variant not in _emu:
        tag = "".join(ch if ch.isalnum() else "_" for ch in variant)
        so = os.path.join(ROOT, "tools", f"libhost_emulation{'_' + tag if tag else ''}.so")
        defs = ["-D" + d for d in variant.split()] if variant else []
        src = os.path.join(ROOT, "tools", "host_emulation.cpp")
        deps = [src] + [os.path.join(ROOT, "thejoker_b200", "csrc", f) for f in
                        ("kepler.cuh", "linalg.cuh", "marginal_ll.cuh", "star_tables.hpp",
                         "accept.cuh", "prior_gen.cuh")]
        if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
            subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared",
                            "-ffp-contract=off"] + defs + [src, "-o", so], check=True)
        lib = ctypes.CDLL(so)
        dp = ctypes.POINTER(ctypes.c_double)
        lib.emu_design_column.argtypes = [ctypes.c_double] * 4 + [dp, ctypes.c_int, dp,
                                                                  ctypes.POINTER(ctypes.c_int)]
        lib.emu_marginal_ll.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_double, dp, dp, dp, dp,
                                        dp, dp, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                        ctypes.c_double, ctypes.c_int, ctypes.c_int, dp,
                                        ctypes.c_long, dp]
        lib.emu_ll_to_key.restype = ctypes.c_longlong
        lib.emu_ll_to_key.argtypes = [ctypes.c_double]
        lib.emu_key_to_ll.restype = ctypes.c_double
        lib.emu_key_to_ll.argtypes = [ctypes.c_longlong]
        lib.emu_pcg64_double.restype = ctypes.c_double
        lib.emu_pcg64_double.argtypes = [ctypes.c_ulonglong] * 5
        lib.emu_sincos_rev.argtypes = [ctypes.c_double, dp, dp]
        up = ctypes.POINTER(ctypes.c_uint)
        lib.emu_philox4x32_10.argtypes = [up, up, up]
        lib.emu_prior_rows.argtypes = [ctypes.POINTER(ctypes.c_int), dp, dp, dp, ctypes.c_ulonglong,
                                       ctypes.c_longlong, ctypes.c_long, dp]
        lib.emu_prior_uniforms.argtypes = [ctypes.c_ulonglong, ctypes.c_ulonglong, ctypes.c_int, dp]
        _emu[variant] = lib
    return _emu[variant]


def emu_marginal_ll(spec, chunk, force_jit=False, variant=""):
    lib = host_emulation(variant)
    dp = ctypes.POINTER(ctypes.c_double)
    p = lambda a: np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(dp)
    chunk = np.ascontiguousarray(chunk, dtype=np.float64)
    ll = np.zeros(len(chunk))
    keep = [np.ascontiguousarray(spec[k], dtype=np.float64) for k in
            ("t", "rv", "ivar", "trend_M", "mu", "Lambda")]
    max_K = spec["max_K"] if np.isfinite(spec["max_K"]) else 1e300
    rc = lib.emu_marginal_ll(spec["n_times"], spec["n_linear"], spec["t0"],
                             *[k.ctypes.data_as(dp) for k in keep], spec["K_prior_kind"],
                             spec["sigma_K0"], spec["P0"], max_K, spec["jitter_mode"],
                             int(force_jit), p(chunk), len(chunk), ll.ctypes.data_as(dp))
    assert rc == 0
    return ll


def emu_prior_rows(gen, index0, n):
    """Host build of csrc/prior_gen.cuh: rows [P, e, omega, M0, s] of the generated prior
    samples index0 .. index0 + n for a ``_lib.TjbPriorGen`` (JokerPrior.device_generator)."""
    lib = host_emulation()
    kind = (ctypes.c_int * 5)(*[gen.par[k].kind for k in range(5)])
    arr = lambda f: (ctypes.c_double * 5)(*[getattr(gen.par[k], f) for k in range(5)])
    rows = np.zeros((int(n), 5))
    lib.emu_prior_rows(kind, arr("p0"), arr("p1"), arr("scale"), gen.seed, int(index0), int(n),
                       rows.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
    return rows


def rel_err(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)


def reference_gate(got, ref, truth, tol=1e-10, label=""):
    """The north-star gate against the reference, with its escape hatch made explicit and
    *bounded*: every ll must be within `tol` of the quad-precision truth; against the
    reference it must be within `tol` too, except where the reference's own double
    arithmetic is off the truth by at least 90 % of the difference (its N x N chi2 form
    cancels on flat / high-S/N data) -- that set is counted, printed and must consist of
    samples where the reference misses the truth by more than tol/2.
    Returns the report dict."""
    r_ref, r_truth, ref_truth = rel_err(got, ref), rel_err(got, truth), rel_err(ref, truth)
    over = r_ref > tol
    explained = over & (ref_truth >= 0.9 * r_ref) & (ref_truth > 0.5 * tol)
    rep = dict(n=len(got), max_rel_vs_ref=float(r_ref.max()), max_rel_vs_truth=float(r_truth.max()),
               ref_vs_truth_max=float(ref_truth.max()), n_over_tol_vs_ref=int(over.sum()),
               n_reference_off_truth=int(explained.sum()),
               n_unexplained=int((over & ~explained).sum()))
    print(f"\n[reference gate {label}] " + ", ".join(f"{k}={v:.3g}" if isinstance(v, float)
                                                      else f"{k}={v}" for k, v in rep.items()))
    assert rep["max_rel_vs_truth"] <= tol, rep
    assert rep["n_unexplained"] == 0, rep
    return rep


def accept_sets_match(got_idx, ll, uu, max_keep=None, tol=1e-12, ll_other=None, llmax=None,
                      index_base=0):
    """The accepted index set against ``where(exp(ll - max) > u)`` computed on the host
    from `ll` (likelihood_helpers.py:107-109): identical except for samples within `tol`
    of the threshold, which are returned as a count (BASELINE.json north_star).  If the
    lls the device used (`ll_other`) are given, a sample is also excused when its
    threshold u lies between exp(ll - max) and exp(ll_other - max), i.e. when the (<= 1e-10
    relative) difference of the two implementations' ll moves it across; those are counted
    separately.  The comparison is made on the sets minus those indices -- never skipped.
    `llmax` overrides ll.max() (a slice of a larger run); `index_base` is subtracted from
    got_idx.  Returns (n_near, n_moved)."""
    m = ll.max() if llmax is None else llmax
    a = np.exp(ll - m)
    want = np.where(a > uu)[0]
    near = np.abs(a - uu) <= tol
    moved = np.zeros(len(ll), dtype=bool)
    if ll_other is not None:
        b = np.exp(ll_other - m)
        near |= np.abs(b - uu) <= tol
        moved = ((a > uu) != (b > uu)) & ~near
    got_idx = np.asarray(got_idx) - index_base
    if max_keep is not None:
        # truncated lists: compare the common prefix range only
        hi = min(got_idx[-1] if len(got_idx) else -1, want[:max_keep][-1] if len(want) else -1)
        want, got_idx = want[want <= hi], got_idx[got_idx <= hi]
    diff = np.setxor1d(got_idx, want)
    ok = near[diff] | moved[diff]
    assert ok.all(), (diff[~ok][:10], len(diff))
    return int(near.sum()), int(moved[diff].sum())


def mode_chunk(spec, n=400, sigma=0.5, seed=0):
    """Prior rows scattered tightly around the true orbit of the synthetic star (the
    posterior mode): the ill-conditioned regime where chi2 << y^T C^-1 y."""
    P0 = 51.8239
    M0p = 2.592 - 2 * np.pi * (spec["t0"] - 51544.5) / P0
    rng = np.random.default_rng(seed)
    sc = sigma / 0.5
    chunk = np.zeros((n, 5))
    chunk[:, 0] = P0 * (1 + rng.normal(0, 1e-5 * sc, n))
    chunk[:, 1] = 0.3 + rng.normal(0, 2e-3 * sc, n)
    chunk[:, 2] = 0.283 + rng.normal(0, 5e-3 * sc, n)
    chunk[:, 3] = M0p + rng.normal(0, 5e-3 * sc, n)
    return chunk


# ---- stand-ins for the native multi-star call (host-side tests without a GPU) -----------
def fake_multistar_stars(n_stars):
    """(prior with one survey offset, n_stars two-survey stars with 10..16 epochs)."""
    import thejoker_b200 as tj
    from thejoker_b200 import units as u
    from thejoker_b200.prior import Normal
    from thejoker_b200.synthetic import make_noisy_data

    prior = default_prior(1, sigma_K0=25.0, v0_offsets=[Normal("dv0_1", 0.0, 5.0, u.km / u.s)])
    stars = []
    for i in range(n_stars):
        full, _ = make_noisy_data(10 + i % 7, seed=i)
        stars.append([tj.RVData(full._t_bmjd[:4], full.rv[:4], full.rv_err[:4]),
                      tj.RVData(full._t_bmjd[4:], full.rv[4:], full.rv_err[4:])])
    return prior, stars


def fake_multistar_lib(n_prior, keep, n_per, L):
    """An object with a ``tjb_multistar_rejection`` that echoes what it is given: star j
    'accepts' k = n_times % (keep + 1) samples whose rows carry n_times in the nonlinear
    columns and the star's pre-drawn normals in the linear ones; ll = 0..k-1; ll_max is
    derived from the star's generator state.  Returns (lib, list of stars per call)."""
    import ctypes

    calls = []

    class FakeLib:
        @staticmethod
        def tjb_multistar_rejection(device, jobref):
            job = jobref._obj
            n = job.n_stars
            calls.append(n)
            assert job.n_prior == n_prior and job.max_keep == keep and job.n_per == n_per
            shape = lambda p, sh, t: np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(t)), sh)
            counts = shape(job.h_counts, (n, 3), ctypes.c_int64)
            rows = shape(job.h_rows, (n, keep * n_per, 5 + L), ctypes.c_double)
            nrm = shape(job.h_normals, (n, keep, n_per, L), ctypes.c_double)
            ll = shape(job.h_ll, (n, keep), ctypes.c_double)
            llmax = shape(job.h_llmax, (n,), ctypes.c_double)
            for j in range(n):
                sp = job.specs[j]
                k = sp.n_times % (keep + 1)
                counts[j] = (k + 10, k, 1)
                rows[j, : k * n_per, :5] = sp.n_times
                rows[j, : k * n_per, 5:] = nrm[j, :k].reshape(k * n_per, L)
                ll[j, :k] = np.arange(k)
                llmax[j] = job.pcg[j].state_lo % 1000
            return 0

    return FakeLib, calls


def fake_multistar_joker(prior, n_prior, L, **kw):
    """A MultiStarJoker whose device state is filled in by hand (no CUDA)."""
    import types

    import thejoker_b200 as tj

    ms = tj.MultiStarJoker(prior, None, devices=[0], streams_per_device=2, **kw)
    col = types.SimpleNamespace(data_ptr=lambda: 8)
    ms._dev = {0: dict(cols=[col] * 4, s=None, slots=[])}
    ms._host_cols = [np.zeros(n_prior)] * 5
    ms._s_const = 0.0
    ms._helper0 = types.SimpleNamespace(n_linear=L)
    return ms
