"""Host-side logic (no GPU): units, RVData, JokerPrior, JokerSamples, design matrices,
the sharding rule, spec extraction, the C-ABI library's exported symbols, and the
device math compiled for the host (tools/host_emulation.cpp) against the oracle."""
import ctypes
import os
import sys
import re

import numpy as np
import pytest
from helpers import (ROOT, default_prior, emu_marginal_ll, host_emulation, prior_chunk, rel_err,
                     star_spec)

import thejoker_b200 as tj
from thejoker_b200 import _lib
from thejoker_b200 import units as u
from thejoker_b200.data_helpers import validate_prepare_data
from thejoker_b200.likelihood_helpers import (get_constant_term_design_matrix,
                                              get_trend_design_matrix)
from thejoker_b200.prior import Normal
from thejoker_b200.sharding import batch_tasks, merge_accepted, shard_ranges
from thejoker_b200.synthetic import make_data


# ---- units -------------------------------------------------------------------
def test_units_roundtrip():
    q = 5 * u.km / u.s
    assert np.isclose(q.to_value(u.m / u.s), 5000.0)
    assert np.isclose((1 * u.year).to_value(u.day), 365.25)
    assert (u.km / u.s).is_equivalent(u.m / u.s) and not (u.km / u.s).is_equivalent(u.day)
    assert np.isclose((0.1 * u.km / u.s / u.year).to_value(u.km / u.s / u.day), 0.1 / 365.25)
    assert u.as_unit("km / s") == u.km / u.s
    with pytest.raises(u.UnitsError):
        q.to_value(u.day)


# ---- RVData (thejoker/tests/test_data.py, hot-path subset) ----------------------
def test_rvdata_sort_clean_tref():
    t = np.array([5.0, 1.0, np.nan, 3.0]) + 55000
    rv = np.array([1.0, 2.0, 3.0, np.inf]) * u.km / u.s
    err = np.array([0.1, 0.2, 0.3, 0.4]) * u.km / u.s
    d = tj.RVData(t, rv, err)
    assert len(d) == 2 and np.all(np.diff(d._t_bmjd) > 0)
    assert d._t_ref_bmjd == 55001.0 and np.allclose(d.rv.value, [2.0, 1.0])
    assert np.allclose(d.ivar.value, 1 / np.array([0.2, 0.1]) ** 2)
    d2 = tj.RVData(t[:2], rv[:2], err[:2], t_ref=False)
    assert d2._t_ref_bmjd == 0.0
    with pytest.raises(ValueError):
        tj.RVData(t, rv[:3], err[:3])
    with pytest.raises(u.UnitsError):
        tj.RVData(t, np.ones(4) * u.day, err)
    d_ms = tj.RVData(t[:2], np.array([1000.0, 2000.0]) * u.m / u.s, np.array([100.0, 100.0]) * u.m / u.s)
    assert np.allclose(d_ms.rv.to_value(u.km / u.s), [2.0, 1.0])


# ---- design matrix (thejoker/tests/test_likelihood_helpers.py:8-36) --------------
def test_design_matrix():
    rnd = np.random.default_rng(42)
    sizes = (8, 4, 3)
    datas = []
    for k, n in enumerate(sizes):
        # chronological, non-overlapping surveys
        t = 55000 + 100 * k + np.sort(rnd.uniform(0, 90, n))
        datas.append(tj.RVData(t, rnd.normal(0, 10, n) * u.km / u.s, np.full(n, 0.5) * u.km / u.s))
    data, ids, M = validate_prepare_data(datas, 1, 2)
    assert np.allclose(M[:, 0], 1.0)
    idx = np.arange(len(data))
    m1 = (idx >= 8) & (idx < 12)
    assert np.allclose(M[m1, 1], 1.0) and np.allclose(M[~m1, 1], 0.0)
    m2 = idx >= 12
    assert np.allclose(M[m2, 2], 1.0) and np.allclose(M[~m2, 2], 0.0)
    # interleaved surveys: indicator columns must follow the time sort (the reference
    # leaves ids in concatenation order, data_helpers.py:117-131)
    tA, tB = 55000 + np.array([0.0, 2.0, 4.0]), 55000 + np.array([1.0, 3.0])
    dA = tj.RVData(tA, np.zeros(3) * u.km / u.s, np.ones(3) * u.km / u.s)
    dB = tj.RVData(tB, np.ones(2) * u.km / u.s, np.ones(2) * u.km / u.s)
    data, ids, M = validate_prepare_data([dA, dB], 2, 1)
    assert np.array_equal(M[:, 1], data.rv.value)  # rv==1 marks survey B
    assert np.allclose(M[:, 2], data._t_bmjd - data._t_ref_bmjd)
    with pytest.raises(ValueError):
        validate_prepare_data(dA, 1, 1)
    with pytest.raises(ValueError):
        validate_prepare_data([dA, dB], 1, 0)
    assert get_constant_term_design_matrix(dA).shape == (3, 1)
    assert get_trend_design_matrix(dA, None, 3).shape == (3, 3)


# ---- batch_tasks (thejoker/tests/test_utils.py:21-33) ---------------------------
def test_batch_tasks():
    N, start_idx = 10000, 1103
    tasks = batch_tasks(N, n_batches=16, start_idx=start_idx)
    assert tasks[0][0][0] == start_idx and tasks[-1][0][1] == N + start_idx
    tasks = batch_tasks(N, n_batches=16, start_idx=start_idx, arr=np.random.random(size=8 * N))
    assert sum(t[0].size for t in tasks) == N
    assert shard_ranges(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert shard_ranges(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    assert [b - a for a, b in shard_ranges(1 << 28, 8)] == [1 << 25] * 8
    idx, tot = merge_accepted([np.array([1, 5]), np.array([], dtype=np.int64), np.array([9])],
                              [2, 0, 1], 2)
    assert list(idx) == [1, 5] and tot == 3


# ---- prior / samples ----------------------------------------------------------------
def test_prior_default_and_sample():
    prior = default_prior(2)
    assert prior.par_names == ["P", "e", "omega", "M0", "s", "K", "v0", "v1"]
    s = prior.sample(size=1000, rng=np.random.default_rng(1), return_logprobs=True)
    assert len(s) == 1000 and np.all(np.isfinite(s["ln_prior"].value))
    P = s["P"].to_value(u.day)
    assert P.min() >= 2 and P.max() <= 1024
    assert np.all((s["e"].value > 0) & (s["e"].value < 1)) and np.all(s["s"].value == 0)
    assert s._uniform_s
    full = prior.sample(size=100, generate_linear=True, rng=np.random.default_rng(2))
    assert set(full.par_names) == set(prior.par_names)
    # K prior scale follows sigma_K0 (P/P0)^(-1/3) / sqrt(1-e^2), clipped (distributions.py:143-147)
    K = prior.pars["K"]
    assert np.isclose(K.sigma_of(365.25, 0.0), 30.0)
    assert np.isclose(K.sigma_of(1e-12, 0.0), 500.0)
    with pytest.raises(ValueError):
        tj.JokerPrior.default(sigma_K0=30 * u.km / u.s, sigma_v=100 * u.km / u.s)  # no P range
    with pytest.raises(ValueError):
        tj.JokerPrior(pars={"P": prior.pars["P"]})
    a = prior.sample(size=10, rng=np.random.default_rng(5))
    b = prior.sample(size=10, rng=np.random.default_rng(5))
    assert np.array_equal(a["P"].value, b["P"].value)


def test_samples_pack_unpack(tmp_path):
    prior = default_prior(1)
    s = prior.sample(size=50, generate_linear=True, rng=np.random.default_rng(0))
    packed, units = s.pack()
    assert packed.shape == (50, 5) and list(units) == ["P", "e", "omega", "M0", "s"]
    packed_ms, _ = s.pack(units={"s": u.m / u.s})
    assert np.allclose(packed_ms[:, 4], packed[:, 4])
    spec, data, pr = star_spec(8, 1)
    raw = np.hstack([packed, np.zeros((50, 2))])
    un = tj.JokerSamples.unpack(raw, spec["internal_units"], t_ref=1.0, poly_trend=1, n_offsets=0)
    assert list(un.keys()) == ["P", "e", "omega", "M0", "s", "K", "v0"] and un.t_ref == 1.0
    assert len(un[3:7]) == 4 and len(un[5]) == 1
    fn = str(tmp_path / "samples.npz")
    s.write(fn)
    back = tj.JokerSamples.read(fn)
    assert np.array_equal(back["P"].value, s["P"].value) and back["K"].unit == s["K"].unit
    s.write(fn, append=True)
    assert len(tj.JokerSamples.read(fn)) == 100
    k = tj.JokerSamples(poly_trend=1)
    k["K"] = np.array([-1.0, 2.0]) * u.km / u.s
    k["omega"] = np.array([0.5, 0.5]) * u.rad
    k.wrap_K()
    assert np.allclose(k["K"].value, [1, 2]) and np.allclose(k["omega"].value, [0.5 + np.pi, 0.5])
    with pytest.raises(ValueError):
        k["bogus"] = np.zeros(2)
    with pytest.raises(u.UnitsError):
        k["P"] = np.ones(2) * u.km
    assert len(s.median_period()) == 1


def test_extract_spec_layout():
    """Linear-parameter order [K, v0, dv0_*, v1, ...] (pyx:143-148, 204-252)."""
    spec, _, prior = star_spec(20, 2, n_surveys=3)
    assert spec["n_linear"] == 1 + 2 + 2
    assert list(spec["internal_units"]) == ["P", "e", "omega", "M0", "s", "K", "v0", "dv0_1",
                                            "dv0_2", "v1"]
    assert np.allclose(spec["Lambda"][1:], [100.0**2, 25.0, 25.0, 0.25])
    assert spec["K_prior_kind"] == 0 and spec["sigma_K0"] == 30.0 and spec["P0"] == 365.25
    assert spec["max_K"] == 500.0
    spec2, _, _ = star_spec(8, 1, normal_K=10.0)
    assert spec2["K_prior_kind"] == 1 and spec2["Lambda"][0] == 100.0
    data, _ = make_data(8, rng=np.random.default_rng(0))
    with pytest.raises(ValueError):  # pyx:174-179
        tj.extract_spec(data, default_prior(1), np.ones((8, 3)))
    prior_ms = tj.JokerPrior.default(P_min=2 * u.day, P_max=10 * u.day, sigma_K0=30000 * u.m / u.s,
                                     sigma_v=1e5 * u.m / u.s)
    sp = tj.extract_spec(data, prior_ms, np.ones((8, 1)))
    assert np.isclose(sp["sigma_K0"], 30.0) and np.isclose(sp["Lambda"][1], 100.0**2)


# ---- C ABI ---------------------------------------------------------------------------
def test_cabi_exports_every_declared_symbol():
    """The shared library loads and exports every function include/thejoker_b200.h
    declares (no compute calls: there is no GPU here)."""
    _lib.build()
    hdr = open(os.path.join(ROOT, "include", "thejoker_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(tjb_[A-Za-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name)
    assert lib.tjb_version() == 100
    for x in (-np.inf, -1e300, -1.5, -0.0, 0.0, 2.5, np.inf):
        assert lib.tjb_key_to_double(lib.tjb_double_to_key(x)) == x
    keys = [lib.tjb_double_to_key(x) for x in (-np.inf, -2.0, -1.0, 0.0, 1.0, np.inf, np.nan)]
    assert keys == sorted(keys)  # NaN sorts above +inf, like numpy.max propagates it


def test_header_is_plain_c_and_links(tmp_path):
    """include/thejoker_b200.h is what a C / cgo / JNI binding would include: it must compile
    as strict C99 (no C++-isms, no torch types) and a C program must link against the
    library and get a clean error back without a GPU."""
    import shutil
    import subprocess

    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    _lib.build()
    src = tmp_path / "cabi.c"
    src.write_text(
        "#include <stdio.h>\n#include <string.h>\n#include \"thejoker_b200.h\"\n"
        "int main(void) {\n"
        "  TjbHandle *h = NULL;\n"
        "  TjbSpec spec; TjbPcg64 pcg; TjbPriorGen gen; TjbMultiStarJob job;\n"
        "  (void)spec; (void)pcg; (void)gen; (void)job;\n"
        "  int rc = tjb_create(NULL, 0, &h);\n"
        "  printf(\"%d %d %s\\n\", tjb_version(), rc, tjb_last_error());\n"
        "  return (rc < 0 && strlen(tjb_last_error()) > 0) ? 0 : 1;\n}\n")
    exe = tmp_path / "cabi"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror",
                    "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L", libdir, "-l:" + os.path.basename(_lib.LIB_PATH),
                    "-Wl,-rpath," + libdir], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.split()[0] == "100"


def test_no_gpu_fails_loudly():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    spec, data, prior = star_spec(8, 1)
    with pytest.raises(_lib.TjbError):
        tj.TheJoker(prior)._make_joker_helper(data)


def test_multistar_job_validation():
    """tjb_multistar_rejection checks its job before touching CUDA (no GPU here: a valid
    job must fail with the no-device error, not crash)."""
    import ctypes

    import torch

    lib = _lib.load()
    spec, data, prior = star_spec(8, 1)
    cs = (_lib.TjbSpec * 2)()
    cs[0] = tj.CJokerHelper._c_spec(spec)
    cs[1] = tj.CJokerHelper._c_spec(spec)
    pcg = (_lib.TjbPcg64 * 2)()
    counts = np.zeros((2, 3), dtype=np.int64)
    idx = np.zeros((2, 4), dtype=np.int64)
    dummy = ctypes.c_void_p(8)  # never dereferenced before the device check
    vp = lambda a: ctypes.c_void_p(a.ctypes.data)

    def job(**kw):
        base = dict(n_stars=2, specs=cs, pcg=pcg, d_P=dummy, d_e=dummy, d_omega=dummy,
                    d_M0=dummy, d_s=None, s_const=0.0, n_prior=16, max_keep=4, near_tol=1e-12,
                    n_per=0, clamp_K=0, n_slots=2, h_idx=vp(idx), h_counts=vp(counts))
        base.update(kw)
        return _lib.TjbMultiStarJob(**base)

    def run(j):
        return lib.tjb_multistar_rejection(0, ctypes.byref(j)), lib.tjb_last_error().decode()

    assert lib.tjb_multistar_rejection(0, None) == -1
    assert run(job(n_stars=0))[0] == 0  # nothing to do
    for bad in (dict(n_prior=0), dict(n_slots=0), dict(n_slots=65), dict(max_keep=-1),
                dict(d_P=None), dict(h_counts=None), dict(specs=None), dict(n_per=1)):
        rc, msg = run(job(**bad))
        assert rc == -1 and msg, bad
    cs[1].n_linear = 3
    rc, msg = run(job())
    assert rc == -1 and "n_linear" in msg
    cs[1].n_linear = cs[0].n_linear
    if not torch.cuda.is_available():
        rc, msg = run(job())
        assert rc == -2 and "no CUDA device" in msg


def test_multistar_native_pipeline_host_side(monkeypatch):
    """The host side of the native multi-star engine -- chunking, per-star specs, child
    generators, pre-drawn normals, unpacking -- with the library call replaced by a stand-in
    that echoes what it was given (no GPU here; tests/test_gpu_parity.py runs the real one)."""
    from helpers import fake_multistar_stars

    n_prior, n_stars, keep, n_per, L = 64, 45, 5, 2, 3
    prior, stars = fake_multistar_stars(n_stars)
    from helpers import fake_multistar_lib, fake_multistar_joker

    FakeLib, calls = fake_multistar_lib(n_prior, keep, n_per, L)
    monkeypatch.setattr(_lib, "load", lambda: FakeLib)
    ms = fake_multistar_joker(prior, n_prior, L, rng=np.random.default_rng(9))
    out = ms.rejection_sample(stars, max_posterior_samples=keep, n_linear_samples=n_per,
                              return_logprobs=True)
    assert sum(calls) == n_stars and calls[0] == 8 and len(calls) > 2  # growing chunks
    seqs = np.random.default_rng(9).bit_generator._seed_seq.spawn(n_stars)
    for i, smp in enumerate(out):
        n_t = 10 + i % 7
        k = n_t % (keep + 1)
        assert len(smp) == k * n_per
        child = np.random.Generator(np.random.PCG64(seqs[i]))
        assert ms.last_stats[i] == dict(n_accepted=k + 10, n_near_threshold=1,
                                        ll_max=float((child.bit_generator.state["state"]["state"]
                                                      & ((1 << 64) - 1)) % 1000))
        child.bit_generator.advance(n_prior)
        z = child.standard_normal((k, n_per, L)).reshape(-1, L)
        assert np.array_equal(smp["P"].value, np.full(k * n_per, float(n_t)))
        assert np.array_equal(smp["K"].value, z[:, 0]) and np.array_equal(smp["dv0_1"].value, z[:, 2])
        assert np.array_equal(smp["ln_likelihood"], np.repeat(np.arange(k), n_per))
    # settings the native loop does not cover fall back to the Python engine
    assert tj.MultiStarJoker(prior, None, engine="python").engine == "python"
    with pytest.raises(ValueError):
        tj.MultiStarJoker(prior, None, engine="cpu")


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "thejoker_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), fn


# ---- device math on the host ---------------------------------------------------------
LEGACY = "TJB_TRIM=0 TJB_PHASE_FIXED=0 TJB_HALLEY=0 TJB_TRIG_TABLE_LOG2=10 TJB_EPOCHS_PER_ITER=2"  # round-1 loop


@pytest.mark.parametrize("variant", ["", "TJB_TRIG_TABLE=0", "TJB_TRIM=1 TJB_TRIG_TABLE_LOG2=11", LEGACY])
def test_sincos_units(variant):
    """The FP64 sin/cos of the epoch loop (table node + short polynomial, or the minimax
    polynomial back-end) against 40-digit mpmath, over many revolutions."""
    import mpmath as mp

    lib = host_emulation(variant)
    rng = np.random.default_rng(0)
    s, c = ctypes.c_double(), ctypes.c_double()
    worst = 0.0
    mp.mp.dps = 40
    revs = np.concatenate([rng.uniform(-3, 3, 3000), rng.uniform(-500, 500, 500),
                           np.arange(-8, 9) / 8.0, (np.arange(0, 4096) + 0.5) / 4096])
    for rev in revs:
        lib.emu_sincos_rev(float(rev), ctypes.byref(s), ctypes.byref(c))
        ang = 2 * mp.pi * mp.mpf(float(rev))
        worst = max(worst, abs(float(mp.sin(ang) - mp.mpf(s.value))),
                    abs(float(mp.cos(ang) - mp.mpf(c.value))))
    assert worst < 3.5e-16


def test_kepler_column_matches_oracle():
    from oracle.oracle import OracleHelper

    lib = host_emulation()
    spec, _, _ = star_spec(64, 1)
    orc = OracleHelper.from_spec(spec)
    chunk = prior_chunk(3000)
    chunk[:200, 1] = np.random.default_rng(1).uniform(0.8, 0.999, 200)  # high-e tail
    dt = np.ascontiguousarray(spec["t"] - spec["t0"])
    dp = ctypes.POINTER(ctypes.c_double)
    z = np.zeros(64)
    st = (ctypes.c_int * 3)()
    worst, not_conv = 0.0, 0
    for row in chunk:
        lib.emu_design_column(row[0], row[1], row[2], row[3], dt.ctypes.data_as(dp), 64,
                              z.ctypes.data_as(dp), st)
        zo = orc.design_column(row)
        # the phase 2 pi dt / P - M0 is rounded differently (both ~1e-13 rad at |M| ~ 500)
        scale = 1.0 / (1.0 - row[1]) ** 2
        worst = max(worst, np.max(np.abs(z - zo)) / scale)
        not_conv += st[2]
    assert not_conv == 0
    assert worst < 2e-12


@pytest.mark.parametrize("N,pt,sl,kw", [
    (16, 1, None, {}), (64, 1, None, {}), (64, 2, (-2.0, 1.0), {}), (3, 1, None, {"normal_K": 10.0}),
    (20, 1, None, {"n_surveys": 2}), (12, 3, None, {}), (24, 2, None, {"n_surveys": 3}),
])
def test_host_emulated_ll_matches_oracle(N, pt, sl, kw):
    """The kernel's algebra (Gram sums + LDL^T + determinant lemma), compiled for the
    host, against the O(N^3) restatement of the reference and the quad truth."""
    from oracle.oracle import OracleHelper

    spec, _, _ = star_spec(N, pt, **kw)
    chunk = prior_chunk(1500, s_lognormal=sl)
    orc = OracleHelper.from_spec(spec)
    ref = orc.batch_marginal_ln_likelihood(chunk, n_threads=0)
    truth, _ = orc.truth_ll(chunk)
    for force_jit in ([True] if sl is not None else [False, True]):
        got = emu_marginal_ll(spec, chunk, force_jit=force_jit)
        r_ref, r_truth, ref_truth = rel_err(got, ref), rel_err(got, truth), rel_err(ref, truth)
        # north-star gate: 1e-10 relative vs the reference algorithm -- or at least as
        # close to the exact value as the reference algorithm itself is
        ok = (r_ref <= 1e-10) | (r_truth <= ref_truth)
        assert ok.all(), (r_ref.max(), r_truth.max(), ref_truth.max())
        assert np.max(r_truth) < 1e-10


VARIANTS = [LEGACY, "TJB_NEED_LOG2=16", "TJB_NEED_LOG2=17", "TJB_VOTE_D2=1", "TJB_VOTE_D2=1 TJB_EPOCHS_PER_ITER=4",
            "TJB_TRIM=1", "TJB_PHASE_FIXED=1", "TJB_TRIM=1 TJB_PHASE_FIXED=1",
            "TJB_TRIM=1 TJB_PHASE_FIXED=1 TJB_TRIG_TABLE=0",
            "TJB_TRIM=1 TJB_PHASE_FIXED=1 TJB_HALLEY=1",
            "TJB_TRIM=1 TJB_PHASE_FIXED=1 TJB_HALLEY=1 TJB_TRIG_TABLE_LOG2=11",
            "TJB_TRIM=1 TJB_PHASE_FIXED=1 TJB_HALLEY=1 TJB_TRIG_TABLE_LOG2=11 TJB_EPOCHS_PER_ITER=3",
            "TJB_TRIM=1 TJB_PHASE_FIXED=1 TJB_EPOCHS_PER_ITER=4",
            # round 2c defaults: two-level table, z reciprocal from the step's, 4 epochs / iteration
            "TJB_TRIG2=0 TJB_EPOCHS_PER_ITER=3", "TJB_TRIG2=0", "TJB_XZ=1",
            "TJB_XZ=1 TJB_WIDE_THREADS=0", "TJB_EPOCHS_PER_ITER=2", "TJB_EPOCHS_PER_ITER=3"]


@pytest.mark.parametrize("variant", VARIANTS)
def test_host_emulated_tuning_variants(variant):
    """Compile-time variants of the epoch loop that are built for timing on the GPU
    (tools/build_variants.sh) must keep the shipped configuration's numerics: same Kepler
    column to a few ulp (also on the high-eccentricity rare path) and the same ll gate."""
    from oracle.oracle import OracleHelper

    base, var = host_emulation(), host_emulation(variant)
    spec, _, _ = star_spec(64, 1)
    chunk = prior_chunk(3000)
    chunk[:300, 1] = np.random.default_rng(1).uniform(0.8, 0.999, 300)  # exercises extra passes
    dt = np.ascontiguousarray(spec["t"] - spec["t0"])
    dp = ctypes.POINTER(ctypes.c_double)
    za, zb = np.zeros(64), np.zeros(64)
    sa, sb = (ctypes.c_int * 3)(), (ctypes.c_int * 3)()
    worst, worst_orc, extra = 0.0, 0.0, 0
    orc = OracleHelper.from_spec(spec)
    for row in chunk:
        base.emu_design_column(*row[:4], dt.ctypes.data_as(dp), 64, za.ctypes.data_as(dp), sa)
        var.emu_design_column(*row[:4], dt.ctypes.data_as(dp), 64, zb.ctypes.data_as(dp), sb)
        assert sb[2] == 0
        extra += sb[1]
        worst = max(worst, np.max(np.abs(za - zb)) * (1.0 - row[1]) ** 2)
        worst_orc = max(worst_orc, np.max(np.abs(zb - orc.design_column(row))) * (1.0 - row[1]) ** 2)
    assert extra > 0          # the rare path ran
    # both round the phase x4 + d4 (|x4| ~ 1e5 angle units: ulp ~ 9e-14 rad) differently
    assert worst < 2e-13
    assert worst_orc < 2e-12  # the gate of test_kepler_column_matches_oracle
    # even / odd epoch counts (the remainder loop after the groups of kEpochsPerIter), L up to 5
    for N, pt, sl in ((64, 1, None), (20, 2, (-2.0, 1.0)), (33, 1, None), (3, 1, None),
                      (17, 4, (-2.0, 1.0))):
        spec, _, _ = star_spec(N, pt)
        chunk = prior_chunk(1500, s_lognormal=sl)
        orc = OracleHelper.from_spec(spec)
        truth, _ = orc.truth_ll(chunk)
        got = emu_marginal_ll(spec, chunk, force_jit=sl is not None, variant=variant)
        ref = emu_marginal_ll(spec, chunk, force_jit=sl is not None)
        # (differently rounded phases: ~1e-13 rad, times the conditioning of ll)
        assert np.max(rel_err(got, ref)) < 2e-11, (N, pt)
        assert np.max(rel_err(got, truth)) < 1e-10, (N, pt)


@pytest.mark.parametrize("variant", ["", "TJB_TRIM=1 TJB_PHASE_FIXED=1",
                                     "TJB_TRIM=1 TJB_PHASE_FIXED=1 TJB_HALLEY=1",
            "TJB_TRIM=1 TJB_PHASE_FIXED=1 TJB_HALLEY=1 TJB_TRIG_TABLE_LOG2=11",
            "TJB_TRIM=1 TJB_PHASE_FIXED=1 TJB_HALLEY=1 TJB_TRIG_TABLE_LOG2=11 TJB_EPOCHS_PER_ITER=3",
            "TJB_TRIM=1 TJB_PHASE_FIXED=1 TJB_EPOCHS_PER_ITER=4",
                                     "TJB_TRIG2=0 TJB_EPOCHS_PER_ITER=3", "TJB_XZ=1"])
def test_kepler_solver_extreme_cases(variant):
    """e -> 1 at M -> 0, phases beyond the FP32 stage's range (P = 0.05 d over 10 000 d):
    the safeguarded extra passes (kepler.cuh::solve_extra_passes) always converge to a
    finite value -- one NaN ll would poison the max of the whole prior cache and the
    rejection step would accept nothing -- and the value is right (50-digit mpmath)."""
    import mpmath as mp

    lib = host_emulation(variant)
    dp = ctypes.POINTER(ctypes.c_double)
    rng = np.random.default_rng(3)
    z, st = np.zeros(64), (ctypes.c_int * 3)()
    n = 3000
    for span, P_min in ((155.0, 2.0), (10000.0, 0.05)):
        dt = np.sort(rng.uniform(0, span, 64))
        P = np.exp(rng.uniform(np.log(P_min), np.log(1000), n))
        e = 1 - 10 ** rng.uniform(-7, -1, n)
        om, M0 = rng.uniform(-np.pi, np.pi, (2, n))
        for i in range(n):
            lib.emu_design_column(P[i], e[i], om[i], M0[i], dt.ctypes.data_as(dp), 64,
                                  z.ctypes.data_as(dp), st)
            assert st[2] == 0 and np.isfinite(z).all(), (P[i], e[i], om[i], M0[i])
    # accuracy at the hard corner, scaled by the conditioning d z / d M ~ 1 / (1 - e cosE)^2
    mp.mp.dps = 50
    worst, one, zz = 0.0, np.zeros(1), np.zeros(1)
    for trial in range(120):
        e = 1 - 10 ** rng.uniform(-7, -2)
        P = float(np.exp(rng.uniform(np.log(2), np.log(1000))))
        om, M0 = rng.uniform(-np.pi, np.pi, 2)
        Mt = 10 ** rng.uniform(-9, -1) * rng.choice([-1, 1])   # a tiny mean anomaly
        one[0] = (Mt + M0) * P / (2 * np.pi) + P * rng.integers(0, 3)
        lib.emu_design_column(P, e, om, M0, one.ctypes.data_as(dp), 1, zz.ctypes.data_as(dp), st)
        assert st[2] == 0
        M = 2 * mp.pi * mp.mpf(float(one[0])) / mp.mpf(P) - mp.mpf(float(M0))
        M -= 2 * mp.pi * mp.nint(M / (2 * mp.pi))
        lo, hi = M - mp.mpf(e), M + mp.mpf(e)
        for _ in range(180):
            mid = (lo + hi) / 2
            if mid - mp.mpf(e) * mp.sin(mid) - M > 0:
                hi = mid
            else:
                lo = mid
        E = (lo + hi) / 2
        a, b = mp.cos(mp.mpf(float(om))), -mp.sqrt(1 - mp.mpf(e) ** 2) * mp.sin(mp.mpf(float(om)))
        f1 = 1 - mp.mpf(e) * mp.cos(E)
        zt = (a * (mp.cos(E) - mp.mpf(e)) + b * mp.sin(E)) / f1 + mp.mpf(e) * a
        worst = max(worst, float(abs(mp.mpf(float(zz[0])) - zt) * f1 ** 2))
    assert worst < 1e-14


def test_phase_reduction_variants():
    """How often the one-pass FP64 step is not enough, against the time baseline of the data
    (default prior, P >= 2 d): the round-1 FP32 stage rounds the unreduced phase to float, so
    its starter degrades with the number of revolutions; the fixed-point reduction
    (TJB_PHASE_FIXED, the default now) does not.  Extra passes are the same code on the GPU,
    where one lane that needs them sends its whole warp through the rare path.  The shipped
    loop (Halley step, threshold 2^-17) takes the rare path more often than the third-order
    step did (2^-13), still independent of the baseline."""
    libs = {"": host_emulation(LEGACY),
            "fixed": host_emulation("TJB_TRIM=1 TJB_PHASE_FIXED=1 TJB_HALLEY=0"),
            "shipped": host_emulation()}
    chunk = prior_chunk(2048)
    dp = ctypes.POINTER(ctypes.c_double)
    z, st = np.zeros(64), (ctypes.c_int * 3)()
    warp_rate = {}
    for span in (155.0, 4000.0):
        dt = np.sort(np.random.default_rng(0).uniform(0, span, 64))
        for name, lib in libs.items():
            extra = []
            for row in chunk:
                lib.emu_design_column(*row[:4], dt.ctypes.data_as(dp), 64, z.ctypes.data_as(dp), st)
                assert st[2] == 0
                extra.append(st[1])
            # share of (warp, epoch) pairs in which some lane needs another pass, if the
            # lanes' epochs were independent: an upper bound of the rare-path rate
            warp_rate[name, span] = np.reshape(extra, (-1, 32)).sum(axis=1).mean() / 64
    assert warp_rate["", 155.0] < 0.01 and warp_rate["fixed", 155.0] < 0.01
    assert warp_rate["", 4000.0] > 3 * warp_rate["fixed", 4000.0]
    assert warp_rate["fixed", 4000.0] < 1.5 * warp_rate["fixed", 155.0] + 1e-3
    assert warp_rate["shipped", 155.0] < 0.05
    assert warp_rate["shipped", 4000.0] < 1.5 * warp_rate["shipped", 155.0] + 1e-3


def test_host_emulated_uniform_jitter_paths_agree():
    spec, _, _ = star_spec(32, 2)
    chunk = prior_chunk(500, s_const=0.37)
    a = emu_marginal_ll(spec, chunk, force_jit=False)
    b = emu_marginal_ll(spec, chunk, force_jit=True)
    assert np.max(rel_err(a, b)) < 1e-12
    spec0 = dict(spec, jitter_mode=0)
    c = emu_marginal_ll(spec0, chunk, force_jit=True)
    chunk0 = chunk.copy()
    chunk0[:, 4] = 0
    assert np.max(rel_err(c, emu_marginal_ll(spec, chunk0))) < 1e-13


def test_pcg64_leapfrog_matches_numpy():
    lib = host_emulation()
    rng = np.random.default_rng(20261017)
    st = rng.bit_generator.state["state"]
    m = (1 << 64) - 1
    args = (st["state"] >> 64, st["state"] & m, st["inc"] >> 64, st["inc"] & m)
    want = rng.random(5000)
    for i in (0, 1, 2, 255, 256, 4095, 4999):
        assert lib.emu_pcg64_double(*args, i) == want[i]
    r2 = np.random.default_rng(20261017)
    r2.bit_generator.advance(5000)
    assert np.array_equal(rng.standard_normal(4), r2.standard_normal(4))


def test_ll_key_order():
    lib = host_emulation()
    xs = np.array([-np.inf, -1e10, -3.0, -1e-300, 0.0, 1e-300, 7.0, np.inf])
    ks = [lib.emu_ll_to_key(x) for x in xs]
    assert ks == sorted(ks) and all(lib.emu_key_to_ll(k) == x for k, x in zip(ks, xs))
    assert lib.emu_ll_to_key(np.nan) > lib.emu_ll_to_key(np.inf)


def test_samples_analysis():
    """thejoker/tests/test_samples_analysis.py in miniature."""
    from thejoker_b200 import samples_analysis as sa

    data, _ = make_data(30, rng=np.random.default_rng(0))
    s = tj.JokerSamples()
    s["P"] = np.array([51.8, 51.81, 51.79]) * u.day
    s["ln_prior"] = np.array([0.0, 0.0, 0.0])
    s["ln_likelihood"] = np.array([-3.0, -1.0, -2.0])
    best, idx = sa.MAP_sample(s, return_index=True)
    assert idx == 1 and best["P"].value[0] == 51.81
    assert sa.is_P_unimodal(s, data)
    wide = tj.JokerSamples()
    wide["P"] = np.array([20.0, 51.8, 300.0]) * u.day
    assert not sa.is_P_unimodal(wide, data)
    two = tj.JokerSamples()
    two["P"] = np.concatenate([np.full(5, 51.8) + 1e-3 * np.arange(5),
                               np.full(4, 103.6) + 1e-3 * np.arange(4)]) * u.day
    ok, reps, counts = sa.is_P_Kmodal(two, data, n_clusters=2)
    assert ok and sorted(counts) == [4, 5] and np.allclose(sorted(reps.value), [51.8, 103.6], atol=0.1)
    one = s[0]
    assert 0 < sa.max_phase_gap(one, data) < 1 and 0 < sa.phase_coverage(one, data) <= 1
    assert np.isclose(sa.periods_spanned(one, data), np.ptp(data._t_bmjd) / 51.8)
    assert sa.phase_coverage_per_period(one, data) >= 1
    with pytest.raises(ValueError):
        sa.MAP_sample(wide)


def test_prior_cache_roundtrip(tmp_path):
    """Native SoA prior cache (section 8 f1): units converted once, constant jitter stored
    as a scalar, memory-mapped shards."""
    from thejoker_b200.cache import PriorCache, read_reference_hdf5, write_prior_cache

    prior = default_prior(1)
    s = prior.sample(size=1000, rng=np.random.default_rng(0), return_logprobs=True)
    path = write_prior_cache(s, str(tmp_path / "cache"))
    c = PriorCache(path)
    assert len(c) == 1000 and c.s_const == 0.0 and c.has_ln_prior
    cols = c.columns(lo=100, hi=200)
    assert np.array_equal(cols[0], s["P"].to_value(u.day)[100:200]) and cols[4] == 0.0
    back = c.to_samples()
    assert np.array_equal(back["e"].value, s["e"].value) and back._uniform_s
    from thejoker_b200.prior import LogNormal
    prior_s = default_prior(1, s=LogNormal("s", 0.0, 1.0, u.m / u.s))
    s2 = prior_s.sample(size=50, rng=np.random.default_rng(0))
    c2 = PriorCache(write_prior_cache(s2, str(tmp_path / "cache2"), rv_unit=u.km / u.s))
    assert c2.s_const is None
    assert np.allclose(c2.columns(rv_unit=u.m / u.s)[4], s2["s"].to_value(u.m / u.s))
    with pytest.raises(OSError):
        write_prior_cache(s, path)
    # no h5py needed: the dependency-free reader (hdf5_min) takes over, and a missing file is
    # an ordinary OSError
    with pytest.raises(OSError):
        read_reference_hdf5(str(tmp_path / "nope.hdf5"))


@pytest.mark.parametrize("kind", ["cache_dir", "npz"])
def test_read_batch_family(tmp_path, kind):
    """thejoker/tests/test_utils.py::test_read_batch_slice / _idx / _random_batch on the
    native cache directory and on a JokerSamples.write file."""
    from thejoker_b200.cache import (read_batch, read_batch_idx, read_batch_slice,
                                     read_random_batch, write_prior_cache)

    prior = default_prior(1)
    ps = prior.sample(size=100, rng=np.random.default_rng(3), return_logprobs=True)
    ps["s"] = np.linspace(0.0, 1.0, 100) * u.km / u.s
    ps._uniform_s = False
    if kind == "cache_dir":
        fn = write_prior_cache(ps, str(tmp_path / "cache"))
    else:
        fn = str(tmp_path / "samples.npz")
        ps.write(fn)
    P, om, sv = ps["P"].to_value(u.day), ps["omega"].to_value(u.rad), ps["s"].to_value(u.km / u.s)
    for read_func in (read_batch_slice, read_batch):
        batch = read_func(fn, ["P", "omega"], slice(10, 20))
        assert batch.shape == (10, 2)
        assert np.allclose(batch[:, 0], P[10:20]) and np.allclose(batch[:, 1], om[10:20])
        batch = read_func(fn, ["s", "P"], slice(0, 100), units={"s": u.m / u.s})
        assert batch.shape == (100, 2)
        assert np.allclose(batch[:, 0], sv * 1e3) and np.allclose(batch[:, 1], P)
        assert read_func(fn, ["P", "e"], slice(0, 100, 2)).shape == (50, 2)
    assert np.array_equal(read_batch(fn, ["P"], (5, 9)), read_batch_slice(fn, ["P"], slice(5, 9)))
    idx = np.arange(10, 20)
    for read_func in (read_batch_idx, read_batch):
        batch = read_func(fn, ["P", "omega"], idx, units=None)
        assert batch.shape == (10, 2) and np.allclose(batch[:, 0], P[idx])
        batch = read_func(fn, ["s", "P"], idx, units={"s": u.m / u.s})
        assert np.allclose(batch[:, 0], sv[idx] * 1e3) and np.allclose(batch[:, 1], P[idx])
    for read_func in (read_random_batch, read_batch):
        assert read_func(fn, ["P", "e"], 10, units=None).shape == (10, 2)
        full = read_func(fn, ["P", "e"], 100, units=None, rng=np.random.default_rng(1))
        assert full.shape == (100, 2) and np.allclose(np.sort(full[:, 0]), np.sort(P))
    with pytest.raises(ValueError):
        read_batch(fn, ["P"], "nope")
    assert np.allclose(read_batch(fn, ["ln_prior"], slice(0, 5))[:, 0], ps["ln_prior"].value[:5])


def test_table_header_parser_on_astropy_headers():
    """parse_table_column_meta on headers written by astropy's own table serialiser (the
    reference's docs/examples/*.ecsv, same YAML as ``samples.__table_column_meta__``):
    units come from the datatype entries, the ``!astropy...`` nodes of the serialised
    Quantity columns must not leak into them, and a serialised Time becomes an MJD."""
    import json

    from thejoker_b200.cache import _meta_kwargs, parse_table_column_meta

    g = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_table_headers.json")))
    units, cols, meta = parse_table_column_meta(g["ecsv:data.ecsv"])
    assert cols == ["bjd", "rv", "rv_err"]
    assert units == {"bjd": "", "rv": "km / s", "rv_err": "km / s"}
    assert abs(_meta_kwargs(meta)["t_ref"] - 58511.29511355243) < 1e-9
    assert "__serialized_columns__" not in meta
    units, _, meta = parse_table_column_meta(g["ecsv:data-triple.ecsv"])
    assert units["rv"] == "km / s" and _meta_kwargs(meta) == {}
    for key in ("prior_samples", "prior_samples_units_in_meta_only"):
        lines = [ln.encode() for ln in g[key]]  # h5py hands back bytes
        units, cols, meta = parse_table_column_meta(lines)
        assert cols == ["P", "e", "omega", "M0", "s", "ln_prior"]
        assert units == {"P": "d", "e": "", "omega": "rad", "M0": "rad", "s": "m / s",
                         "ln_prior": ""}
        assert _meta_kwargs(meta) == {"poly_trend": 1, "n_offsets": 0}
    with pytest.raises(ValueError):
        parse_table_column_meta(["meta: {a: 1}"])


def _fake_h5py(monkeypatch, files):
    """A module with h5py's read interface over in-memory arrays: File(name, 'r') is a
    context manager mapping dataset names to objects with len / dtype / slicing / [()]."""
    import types

    class Dataset:
        def __init__(self, arr):
            self._a = arr
            self.n_reads, self.max_rows = 0, 0

        dtype = property(lambda self: self._a.dtype)

        def __len__(self):
            return len(self._a)

        def __getitem__(self, key):
            out = self._a[key]
            self.n_reads += 1
            self.max_rows = max(self.max_rows, np.size(out))
            return out

    class File(dict):
        def __init__(self, name, mode="r"):
            assert mode == "r"
            super().__init__(files[name])

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

    store = {name: {k: Dataset(v) for k, v in dsets.items()} for name, dsets in files.items()}
    files.clear()
    files.update(store)
    mod = types.ModuleType("h5py")
    mod.File = File
    monkeypatch.setitem(sys.modules, "h5py", mod)
    return store


def test_reference_hdf5_cache_reader_and_converter(tmp_path, monkeypatch):
    """The reference's prior-cache layout (compound-row dataset ``samples`` + YAML header
    dataset, samples.py:535-563) through read_reference_hdf5 and the blockwise
    convert_reference_hdf5, with h5py replaced by an in-memory stand-in (h5py is not in this
    image) and the header in astropy's format."""
    import json

    from thejoker_b200.cache import (PriorCache, convert_reference_hdf5, read_reference_hdf5,
                                     write_prior_cache)

    g = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_table_headers.json")))
    n = 1000
    rng = np.random.default_rng(4)
    rows = np.zeros(n, dtype=[(c, "<f8") for c in ("P", "e", "omega", "M0", "s", "ln_prior")])
    rows["P"], rows["e"] = rng.uniform(2, 100, n), rng.uniform(0, 0.9, n)
    rows["omega"], rows["M0"] = rng.uniform(0, 6, n), rng.uniform(0, 6, n)
    rows["s"], rows["ln_prior"] = rng.uniform(0, 300, n), rng.normal(size=n)  # s in m/s
    hdr = np.array([ln.encode() for ln in g["prior_samples"]])
    const = rows.copy()
    const["s"] = 125.0
    files = {"prior.hdf5": {"samples": rows, "samples.__table_column_meta__": hdr},
             "const.hdf5": {"samples": const, "samples.__table_column_meta__": hdr}}
    store = _fake_h5py(monkeypatch, files)

    smp = read_reference_hdf5("prior.hdf5")
    assert len(smp) == n and smp.poly_trend == 1 and smp.n_offsets == 0
    assert np.array_equal(smp["P"].to_value(u.day), rows["P"])
    assert np.allclose(smp["s"].to_value(u.km / u.s), rows["s"] / 1e3, rtol=1e-15)
    assert np.array_equal(smp["ln_prior"].value, rows["ln_prior"])
    part = read_reference_hdf5("prior.hdf5", lo=100, hi=164)
    assert np.array_equal(part["e"].value, rows["e"][100:164])

    store["prior.hdf5"]["samples"].max_rows = 0
    out = convert_reference_hdf5("prior.hdf5", str(tmp_path / "c1"), rows_per_block=128)
    assert store["prior.hdf5"]["samples"].max_rows <= 128  # never the whole table at once
    cache = PriorCache(out)
    cols = cache.columns()
    for i, c in enumerate(("P", "e", "omega", "M0")):
        assert np.array_equal(cols[i], rows[c])
    assert np.allclose(cols[4], rows["s"] / 1e3, rtol=1e-15) and cache.s_const is None
    assert np.array_equal(cache.ln_prior(), rows["ln_prior"])
    # the converted cache equals what write_prior_cache makes of the same samples
    write_prior_cache(smp, str(tmp_path / "c2"))
    for c in ("P", "e", "omega", "M0", "s", "ln_prior"):
        assert np.allclose(np.load(tmp_path / "c1" / f"{c}.npy"), np.load(tmp_path / "c2" / f"{c}.npy"),
                           rtol=1e-15, atol=0)
    # a constant jitter column collapses to a scalar; rv_unit converts it
    c3 = PriorCache(convert_reference_hdf5("const.hdf5", str(tmp_path / "c3"), rv_unit=u.m / u.s,
                                           rows_per_block=300))
    assert c3.s_const == 125.0 and not os.path.exists(tmp_path / "c3" / "s.npy")
    assert c3.columns(rv_unit=u.km / u.s)[4] == pytest.approx(0.125)
    with pytest.raises(OSError):
        convert_reference_hdf5("prior.hdf5", str(tmp_path / "c1"))
    # the reference's read_batch family over the converted cache (utils.py:106-245)
    from thejoker_b200.cache import read_batch

    b = read_batch(out, ["P", "e", "s"], (10, 20), units={"s": u.m / u.s})
    assert np.array_equal(b[:, 0], rows["P"][10:20]) and np.allclose(b[:, 2], rows["s"][10:20])


def test_hdf5_min_reads_a_real_libhdf5_file():
    """hdf5_min on bytes written by libhdf5 itself (not by anything in this repository):
    scipy ships MATLAB's ``testhdf5_7.4_GLNX86.mat`` -- an HDF5 file behind a 512-byte user
    block, version-0 superblock, old-style root group (B-tree + local heap + symbol-table
    node), one contiguous float64 dataset holding 0 : pi/4 : 2 pi."""
    import scipy.io

    from thejoker_b200 import hdf5_min
    from thejoker_b200.samples import file_format

    path = os.path.join(os.path.dirname(scipy.io.__file__), "matlab", "tests", "data",
                        "testhdf5_7.4_GLNX86.mat")
    if not os.path.exists(path):
        pytest.skip("scipy's test data is not installed")
    assert file_format(path) == "hdf5"
    with hdf5_min.File(path) as f:
        assert f.keys() == ["testdouble"] and "testdouble" in f
        d = f["testdouble"]
        assert d.shape == (9, 1) and d.dtype == np.dtype("<f8")
        assert np.allclose(d[()].ravel(), np.arange(9) * np.pi / 4, rtol=0, atol=1e-15)
        with pytest.raises(KeyError):
            f["nope"]


@pytest.mark.parametrize("two_level,user_block", [(False, 0), (True, 0), (True, 1024)])
def test_reference_hdf5_layout_through_hdf5_min(tmp_path, two_level, user_block):
    """The reference's prior-cache file layout -- compound rows in a *chunked*, resizable
    dataset (``maxshape=(None,)``, samples.py:535-545) indexed by a version-1 B-tree, YAML
    header in a fixed-string dataset -- written byte by byte from the HDF5 format
    specification (tests/hdf5_writer.py; no HDF5 library is in the image) and read back with
    the dependency-free reader through the public entry points: read_reference_hdf5,
    JokerSamples.read (format detected from the bytes), convert_reference_hdf5, read_batch."""
    import json

    from hdf5_writer import write_reference_style_hdf5

    from thejoker_b200.cache import (PriorCache, convert_reference_hdf5, read_batch,
                                     read_reference_hdf5)

    g = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_table_headers.json")))
    n = 1000  # not a multiple of the chunk size: the last chunk is partly unused
    rng = np.random.default_rng(4)
    rows = np.zeros(n, dtype=[(c, "<f8") for c in ("P", "e", "omega", "M0", "s", "ln_prior")])
    rows["P"], rows["e"] = rng.uniform(2, 100, n), rng.uniform(0, 0.9, n)
    rows["omega"], rows["M0"] = rng.uniform(0, 6, n), rng.uniform(0, 6, n)
    rows["s"], rows["ln_prior"] = rng.uniform(0, 300, n), rng.normal(size=n)  # s in m/s
    path = write_reference_style_hdf5(str(tmp_path / "prior_samples.hdf5"), rows,
                                      [ln.encode() for ln in g["prior_samples"]], chunk_rows=64,
                                      two_level=two_level, user_block=user_block)
    assert "h5py" not in sys.modules
    smp = read_reference_hdf5(path)
    assert len(smp) == n and smp.poly_trend == 1 and smp.n_offsets == 0
    for c in ("P", "e", "omega", "M0"):
        assert np.array_equal(smp[c].value, rows[c])
    assert np.allclose(smp["s"].to_value(u.km / u.s), rows["s"] / 1e3, rtol=1e-15)
    assert np.array_equal(smp["ln_prior"].value, rows["ln_prior"])
    # row ranges that start / end inside chunks and cross B-tree nodes
    for lo, hi in ((0, 1), (63, 65), (100, 164), (500, 1000), (999, 1000), (10, 10)):
        part = read_reference_hdf5(path, lo=lo, hi=hi)
        assert np.array_equal(part["e"].value, rows["e"][lo:hi])
    # the name does not matter: JokerSamples.read tells the formats apart by their bytes
    other = str(tmp_path / "cache.dat")
    os.replace(path, other)
    again = tj.JokerSamples.read(other)
    assert np.array_equal(again["M0"].value, rows["M0"])
    out = convert_reference_hdf5(other, str(tmp_path / "native"), rows_per_block=128)
    cols = PriorCache(out).columns()
    for i, c in enumerate(("P", "e", "omega", "M0")):
        assert np.array_equal(cols[i], rows[c])
    b = read_batch(out, ["P", "e", "s"], (10, 20), units={"s": u.m / u.s})
    assert np.array_equal(b[:, 0], rows["P"][10:20]) and np.allclose(b[:, 2], rows["s"][10:20])


def test_hdf5_min_refuses_what_it_does_not_implement(tmp_path):
    """Never a wrong answer: files in the new-style format or not HDF5 at all raise."""
    from thejoker_b200 import hdf5_min

    p = tmp_path / "v2.h5"
    p.write_bytes(b"\x89HDF\r\n\x1a\n" + bytes([2]) + b"\x00" * 100)
    with pytest.raises(NotImplementedError, match="superblock version 2"):
        hdf5_min.File(str(p))
    q = tmp_path / "junk.h5"
    q.write_bytes(b"not an hdf5 file" * 10)
    with pytest.raises(OSError):
        hdf5_min.File(str(q))
    with pytest.raises(OSError):
        tj.JokerSamples.read(str(q))


def test_samples_file_round_trip_under_the_reference_file_name(tmp_path):
    """``prior.sample(...).write('prior_samples.hdf5')`` followed by reading the
    same path back.  The container is this package's .npz whatever the name (a warning says
    so); reading detects the format from the file's bytes, so the round trip works."""
    prior = default_prior(1, sigma_K0=25.0)
    ps = prior.sample(size=257, rng=np.random.default_rng(2), return_logprobs=True)
    path = str(tmp_path / "prior_samples.hdf5")
    with pytest.warns(UserWarning, match="npz"):
        ps.write(path)
    back = tj.JokerSamples.read(path)
    for k in ("P", "e", "omega", "M0", "s", "ln_prior"):
        assert np.array_equal(np.asarray(back[k].value), np.asarray(ps[k].value))
    with pytest.raises(OSError):
        ps.write(path)  # exists


def test_poly_trend_zero_is_rejected_like_the_reference():
    """poly_trend=0 still yields the constant column (likelihood_helpers.py:21, 34-37), so
    the reference's shape check (pyx:174-179) raises; same here."""
    with pytest.raises(ValueError):
        star_spec(10, 0)


@pytest.mark.parametrize("N,pt,tol", [(30, 4, 1e-10), (40, 6, 1e-9), (24, 7, 1e-7)])
def test_host_emulated_ll_extreme_n_linear(N, pt, tol):
    """n_linear up to 8 (K + sextic trend), against the quad truth.  Raw monomials
    1, dt, dt^2.. are nearly collinear, so accuracy degrades with the order in any
    double-precision evaluation (the reference's N x N LU is worse); the tolerance per
    order documents what the L x L LDL^T delivers."""
    from oracle.oracle import OracleHelper

    spec, _, _ = star_spec(N, pt)
    assert spec["n_linear"] == 1 + pt
    chunk = prior_chunk(300, s_lognormal=(-2.0, 1.0))
    truth, _ = OracleHelper.from_spec(spec).truth_ll(chunk)
    for force_jit in (False, True):
        c = chunk.copy()
        if not force_jit:
            c[:, 4] = 0.3
            truth_c, _ = OracleHelper.from_spec(spec).truth_ll(c)
        else:
            truth_c = truth
        got = emu_marginal_ll(spec, c, force_jit=force_jit)
        assert np.max(rel_err(got, truth_c)) < tol, (N, pt, force_jit)


def test_thejoker_init_validation():
    """thejoker/tests/test_sampler.py::test_init (constructor checks need no GPU)."""
    prior = default_prior(1)
    tj.TheJoker(prior)
    tj.TheJoker(prior, rng=np.random.default_rng(1), tempfile_path="/tmp/_tjb_test")
    with pytest.raises(TypeError):
        tj.TheJoker("jsdfkj")
    with pytest.raises(TypeError):
        tj.TheJoker(prior, rng=np.random.RandomState(1))
    with pytest.raises(TypeError):
        tj.TheJoker(prior, pool="sdfks")

    class Pool:  # anything with map / close is accepted, as schwimmbad pools are
        def map(self, *a):
            pass

        def close(self):
            pass

    j = tj.TheJoker(prior, pool=Pool(), devices=[0, 1])
    assert j.devices == [0, 1] and os.path.isdir(j.tempfile_path)


@pytest.mark.parametrize("sigma", [0.5, 0.03, 0.003])
def test_conditioning_at_the_posterior_mode(sigma):
    """at the posterior mode of high-S/N data the
    likelihood is ill-conditioned in its *inputs* (d ll / d z ~ K resid / sigma^2, so the
    ~1e-13 rounding of the orbital phase alone moves ll by 1e-9..1e-8 relative at
    sigma = 3 m/s); the reference algorithm and this kernel are then both ~kappa * 1e-16
    from the quad truth and from each other, and a 1e-10 gate is not meaningful.  What is
    checked: the kernel's error stays within a small factor of the reference algorithm's."""
    from helpers import mode_chunk
    from oracle.oracle import OracleHelper

    spec, _, _ = star_spec(64, 1, sigma=sigma)
    chunk = mode_chunk(spec, 400, sigma)
    orc = OracleHelper.from_spec(spec)
    ref = orc.batch_marginal_ln_likelihood(chunk, 0)
    truth, kappa = orc.truth_ll(chunk)
    emu = emu_marginal_ll(spec, chunk)
    e_emu, e_ref = rel_err(emu, truth), rel_err(ref, truth)
    assert np.median(kappa) > 1e4
    assert np.median(e_emu) <= 3 * np.median(e_ref) + 1e-14
    assert e_emu.max() <= 5 * e_ref.max() + 1e-13
    assert e_emu.max() < 5e-16 * np.median(kappa)


# ---- counter-based prior sampler (csrc/prior_gen.cuh, host build) ---------------------------
def test_philox_known_answers():
    """Philox4x32-10 against the known-answer vectors of the Random123 distribution
    (kat_vectors: zero, all-ones and the digits-of-pi counter / key)."""
    lib = host_emulation()
    kat = [([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
           ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
           ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
            [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for ctr, key, want in kat:
        out = (ctypes.c_uint * 4)()
        lib.emu_philox4x32_10((ctypes.c_uint * 4)(*ctr), (ctypes.c_uint * 2)(*key), out)
        assert list(out) == want


def test_prior_generator_distributions():
    """The generated prior against scipy's distributions (KS), for the reference's default
    prior (UniformLog P, Kipping13Global e, uniform angles; distributions.py:17-51, 171-176;
    prior.py:437-479) and a LogNormal jitter in m/s converted to km/s; independence of
    the columns; and the contract that sample g depends only on (seed, g)."""
    from scipy import stats

    from helpers import emu_prior_rows
    from thejoker_b200.prior import LogNormal

    prior = default_prior(1, sigma_K0=25.0, P_min=2.0, P_max=1024.0,
                          s=LogNormal("s", np.log(200.0), 0.5, u.m / u.s))
    gen = prior.device_generator(987654321, u.km / u.s)
    n = 200_000
    rows = emu_prior_rows(gen, 0, n)
    assert rows[:, 0].min() >= 2 and rows[:, 0].max() <= 1024
    assert rows[:, 1].min() > 0 and rows[:, 1].max() < 1
    assert np.abs(rows[:, 2:4]).max() <= np.pi
    pv = [stats.kstest(np.log(rows[:, 0]), stats.uniform(np.log(2), np.log(512)).cdf).pvalue,
          stats.kstest(rows[:, 1], stats.beta(0.867, 3.03).cdf).pvalue,
          stats.kstest(rows[:, 2], stats.uniform(-np.pi, 2 * np.pi).cdf).pvalue,
          stats.kstest(rows[:, 3], stats.uniform(-np.pi, 2 * np.pi).cdf).pvalue,
          stats.kstest(np.log(rows[:, 4] * 1e3), stats.norm(np.log(200.0), 0.5).cdf).pvalue]
    assert min(pv) > 1e-3, pv
    cc = np.corrcoef(rows.T)
    assert np.abs(cc - np.eye(5)).max() < 0.01
    # any window of the index range reproduces the same samples; another seed does not
    assert np.array_equal(emu_prior_rows(gen, 12345, 100), rows[12345:12445])
    gen2 = prior.device_generator(987654322, u.km / u.s)
    assert not np.any(emu_prior_rows(gen2, 0, 100) == rows[:100])
    # the other Beta shapes of the reference (a >= 1 takes the unboosted gamma)
    from thejoker_b200.prior import Kipping13Long, Kipping13Short

    for cls, (a, b) in ((Kipping13Long, (1.12, 3.09)), (Kipping13Short, (0.697, 3.27))):
        pr = default_prior(1, sigma_K0=25.0, pars={"e": cls("e")})
        e = emu_prior_rows(pr.device_generator(5, u.km / u.s), 0, 100_000)[:, 1]
        assert stats.kstest(e, stats.beta(a, b).cdf).pvalue > 1e-3
    # a family without a device sampler -> no generator (the caller samples on the host)
    from thejoker_b200.prior import Distribution

    class Odd(Distribution):
        def draw(self, rng, size, **ctx):
            return rng.uniform(size=size)

    assert default_prior(1, sigma_K0=25.0, pars={"e": Odd("e", u.one)}).device_generator(1, u.km / u.s) is None


def test_prior_generator_ln_prior_rows():
    """ln_prior_rows (prior.py:388-400 restated for packed rows in internal units) equals
    the per-parameter logp sum of JokerPrior.sample(return_logprobs=True)."""
    from thejoker_b200.prior import LogNormal

    prior = default_prior(1, sigma_K0=25.0, s=LogNormal("s", np.log(200.0), 0.5, u.m / u.s))
    smp = prior.sample(size=1000, rng=np.random.default_rng(3), return_logprobs=True)
    rows = np.stack([smp["P"].to_value(u.day), smp["e"].value, smp["omega"].to_value(u.rad),
                     smp["M0"].to_value(u.rad), smp["s"].to_value(u.km / u.s)], axis=1)
    assert np.allclose(prior.ln_prior_rows(rows, u.km / u.s), smp["ln_prior"].value, rtol=1e-12,
                       atol=1e-12)
