"""Pins the CPU oracle to the reference's own compiled Cython operator.

tests/golden/ref_*.npz hold outputs of thejoker/src/fast_likelihood.pyx itself --
translated by Cython and compiled unmodified from /root/reference
(oracle/ref_build/build_ref.py), driven through CJokerHelper.__init__ and its public
methods (oracle/ref_cython.py, tests/golden/make_ref_golden.py).  Only twobody's
c_rv_from_elements is not the reference's code (third party, absent: the oracle's
restatement is linked in its place).

* against the committed vectors: always (CPU, any machine).
* against the live extension: when oracle/_ref/ is present (built in the container that
  has /root/reference; the .so also travels to the GPU box) -- there the comparison is
  exact (==), on fresh seeded inputs.
"""
import glob
import os

import numpy as np
import pytest
from helpers import prior_chunk, star_spec

from oracle import ref_cython
from oracle.oracle import OracleHelper, iterative_rejection_indices, rejection_accept

_REF_ALL = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "ref_*.npz")))
REF_REJECTION = [p for p in _REF_ALL if os.path.basename(p).startswith("ref_rejection_")]
REF_GOLDEN = [p for p in _REF_ALL if os.path.basename(p).startswith("ref_n")]  # ref_n<N>_...
SPEC_KEYS = ("t", "rv", "ivar", "t0", "trend_M", "mu", "Lambda", "K_prior_kind", "sigma_K0", "P0",
             "max_K")
MATS = ("a", "A", "Ainv", "b", "B", "Binv")


def _spec(z):
    return {k: (z[k] if z[k].ndim else z[k].item()) for k in SPEC_KEYS}


def test_reference_vectors_are_committed():
    assert len(REF_GOLDEN) >= 7


@pytest.mark.parametrize("path", REF_GOLDEN, ids=[os.path.basename(p)[4:-4] for p in REF_GOLDEN])
def test_oracle_matches_reference_vectors(path):
    """ll, every matrix the worker leaves behind, and the posterior draw (same rng stream).
    The vectors were bit-identical where they were minted; the tolerance only allows for
    libm / OpenBLAS kernel dispatch on a different CPU."""
    z = np.load(path)
    orc = OracleHelper.from_spec(_spec(z), jitter_mode=0)  # the reference ignores s (pyx:458)
    ll = orc.batch_marginal_ln_likelihood(z["chunk"])
    assert np.allclose(ll, z["ref_ll"], rtol=1e-12, atol=0)
    n_post = len(z["ref_worker_ll"])
    for i in range(n_post):
        l = orc.test_likelihood_worker(z["chunk"][i])
        assert np.isclose(l, z["ref_worker_ll"][i], rtol=1e-12, atol=0)
        for k in MATS:
            ref = z["ref_worker_" + k][i]
            assert np.allclose(getattr(orc, k), ref, rtol=1e-10, atol=1e-14 * np.abs(ref).max()), k
    n_draw = len(z["ref_samples"]) // n_post
    samples, lls = orc.batch_get_posterior_samples(z["chunk"][:n_post], n_draw,
                                                   np.random.default_rng(11))
    assert samples.shape == z["ref_samples"].shape
    assert np.allclose(samples, z["ref_samples"], rtol=1e-9, atol=1e-9)
    assert np.allclose(lls, z["ref_samples_ll"], rtol=1e-12, atol=0)


@pytest.mark.parametrize("path", REF_REJECTION,
                         ids=[os.path.basename(p)[14:-4] for p in REF_REJECTION])
def test_oracle_matches_reference_rejection_drivers(path):
    """thejoker/likelihood_helpers.py:91-229 run on the compiled reference helper vs the
    oracle's restatement of the same steps on the same Generator stream: accepted rows,
    truncation, the linear draws that follow, ln_prior / ln_likelihood bookkeeping, and
    the iterative schedule."""
    z = np.load(path)
    orc = OracleHelper.from_spec(_spec(z), jitter_mode=0)
    chunk, ln_prior = z["chunk"], z["ln_prior"]
    # (1) rejection_sample_inmem(..., n_linear_samples=2, return_all_logprobs=True)
    rng = np.random.default_rng(int(z["seed_rej"]))
    ll = orc.batch_marginal_ln_likelihood(chunk)
    assert np.allclose(ll, z["rej_lls"], rtol=1e-12, atol=0)
    good = rejection_accept(ll, rng.uniform(size=len(ll)))
    assert np.array_equal(chunk[np.repeat(good, 2)], z["rej_raw"][:, :5])
    assert np.array_equal(ln_prior[good], z["rej_ln_prior"])
    assert np.allclose(ll[good], z["rej_ln_likelihood"], rtol=1e-12, atol=0)
    smp, _ = orc.batch_get_posterior_samples(chunk[good], 2, rng)
    assert np.allclose(smp, z["rej_raw"], rtol=1e-9, atol=1e-9)
    # (2) max_posterior_samples=3
    rng = np.random.default_rng(int(z["seed_rej"]))
    good3 = rejection_accept(ll, rng.uniform(size=len(ll)), max_posterior_samples=3)
    smp, _ = orc.batch_get_posterior_samples(chunk[good3], 1, rng)
    assert np.allclose(smp, z["rej3_raw"], rtol=1e-9, atol=1e-9)
    # (3) iterative_rejection_inmem(n_requested=4, init_batch_size=256)
    rng = np.random.default_rng(int(z["seed_iter"]))
    idx, lls_seen = iterative_rejection_indices(
        lambda a, b: orc.batch_marginal_ln_likelihood(chunk[a:b]), len(chunk), rng,
        int(z["iter_n_requested"]), init_batch_size=int(z["iter_init_batch_size"]))
    assert np.array_equal(chunk[idx], z["iter_raw"][:, :5])
    assert np.array_equal(ln_prior[idx], z["iter_ln_prior"])
    assert np.allclose(lls_seen[idx], z["iter_ln_likelihood"], rtol=1e-12, atol=0)
    smp, _ = orc.batch_get_posterior_samples(chunk[idx], 1, rng)
    assert np.allclose(smp, z["iter_raw"], rtol=1e-9, atol=1e-9)


def test_reference_ignores_jitter_and_clamps_only_in_batch_ll():
    """Two quirks of the reference, read off its own outputs: (i) the jitter column does
    not enter the likelihood (pyx:458 fills s_ivar, the algebra reads ivar); (ii) only
    batch_marginal_ln_likelihood clamps Lambda_K at max_K^2 (pyx:464 vs 519-522, 569-572)."""
    z = np.load([p for p in REF_GOLDEN if "l3_jitter" in p][0])
    assert np.all(z["chunk"][:, 4] > 0)
    o0 = OracleHelper.from_spec(_spec(z), jitter_mode=0).batch_marginal_ln_likelihood(z["chunk"])
    o1 = OracleHelper.from_spec(_spec(z), jitter_mode=1).batch_marginal_ln_likelihood(z["chunk"])
    assert np.allclose(o0, z["ref_ll"], rtol=1e-12)
    assert not np.allclose(o1, z["ref_ll"], rtol=1e-3)


live = pytest.mark.skipif(not ref_cython.available(),
                          reason="oracle/_ref not built (needs /root/reference at build time)")

LIVE_CASES = [
    (16, 1, {}), (64, 1, {}), (3, 1, {"normal_K": 10.0}), (20, 1, {"n_surveys": 2}),
    (24, 2, {"n_surveys": 3}), (12, 3, {}), (40, 4, {}), (64, 1, {"sigma": 0.01}),
]


@live
@pytest.mark.parametrize("N,pt,kw", LIVE_CASES)
def test_oracle_equals_live_reference_cython(N, pt, kw):
    """Fresh seeded inputs through both, exact equality."""
    spec, _, _ = star_spec(N, pt, seed=7, **kw)
    ref = ref_cython.RefCythonHelper(spec, poly_trend=spec["n_poly"], n_offsets=spec["n_offsets"])
    orc = OracleHelper.from_spec(spec, jitter_mode=0)
    chunk = prior_chunk(300, seed=99, s_lognormal=(-1.0, 1.0))
    assert np.array_equal(ref.batch_marginal_ln_likelihood(chunk),
                          orc.batch_marginal_ln_likelihood(chunk))
    for row in chunk[:5]:
        ll_ref, mats = ref.test_likelihood_worker(row)
        assert ll_ref == orc.test_likelihood_worker(row)
        for k in MATS:
            assert np.array_equal(mats[k], getattr(orc, k)), k
    s_ref, l_ref = ref.batch_get_posterior_samples(chunk[:6], 4, np.random.default_rng(5))
    s_orc, l_orc = orc.batch_get_posterior_samples(chunk[:6], 4, np.random.default_rng(5))
    assert np.array_equal(np.asarray(s_ref), s_orc) and np.array_equal(np.asarray(l_ref), l_orc)


@live
def test_live_reference_edge_cases():
    """Empty chunk; e = 0; a zero inverse variance is a division by zero in B (pyx:317)
    and comes out non-finite in both; wrong trend_M shape raises ValueError (pyx:175-180)."""
    spec, _, _ = star_spec(10, 1, seed=3)
    ref = ref_cython.RefCythonHelper(spec, 1, 0)
    orc = OracleHelper.from_spec(spec, jitter_mode=0)
    assert ref.batch_marginal_ln_likelihood(np.zeros((0, 5))).shape == (0,)
    chunk = prior_chunk(16, seed=1)
    chunk[:, 1] = 0.0
    assert np.array_equal(ref.batch_marginal_ln_likelihood(chunk),
                          orc.batch_marginal_ln_likelihood(chunk))
    bad = dict(spec)
    bad["trend_M"] = np.ones((10, 2))
    with pytest.raises(ValueError):
        ref_cython.RefCythonHelper(bad, 1, 0)


@live
def test_shims_do_not_leak_into_the_process():
    import sys

    ref_cython.load()
    assert "astropy" not in sys.modules and "thejoker" not in sys.modules


# -- host-side helpers of the path against the reference's own functions --------------------
def _host_golden():
    import json

    with open(os.path.join(os.path.dirname(__file__), "golden", "ref_host_logic.json")) as f:
        return json.load(f)


def test_batch_tasks_matches_reference():
    """thejoker/utils.py:22-72 (outputs minted by tests/golden/make_ref_host_golden.py)."""
    from thejoker_b200.sharding import batch_tasks, shard_ranges

    from oracle.oracle import batch_tasks_ranges

    g = _host_golden()
    for c in g["batch_tasks"]:
        got = batch_tasks(c["n_tasks"], c["n_batches"], start_idx=c["start_idx"], args=["x"])
        assert [[list(t[0]), t[1], t[2]] for t in got] == c["tasks"], c
        assert [list(r) for r in batch_tasks_ranges(c["n_tasks"], c["n_batches"],
                                                    c["start_idx"])] == [t[0] for t in c["tasks"]]
        if c["start_idx"] == 0 and c["n_batches"] > 0 and c["n_tasks"] >= c["n_batches"]:
            assert [list(r) for r in shard_ranges(c["n_tasks"], c["n_batches"])] == \
                [t[0] for t in c["tasks"]]
    c = g["batch_tasks_arr"]
    got = batch_tasks(c["n_tasks"], c["n_batches"], arr=np.array(c["arr"]), start_idx=c["start_idx"])
    assert [[t[0].tolist(), t[1]] for t in got] == c["tasks"]


def test_design_matrices_match_reference():
    """thejoker/likelihood_helpers.py:8-37, 232-233."""
    from thejoker_b200.likelihood_helpers import (get_constant_term_design_matrix,
                                                  get_trend_design_matrix, ln_normal)

    class FakeData:
        def __init__(self, t, t_ref):
            self._t_bmjd, self._t_ref_bmjd = np.asarray(t, float), float(t_ref)

        def __len__(self):
            return len(self._t_bmjd)

    g = _host_golden()
    for c in g["design"]:
        data = FakeData(c["t"], c["t_ref"])
        ids = None if c["ids"] is None else np.array(c["ids"])
        assert np.array_equal(get_constant_term_design_matrix(data, ids), np.array(c["const_M"]))
        M = get_trend_design_matrix(data, ids, c["poly_trend"])
        ref = np.array(c["trend_M"]).reshape(M.shape)
        assert np.array_equal(M, ref), (c["poly_trend"], np.max(np.abs(M - ref) / np.abs(ref)))
    for c in g["ln_normal"]:
        assert np.isclose(ln_normal(c["x"], c["mu"], c["var"]), c["value"], rtol=1e-15, atol=0)


def test_samples_analysis_matches_reference():
    """thejoker/samples_analysis.py:35-135 (is_P_unimodal, max_phase_gap, phase_coverage,
    periods_spanned), outputs of the reference's own functions on seeded epochs."""
    import thejoker_b200 as tj
    from thejoker_b200 import samples_analysis as sa
    from thejoker_b200 import units as u

    for c in _host_golden()["samples_analysis"]:
        t = np.array(c["t"])
        data = tj.RVData(t, np.zeros(len(t)) * u.km / u.s, np.ones(len(t)) * u.km / u.s,
                         t_ref=c["t_ref"])
        many = tj.JokerSamples()
        many["P"] = np.array(c["P_samples"]) * u.day
        one = tj.JokerSamples()
        one["P"] = np.array([c["P"]]) * u.day
        assert bool(sa.is_P_unimodal(many, data)) == c["is_P_unimodal"]
        assert np.isclose(sa.max_phase_gap(one, data), c["max_phase_gap"], rtol=1e-12, atol=1e-15)
        assert np.isclose(sa.phase_coverage(one, data), c["phase_coverage"], rtol=0, atol=1e-15)
        assert np.isclose(sa.periods_spanned(one, data), c["periods_spanned"], rtol=1e-13)


# -- the Kepler function against twobody's own outputs stored in the reference repository ----
def _examples():
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_examples.npz"))


def _example_cases():
    """(name, dt [TCB days since t0], stored rv, list of (P, e, omega, M0, K), constant
    per-epoch term) for the three datasets of docs/examples/make-data.ipynb."""
    g = _examples()
    P, e, om, M0, K, v0 = g["single_truth"]
    yield "single", g["single_dt"], g["single_rv"], [(P, e, om, M0, K)], np.full(257, v0)
    P, e, om, M0, K, v0 = g["triple_truth1"]
    yield ("triple", g["triple_dt"], g["triple_rv"],
           [(P, e, om, M0, K), tuple(g["triple_truth2"])], np.full(257, v0))
    P, e, om, M0, K, v0 = g["survey_truth"]
    const = np.full(17, v0)
    const[int(g["survey_n1"]):] += float(g["survey_offset"])
    yield "survey", g["survey_dt"], g["survey_rv"], [(P, e, om, M0, K)], const


# float64 Julian dates resolve 4.7e-10 d: 2 pi K dt_err / P ~ 5e-10 km/s at K ~ 7, P ~ 42 d
TWOBODY_RV_TOL = 6e-10  # km/s


def test_oracle_kepler_reproduces_twobody_outputs():
    """oracle/joker_oracle.c::orc_rv_from_elements (the restatement of twobody's
    c_rv_from_elements) against radial velocities that twobody itself computed:
    docs/examples/*.ecsv hold noiseless `KeplerOrbit.radial_velocity(t)` values whose true
    elements follow from the notebook's seed (tests/golden/make_ref_examples_golden.py)."""
    import ctypes

    from oracle.oracle import load

    lib = load()
    dp = ctypes.POINTER(ctypes.c_double)
    lib.orc_rv_from_elements.restype = None
    lib.orc_rv_from_elements.argtypes = ([dp, dp, ctypes.c_int] + [ctypes.c_double] * 7
                                         + [ctypes.c_int, ctypes.c_int])
    for name, dt, rv, orbits, const in _example_cases():
        model = const.copy()
        for P, e, om, M0, K in orbits:
            t = np.ascontiguousarray(dt)
            out = np.zeros_like(t)
            lib.orc_rv_from_elements(t.ctypes.data_as(dp), out.ctypes.data_as(dp), len(t), P, K,
                                     e, om, M0, 0.0, 1e-10, 128, 0)
            model += out
        assert np.max(np.abs(model - rv)) < TWOBODY_RV_TOL, (name, np.max(np.abs(model - rv)))


def test_device_kepler_math_reproduces_twobody_outputs():
    """The same for the device Kepler function (csrc/kepler.cuh) compiled for the host."""
    import ctypes

    from helpers import host_emulation

    lib = host_emulation()
    dp = ctypes.POINTER(ctypes.c_double)
    st = (ctypes.c_int * 3)()
    for name, dt, rv, orbits, const in _example_cases():
        model = const.copy()
        for P, e, om, M0, K in orbits:
            t = np.ascontiguousarray(dt)
            z = np.zeros_like(t)
            lib.emu_design_column(P, e, om, M0, t.ctypes.data_as(dp), len(t),
                                  z.ctypes.data_as(dp), st)
            assert st[2] == 0
            model += K * z
        assert np.max(np.abs(model - rv)) < TWOBODY_RV_TOL, (name, np.max(np.abs(model - rv)))
