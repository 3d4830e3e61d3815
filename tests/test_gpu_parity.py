"""Parity of the CUDA path with the CPU oracle, through the C ABI
(libthejoker_b200.so via ctypes).  Everything here needs a B200.

Gates (BASELINE.json north_star): ll within 1e-10 relative of the reference algorithm
on identical inputs -- where the reference algorithm itself is further than that from
the exact value (ill-conditioned B), the CUDA result must be at least as close to the
quad-precision truth; accepted index sets bit-exact except samples within 1e-12 of the
threshold, which are counted."""
import glob
import os

import numpy as np
import pytest
from helpers import accept_sets_match, prior_chunk, reference_gate, rel_err, star_spec

pytestmark = pytest.mark.gpu

_ALL_NPZ = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
# ref_*.npz: minted by the reference's own compiled Cython (golden/make_ref_golden.py);
# the others by the oracle + quad truth (golden/make_golden.py)
GOLDEN = [p for p in _ALL_NPZ if not os.path.basename(p).startswith("ref_")]
REF_REJECTION = [p for p in _ALL_NPZ if os.path.basename(p).startswith("ref_rejection_")]
REF_GOLDEN = [p for p in _ALL_NPZ if os.path.basename(p).startswith("ref_n")]  # ref_n<N>_...


@pytest.fixture(scope="module")
def torch_cuda():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def make_helper(spec_args, device=0, **kw):
    import thejoker_b200 as tj
    from thejoker_b200.data_helpers import validate_prepare_data

    N, pt = spec_args[:2]
    spec, data, prior = star_spec(N, pt, **kw)
    jm = kw.get("jitter_mode", "apply")
    all_data, ids, trend_M = validate_prepare_data(data, prior.poly_trend, prior.n_offsets)
    return tj.CJokerHelper(all_data, prior, trend_M, device=device, jitter_mode=jm), spec, data, prior


def rows_to_idx(column, values):
    """Indices of `values` in a prior column of distinct numbers (accepted rows -> indices)."""
    order = np.argsort(column)
    pos = np.searchsorted(column[order], values)
    idx = order[pos]
    assert np.array_equal(column[idx], values)
    return idx


SPEC_KEYS = ("t", "rv", "ivar", "t0", "trend_M", "mu", "Lambda", "K_prior_kind", "sigma_K0", "P0",
             "max_K", "jitter_mode")


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_golden_vectors(torch_cuda, path):
    """The committed fixtures (inputs + oracle outputs) travel without the oracle."""
    import thejoker_b200 as tj

    z = np.load(path)
    spec = {k: (z[k] if z[k].ndim else z[k].item()) for k in SPEC_KEYS}
    helper = tj.CJokerHelper.from_spec(spec, device=0)
    ll = helper.batch_marginal_ln_likelihood(np.ascontiguousarray(z["chunk"]))
    reference_gate(ll, z["ll"], z["ll_truth"], label=os.path.basename(path)[:-4])
    lls, a, A = helper.posterior_aA(z["chunk"][:16])
    assert np.allclose(a, z["post_a"], rtol=1e-8, atol=1e-10)
    assert np.allclose(A, z["post_A"], rtol=1e-7, atol=1e-14)
    # accept on the fixture's own uniforms
    import torch

    ll_dev = torch.from_numpy(ll).cuda()
    key = helper.new_llmax_key()
    helper.llmax_update(ll_dev, key)
    idx, tot, near = helper.accept(ll_dev, key, uniforms=torch.from_numpy(z["uniforms"]).cuda())
    # the device set against the host rule on the device's own lls: bit-exact minus the
    # near-threshold samples; against the oracle's set: the same minus samples the
    # ll difference (<= 1e-10 relative) moves across the threshold, which are counted
    n_near, _ = accept_sets_match(idx.cpu().numpy(), ll, z["uniforms"])
    assert n_near == near
    _, n_moved = accept_sets_match(idx.cpu().numpy(), z["ll"], z["uniforms"], ll_other=ll)
    assert n_moved <= 1


@pytest.mark.parametrize("path", REF_GOLDEN, ids=[os.path.basename(p)[:-4] for p in REF_GOLDEN])
def test_reference_cython_golden_vectors(torch_cuda, path):
    """Vectors produced by the reference's own compiled fast_likelihood.pyx
    (tests/golden/make_ref_golden.py).  The reference ignores the jitter column
    (pyx:458 is a dead store), hence jitter_mode="reference"; its posterior entry points
    do not clamp Lambda_K at max_K^2 (pyx:519-522, 569-572), hence clamp=False."""
    import thejoker_b200 as tj

    z = np.load(path)
    spec = {k: (z[k] if z[k].ndim else z[k].item()) for k in SPEC_KEYS}
    spec["jitter_mode"] = 0
    helper = tj.CJokerHelper.from_spec(spec, device=0)
    chunk = np.ascontiguousarray(z["chunk"])
    ll = helper.batch_marginal_ln_likelihood(chunk)
    # quad-precision value of the same inputs (tests/golden/make_ref_truth.py): the gate is
    # 1e-10 against the reference AND against the truth; on the flat star (K = 1e-4) the
    # reference's own N x N chi2 form is up to 2e-10 off the truth on a few rows, which the
    # gate counts instead of loosening the tolerance
    name = os.path.basename(path)[:-4]
    truth = np.load(os.path.join(os.path.dirname(path), "ref_truth.npz"))[name]
    rep = reference_gate(ll, z["ref_ll"], truth, label=name)
    assert rep["n_reference_off_truth"] <= (4 if "flat" in path else 0)
    n_post = len(z["ref_worker_ll"])
    lls, a, A = helper.posterior_aA(chunk[:n_post], clamp_K=False)
    reference_gate(lls, z["ref_worker_ll"], truth[:n_post], label=name + " worker")
    assert np.allclose(a, z["ref_worker_a"], rtol=1e-8, atol=1e-10)
    assert np.allclose(A, z["ref_worker_A"], rtol=1e-7, atol=1e-14)
    # the reference's own draw: numpy multivariate_normal(a, inv(Ainv)) with the same rng
    # stream reproduces its linear parameters from the device a / A
    n_draw = len(z["ref_samples"]) // n_post
    samples, ll_rep = helper.batch_get_posterior_samples(chunk[:n_post], n_draw,
                                                         np.random.default_rng(11), draw="numpy",
                                                         clamp_K=False)
    assert samples.shape == z["ref_samples"].shape
    assert np.array_equal(samples[:, :5], z["ref_samples"][:, :5])
    scale = np.sqrt(np.repeat(np.einsum("nii->ni", z["ref_worker_A"]), n_draw, axis=0))
    assert np.max(np.abs(samples[:, 5:] - z["ref_samples"][:, 5:]) / scale) < 1e-6


def test_kepler_function_reproduces_twobody_outputs(torch_cuda):
    """tjb_design_column (the kernel's Kepler function) against radial velocities computed
    by twobody itself: the noiseless example data of the reference repository
    (tests/golden/ref_examples.npz, made by tests/golden/make_ref_examples_golden.py); and
    the marginal likelihood evaluated at the true nonlinear elements sits at the chi2
    floor (the stored rv are exact, so the residual is ~0 against the quoted errors)."""
    import thejoker_b200 as tj

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_examples.npz"))
    tol = 6e-10  # km/s: the float64 Julian dates of the files resolve 4.7e-10 d
    for name in ("single", "survey"):
        dt, rv, err = g[name + "_dt"], g[name + "_rv"], g[name + "_rv_err"]
        P, e, om, M0, K, v0 = g[name + "_truth"]
        n = len(dt)
        if name == "survey":
            n1 = int(g["survey_n1"])
            T = np.stack([np.ones(n), (np.arange(n) >= n1).astype(float)], axis=1)
            lin = np.array([v0, float(g["survey_offset"])])
            mu, Lam = np.zeros(3), np.array([0.0, 1e6, 1e6])
        else:
            T, lin = np.ones((n, 1)), np.array([v0])
            mu, Lam = np.zeros(2), np.array([0.0, 1e6])
        spec = dict(t=dt, rv=rv, ivar=1.0 / err**2, t0=0.0, trend_M=T, mu=mu, Lambda=Lam,
                    K_prior_kind=0, sigma_K0=300.0, P0=365.25, max_K=5000.0, jitter_mode=1)
        helper = tj.CJokerHelper.from_spec(spec, device=0)
        z = helper.design_column(np.array([P, e, om, M0, 0.0]))
        model = K * z + T @ lin
        assert np.max(np.abs(model - rv)) < tol, (name, np.max(np.abs(model - rv)))
        # at the truth the data are reproduced exactly, so a, the posterior mean of the
        # linear parameters, is the truth itself up to the (weak) prior's pull
        ll, a, A = helper.posterior_aA(np.array([[P, e, om, M0, 0.0]]))
        assert np.allclose(a[0], np.concatenate([[K], lin]), rtol=1e-3, atol=1e-3), a[0]
        # and a period 1 % off is far less likely (oracle: 474.6 for the 257-epoch set,
        # 10.3 for the 17 epochs of the two surveys)
        ll_off = helper.batch_marginal_ln_likelihood(np.array([[P * 1.01, e, om, M0, 0.0]]))
        assert ll[0] - ll_off[0] > (400 if name == "single" else 9)


@pytest.mark.parametrize("N,pt,kw", [(16, 1, {}), (64, 1, {}), (33, 2, {"n_surveys": 2}),
                                     (64, 3, {}), (20, 1, {"K": 1e-4})])
def test_live_reference_cython(torch_cuda, N, pt, kw):
    """The CUDA path against the reference's own compiled operator running on this box
    (oracle/_ref travels with the repo; skipped when it was not built): fresh seeded
    inputs, 4096 prior rows per star, ll within 1e-10, a / A of the posterior."""
    from oracle import ref_cython

    if not ref_cython.available():
        pytest.skip("oracle/_ref not built")
    import thejoker_b200 as tj

    spec, _, _ = star_spec(N, pt, seed=11, **kw)
    ref = ref_cython.RefCythonHelper(spec, poly_trend=spec["n_poly"], n_offsets=spec["n_offsets"])
    spec_ref_mode = dict(spec, jitter_mode=0)  # the reference ignores the jitter column
    helper = tj.CJokerHelper.from_spec(spec_ref_mode, device=0)
    chunk = prior_chunk(4096, seed=77, s_lognormal=(-1.0, 1.0))
    want = ref.batch_marginal_ln_likelihood(chunk)
    got = helper.batch_marginal_ln_likelihood(chunk)
    from oracle.oracle import OracleHelper

    truth, _ = OracleHelper.from_spec(spec_ref_mode).truth_ll(chunk)
    flat = "K" in kw  # chi2 cancels to ~1e-12 of its terms in the reference itself
    rep = reference_gate(got, want, truth, label=f"live N={N} pt={pt} {kw}")
    assert rep["n_reference_off_truth"] <= (0.02 * len(chunk) if flat else 0)
    lls, a, A = helper.posterior_aA(chunk[:8], clamp_K=False)
    for i in range(8):
        ll_ref, mats = ref.test_likelihood_worker(chunk[i])
        assert abs(lls[i] - ll_ref) <= 1e-10 * abs(ll_ref) or \
            abs(ll_ref - truth[i]) >= 0.9 * abs(lls[i] - ll_ref)
        assert abs(lls[i] - truth[i]) <= 1e-10 * abs(truth[i])
        assert np.allclose(a[i], mats["a"], rtol=1e-8, atol=1e-10)
        assert np.allclose(A[i], mats["A"], rtol=1e-7, atol=1e-14)
    # the single-row entry point leaves the reference's public attributes behind
    ll1 = helper.test_likelihood_worker(chunk[0])
    ll_ref, mats = ref.test_likelihood_worker(chunk[0])
    assert abs(ll1 - truth[0]) <= 1e-10 * abs(truth[0])
    assert np.allclose(helper.a, mats["a"], rtol=1e-8, atol=1e-10)
    assert np.allclose(helper.A, mats["A"], rtol=1e-7, atol=1e-14)
    assert np.allclose(helper.Ainv, mats["Ainv"], rtol=1e-6)
    assert np.allclose(helper.b, mats["b"], rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("path", REF_REJECTION,
                         ids=[os.path.basename(p)[14:-4] for p in REF_REJECTION])
def test_reference_rejection_driver_vectors(torch_cuda, path):
    """TheJoker.rejection_sample / iterative_rejection_sample (device ll, fused max, device
    PCG64 uniforms, compaction, draws) against what the reference's own
    likelihood_helpers.py:91-229 returned on its compiled helper for the same star, prior
    rows and Generator seed (tests/golden/make_ref_golden.py)."""
    import sys

    import thejoker_b200 as tj

    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from make_golden import CASES as GOLDEN_CASES

    z = np.load(path)
    N, pt, sl, kw = GOLDEN_CASES[os.path.basename(path)[14:-4]]
    spec, data, prior = star_spec(N, pt, **kw)
    # the star regenerated here is the one the reference was given
    assert np.array_equal(spec["t"], z["t"]) and np.array_equal(spec["rv"], z["rv"])
    assert np.array_equal(spec["ivar"], z["ivar"]) and np.array_equal(spec["Lambda"], z["Lambda"])
    chunk = np.ascontiguousarray(z["chunk"])

    def packed(samples):
        return samples.pack(nonlinear_only=False)[0]

    def close(got, ref):
        assert got.shape == ref.shape, (got.shape, ref.shape)
        assert np.array_equal(got[:, :5], ref[:, :5])           # the accepted prior rows
        scale = np.abs(ref[:, 5:]).max(axis=0) + 1e-300
        assert np.max(np.abs(got[:, 5:] - ref[:, 5:]) / scale) < 1e-6

    mk = lambda seed: tj.TheJoker(prior, rng=np.random.default_rng(seed), devices=[0],
                                  jitter_mode="reference", draw="numpy")
    smp, lls = mk(int(z["seed_rej"])).rejection_sample(data, chunk, n_linear_samples=2,
                                                       return_all_logprobs=True, in_memory=True)
    name = os.path.basename(path)[:-4]
    truth = np.load(os.path.join(os.path.dirname(path), "ref_truth.npz"))[name]
    rep = reference_gate(lls, z["rej_lls"], truth, label=name)
    assert rep["n_reference_off_truth"] <= (16 if "flat" in path else 0)
    close(packed(smp), z["rej_raw"])
    smp = mk(int(z["seed_rej"])).rejection_sample(data, chunk, max_posterior_samples=3,
                                                  in_memory=True)
    close(packed(smp), z["rej3_raw"])
    smp = mk(int(z["seed_iter"])).iterative_rejection_sample(
        data, chunk, int(z["iter_n_requested"]), init_batch_size=int(z["iter_init_batch_size"]),
        in_memory=True)
    close(packed(smp), z["iter_raw"])


CASES = [
    ((16, 1), None, {}),
    ((64, 1), None, {}),
    ((64, 2), (-2.0, 1.0), {}),
    ((256, 1), None, {}),
    ((3, 1), None, {"normal_K": 10.0}),
    ((20, 1), None, {"n_surveys": 2}),
    ((24, 2), None, {"n_surveys": 3}),
    ((12, 3), None, {}),
    ((64, 1), None, {"K": 1e-4}),
    ((64, 1), None, {"sigma": 0.01}),
]


@pytest.mark.parametrize("args,sl,kw", CASES)
def test_ll_parity_host_path(torch_cuda, oracle_lib, args, sl, kw):
    """batch_marginal_ln_likelihood (host chunk in, host ll out) vs oracle + truth."""
    helper, spec, _, _ = make_helper(args, **kw)
    n = 1 << 14 if args[0] <= 64 else 1 << 12
    chunk = prior_chunk(n, s_lognormal=sl)
    orc = oracle_lib.OracleHelper.from_spec(spec)
    ref = orc.batch_marginal_ln_likelihood(chunk, n_threads=0)
    truth, kappa = orc.truth_ll(chunk)
    ll = helper.batch_marginal_ln_likelihood(chunk)
    rep = reference_gate(ll, ref, truth, label=f"{args} {kw}")
    # how many samples lean on the truth is bounded: a handful on the flat star (the
    # reference's uncentred chi2 cancels) and with high-order trend columns, none otherwise
    assert rep["n_reference_off_truth"] <= (0.005 * n if (kw.get("K") == 1e-4 or args[1] >= 2) else 0), rep
    # ragged / tiny / empty inputs
    for m in (0, 1, 31, 33, 257):
        part = helper.batch_marginal_ln_likelihood(np.ascontiguousarray(chunk[:m]))
        assert part.shape == (m,)
        if sl is not None and m == 1:
            # a one-row chunk has a uniform jitter by construction and takes the
            # constant-jitter kernel (weights folded into the table in long double); the
            # same row inside a mixed chunk goes through the per-sample-jitter kernel
            assert np.allclose(part, ll[:m], rtol=1e-12, atol=0)
        else:
            assert np.array_equal(part, ll[:m])
    with pytest.raises(ValueError):
        helper.batch_marginal_ln_likelihood(chunk.astype(np.float32))
    with pytest.raises(ValueError):
        helper.batch_marginal_ln_likelihood(chunk[:, :4])


def test_random_configurations(torch_cuda, oracle_lib):
    """Seeded sweep over star configurations the fixed cases do not reach: odd epoch counts
    (remainder loops of both kernel shapes), every n_linear from 2 to 8 (parameter-block,
    shared-memory rows; 1024 x 2 and 640 x 4 shapes), several surveys, noise levels, both
    jitter kernels and both jitter modes -- against the quad-precision truth (1e-10) and,
    where the oracle is cheap, the reference algorithm."""
    rng = np.random.default_rng(20261017)
    n_cfg = 0
    for _ in range(28):
        n_surveys = int(rng.integers(1, 4))
        pt = int(rng.integers(1, 8 - (n_surveys - 1) - 1 + 1))  # L = 1 + pt + n_surveys - 1 <= 8
        N = int(rng.integers(max(4, 2 * n_surveys + 2), 140))
        kw = {"n_surveys": n_surveys, "seed": int(rng.integers(1, 10_000)),
              "sigma": float(rng.choice([0.05, 0.5, 3.0]))}
        if rng.random() < 0.25:
            kw["K"] = 1e-4
        jm = "reference" if rng.random() < 0.2 else "apply"
        sl = (-2.0, 1.0) if rng.random() < 0.5 else None
        helper, spec, _, _ = make_helper((N, pt), jitter_mode=jm, **kw)
        assert helper.n_linear == pt + n_surveys
        chunk = prior_chunk(768, seed=int(rng.integers(1, 10_000)), s_lognormal=sl)
        ll = helper.batch_marginal_ln_likelihood(chunk)
        orc = oracle_lib.OracleHelper.from_spec(spec)
        truth, _ = orc.truth_ll(chunk)
        assert np.isfinite(ll).all()
        # high-order trend columns (unscaled dt^k up to k = 6 over 150 d, as the reference
        # builds them) condition the L x L solve itself: any double-precision evaluation --
        # the host build of this code measures 8e-11 at degree 2 with 50 m/s errors, 4e-10 at
        # degree 3, up to 9e-8 at degree 6 -- so the gate widens with the polynomial degree
        tol = {1: 1e-10, 2: 1e-10, 3: 5e-10, 4: 2e-9}.get(pt, 5e-7)
        assert np.max(rel_err(ll, truth)) < tol, (N, pt, n_surveys, kw, jm, sl)
        if pt <= 2:
            ref = orc.batch_marginal_ln_likelihood(chunk[:256], n_threads=0)
            reference_gate(ll[:256], ref, truth[:256], label=f"random N={N} pt={pt} ns={n_surveys}")
        n_cfg += 1
    assert n_cfg == 28


def test_epoch_rows_kernels_agree(torch_cuda, oracle_lib):
    """The likelihood kernel that reads the epoch rows from its parameter block (uniform
    loads, the default when the table fits) and the one that stages them in shared memory
    (long tables, n_linear > 4) give identical bits; tables beyond the parameter block's
    capacity take the shared-memory kernel by themselves and match the oracle."""
    from thejoker_b200 import _lib

    lib = _lib.load()
    try:
        for shape, sl in (((64, 1), None), ((64, 2), (-2.0, 1.0)), ((21, 1), None),
                          ((256, 1), None), ((3, 3), (-2.0, 1.0))):
            helper, spec, _, _ = make_helper(shape)
            chunk = prior_chunk(20_000, s_lognormal=sl)
            _lib.check(lib.tjb_set_epoch_rows_mode(0))
            a = helper.batch_marginal_ln_likelihood(chunk)
            _lib.check(lib.tjb_set_epoch_rows_mode(1))
            b = helper.batch_marginal_ln_likelihood(chunk)
            assert np.array_equal(a, b), shape
    finally:
        _lib.check(lib.tjb_set_epoch_rows_mode(0))
    # 1000 epochs at L = 2: 4000 doubles > kParamRowDoubles (3072) -> shared-memory rows;
    # 8000 epochs: 250 KB > one CTA's shared memory -> rows read from global memory (the
    # reference has no limit on n_times)
    for N, n in ((1000, 2048), (8000, 96)):
        helper, spec, _, _ = make_helper((N, 1))
        chunk = prior_chunk(n)
        got = helper.batch_marginal_ln_likelihood(chunk)
        orc = oracle_lib.OracleHelper.from_spec(spec)
        truth, _ = orc.truth_ll(chunk)
        assert np.max(rel_err(got, truth)) < 1e-10, N


def test_ll_device_paths_agree(torch_cuda, oracle_lib):
    """SoA / AoS device entry points and the host entry point give identical bits, for
    both kernels (constant jitter folded into the table vs per-sample jitter)."""
    torch = torch_cuda
    helper, spec, _, _ = make_helper((64, 2))
    n = 50_000
    for sl, s_const in ((None, None), (None, 0.3), ((-2.0, 1.0), None)):
        chunk = prior_chunk(n, s_lognormal=sl, s_const=s_const)
        host = helper.batch_marginal_ln_likelihood(chunk)
        dev = torch.from_numpy(chunk).cuda()
        cols = [dev[:, i].contiguous() for i in range(5)]
        key = helper.new_llmax_key()
        if sl is None:
            soa = helper.marginal_ll_soa(*cols[:4], s=None, s_const=s_const or 0.0, llmax_key=key)
            aos = helper.marginal_ll_aos(dev, uniform_s=True)
            soa_j = helper.marginal_ll_soa(*cols[:4], s=cols[4])
            assert np.max(rel_err(soa_j.cpu().numpy(), host)) < 1e-12
        else:
            soa = helper.marginal_ll_soa(*cols[:4], s=cols[4], llmax_key=key)
            aos = helper.marginal_ll_aos(dev, uniform_s=False)
        assert np.array_equal(soa.cpu().numpy(), host)
        assert np.array_equal(aos.cpu().numpy(), host)
        hc = [np.ascontiguousarray(chunk[:, i]) for i in range(5)]
        host_cols = helper.marginal_ln_likelihood_columns(
            *hc[:4], s=hc[4] if sl is not None else None, s_const=s_const or 0.0)
        assert np.array_equal(host_cols, host)
        assert helper.llmax_value(key) == host.max()
        ref = oracle_lib.OracleHelper.from_spec(spec).batch_marginal_ln_likelihood(chunk[:4096], 0)
        assert np.max(rel_err(host[:4096], ref)) < 1e-9


@pytest.mark.parametrize("n", [1, 1000, (1 << 18) + 5, 3 * (1 << 18) + 77, (1 << 21) + 12345])
def test_host_streaming_paths(torch_cuda, n):
    """Host columns streamed through the GPU (pageable memory via the page-locked ring,
    page-locked memory directly; ll to the host or left on the device with its max) give
    the bits of the resident-column kernel, for ragged sizes around the slice size."""
    torch = torch_cuda
    helper, spec, _, _ = make_helper((16, 1))
    for sl in (None, (-2.0, 1.0)):
        chunk = prior_chunk(n, s_lognormal=sl)
        hc = [np.ascontiguousarray(chunk[:, i]) for i in range(5)]
        s = hc[4] if sl is not None else None
        dev = [torch.from_numpy(c).cuda() for c in hc]
        key0 = helper.new_llmax_key()
        ref = helper.marginal_ll_soa(*dev[:4], s=dev[4] if sl is not None else None,
                                     llmax_key=key0).cpu().numpy()
        # pageable in, pageable out
        assert np.array_equal(helper.marginal_ln_likelihood_columns(*hc[:4], s=s), ref)
        # page-locked in and out
        pin = [torch.from_numpy(c).pin_memory().numpy() for c in hc]
        out = torch.empty(n, dtype=torch.float64).pin_memory().numpy()
        helper.marginal_ln_likelihood_columns(*pin[:4], s=pin[4] if sl is not None else None,
                                              out=out)
        assert np.array_equal(out, ref)
        # pageable in, ll stays on the device, running max in the key
        key = helper.new_llmax_key()
        d_ll = helper.marginal_ll_host_columns(*hc[:4], s=s, llmax_key=key)
        assert np.array_equal(d_ll.cpu().numpy(), ref)
        assert helper.llmax_value(key) == helper.llmax_value(key0) == ref.max()
        # read-only (memory-mapped style) sources
        ro = [c.copy() for c in hc]
        for c in ro:
            c.flags.writeable = False
        d_ll = helper.marginal_ll_host_columns(*ro[:4], s=ro[4] if sl is not None else None)
        assert np.array_equal(d_ll.cpu().numpy(), ref)


def test_host_streaming_from_concurrent_threads(torch_cuda):
    """Handles are per star but the staging pool is per device: calls from several host
    threads (each with its own handle) serialise on it and stay correct; so does the
    block cache when handles are created and destroyed concurrently."""
    from concurrent.futures import ThreadPoolExecutor

    import thejoker_b200 as tj

    specs = [star_spec(n_t, 1, seed=5 + n_t)[0] for n_t in (8, 16, 24, 32)]
    chunk = prior_chunk((1 << 19) + 3)
    hc = [np.ascontiguousarray(chunk[:, i]) for i in range(4)]
    refs = []
    for sp in specs:
        h = tj.CJokerHelper.from_spec(sp, device=0)
        refs.append(h.marginal_ln_likelihood_columns(*hc))
        del h

    def work(i):
        outs = []
        for _ in range(3):
            h = tj.CJokerHelper.from_spec(specs[i], device=0)
            outs.append(h.marginal_ll_host_columns(*hc).cpu().numpy())
            outs.append(h.batch_marginal_ln_likelihood(chunk[:70_000]))
            del h
        return outs

    with ThreadPoolExecutor(4) as ex:
        results = list(ex.map(work, range(4)))
    for ref, outs in zip(refs, results):
        for k, o in enumerate(outs):
            assert np.array_equal(o, ref[:len(o)]), k


def test_engine_streamed_equals_resident(torch_cuda):
    """DeviceEngine over host columns: streaming (default) and resident modes give the
    same ll, max, accepted indices and rows, also for sub-ranges (iterative sampler)."""
    from thejoker_b200.sharding import DeviceEngine

    helper, spec, data, prior = make_helper((16, 1), K=1e-4)
    import thejoker_b200 as tj

    make = lambda d: tj.CJokerHelper.from_spec(spec, device=d)
    n = 300_000
    chunk = prior_chunk(n)
    cols = [np.ascontiguousarray(chunk[:, i]) for i in range(4)] + [None]
    outs = []
    for resident in (False, True):
        eng = DeviceEngine(make, cols, devices=[0], resident=resident)
        eng.compute_ll(0, 70_000)
        eng.compute_ll(70_000, n)
        idx, tot, near = eng.accept(np.random.default_rng(3), max_keep=500)
        outs.append((eng.download_ll(0, n), eng.max_value(), idx, tot, eng.rows(idx[:50])))
    for a, b in zip(outs[0], outs[1]):
        assert np.array_equal(a, b)
    assert outs[0][3] > 0


def test_jitter_modes(torch_cuda, oracle_lib):
    """jitter_mode='reference' ignores s like the reference's Cython does
    (fast_likelihood.pyx:458); 'apply' matches the intended semantics
    (src/tests/py_likelihood.py:17-29)."""
    chunk = prior_chunk(4096, s_lognormal=(0.0, 1.0))
    chunk0 = chunk.copy()
    chunk0[:, 4] = 0
    h_ref, spec_ref, _, _ = make_helper((32, 1), jitter_mode="reference")
    assert np.array_equal(h_ref.batch_marginal_ln_likelihood(chunk),
                          h_ref.batch_marginal_ln_likelihood(chunk0))
    o = oracle_lib.OracleHelper.from_spec(spec_ref).batch_marginal_ln_likelihood(chunk, 0)
    assert np.max(rel_err(h_ref.batch_marginal_ln_likelihood(chunk), o)) < 1e-10
    h_app, spec_app, _, _ = make_helper((32, 1), jitter_mode="apply")
    o = oracle_lib.OracleHelper.from_spec(spec_app).batch_marginal_ln_likelihood(chunk, 0)
    assert np.max(rel_err(h_app.batch_marginal_ln_likelihood(chunk), o)) < 1e-10


def test_design_column_and_solver_tail(torch_cuda, oracle_lib):
    """z = cos(f+omega) + e cos(omega) per epoch vs the Newton restatement, including
    e up to 0.999; the fixed-work solver must report no non-converged epochs."""
    helper, spec, _, _ = make_helper((64, 1))
    orc = oracle_lib.OracleHelper.from_spec(spec)
    rng = np.random.default_rng(5)
    chunk = prior_chunk(400)
    chunk[:150, 1] = rng.uniform(0.8, 0.999, 150)
    chunk[150:160, 1] = 0.0
    extra = 0
    for row in chunk:
        z, st = helper.design_column(row, return_stats=True)
        zo = orc.design_column(row)
        assert st[2] == 0
        extra += st[1]
        assert np.max(np.abs(z - zo)) < 3e-12 / (1 - row[1]) ** 2
    print(f"\nextra FP64 Householder passes over {len(chunk) * 64} epochs: {extra}")


def test_extreme_eccentricity_and_phase_stay_finite(torch_cuda):
    """e -> 1 (down to 1 - 1e-7) with mean anomalies down to 1e-9 rad, and phases beyond the
    FP32 stage's range (P = 0.05 d over a 10 000 d baseline): the safeguarded extra passes
    (kepler.cuh::solve_extra_passes) converge for every epoch and every ll is finite -- a
    single NaN would poison the max of the prior cache and nothing would be accepted.  The
    same sweep runs on the host build of the device code in
    tests/test_host_logic.py::test_kepler_solver_extreme_cases (there also against mpmath)."""
    rng = np.random.default_rng(3)
    for span_periods, P_min in ((3.0, 2.0), (193.0, 0.05)):
        helper, spec, _, _ = make_helper((64, 1), t_span_periods=span_periods)
        n = 600
        chunk = prior_chunk(n)
        chunk[:, 0] = np.exp(rng.uniform(np.log(P_min), np.log(1000), n))
        chunk[:, 1] = 1 - 10 ** rng.uniform(-7, -1, n)
        # a tiny mean anomaly at the first epoch for half of the rows
        dt0 = spec["t"][0] - spec["t0"]
        tiny = 10 ** rng.uniform(-9, -2, n // 2) * rng.choice([-1, 1], n // 2)
        chunk[: n // 2, 3] = 2 * np.pi * dt0 / chunk[: n // 2, 0] - tiny
        not_conv = 0
        for row in chunk[:200]:
            z, st = helper.design_column(row, return_stats=True)
            assert np.isfinite(z).all(), row
            not_conv += st[2]
        assert not_conv == 0
        ll = helper.batch_marginal_ln_likelihood(chunk)
        assert np.isfinite(ll).all()
        assert helper.solver_stats(reset=True)["not_converged"] == 0


def test_posterior_aA_and_draws(torch_cuda, oracle_lib):
    """(a, A) vs oracle (dsysv / inverse of Ainv) as in
    test_fast_likelihood.py::test_likelihood_helpers; draws: same RNG consumption as
    rng.multivariate_normal, right mean / covariance."""
    for args, kw in (((3, 1), {"normal_K": 1.0}), ((64, 2), {}), ((20, 1), {"n_surveys": 2})):
        helper, spec, _, _ = make_helper(args, **kw)
        orc = oracle_lib.OracleHelper.from_spec(spec)
        chunk = prior_chunk(16)
        for row in chunk:
            ll = helper.test_likelihood_worker(row)
            ll_o = orc.test_likelihood_worker(row)
            assert np.abs(ll) < 1e8
            assert np.isclose(ll, ll_o, rtol=1e-9)
            assert np.allclose(helper.a, orc.a, rtol=1e-8, atol=1e-10)
            assert np.allclose(helper.A, orc.A, rtol=1e-7, atol=1e-14)
            assert np.allclose(helper.b, orc.b)
            a_t, A_t = orc.truth_aA(row)
            assert np.allclose(helper.a, a_t, rtol=1e-9, atol=1e-11)
        # reference-identical draws through numpy
        r1, r2 = np.random.default_rng(3), np.random.default_rng(3)
        s_np, _ = helper.batch_get_posterior_samples(chunk, 2, r1, draw="numpy")
        s_or, _ = orc.batch_get_posterior_samples(chunk, 2, r2)
        assert np.allclose(s_np, s_or, rtol=1e-6, atol=1e-9)
        # device draws: stream consumption equals multivariate_normal's
        r3 = np.random.default_rng(3)
        s_dev, lls = helper.batch_get_posterior_samples(chunk, 2, r3, draw="device")
        assert np.array_equal(r1.standard_normal(3), r3.standard_normal(3))
        assert s_dev.shape == s_np.shape and np.array_equal(s_dev[:, :5], s_np[:, :5])
    # distribution check on one row
    helper, spec, _, _ = make_helper((16, 1))
    row = prior_chunk(1)
    _, a, A = helper.posterior_aA(row)
    draws, _ = helper.batch_get_posterior_samples(row, 200_000, np.random.default_rng(0),
                                                  draw="device")
    x = draws[:, 5:]
    assert np.allclose(x.mean(0), a[0], atol=5 * np.sqrt(np.diag(A[0]) / len(x)))
    C = np.cov(x.T)
    se = np.sqrt((np.outer(np.diag(A[0]), np.diag(A[0])) + A[0] ** 2) / len(x))
    assert np.all(np.abs(C - A[0]) < 5 * se)


def test_pcg64_uniforms_bit_exact(torch_cuda):
    helper, _, _, _ = make_helper((8, 1))
    for seed, n, off in ((42, 100_000, 0), (7, 65_537, 12_345), (1, 1, 0), (5, 3_000_000, 999_999)):
        rng = np.random.default_rng(seed)
        dev = helper.pcg64_uniform(rng, n, offset=off).cpu().numpy()
        want = np.random.default_rng(seed).random(n + off)[off:]
        assert np.array_equal(dev, want)


def test_accept_counts_nonfinite_lls(torch_cuda):
    """likelihood_helpers.py:173-176 tests np.isfinite over every ll before it accepts: the
    accept kernel counts NaN / +-inf lls (a -inf does not show in the max), and the
    in-memory iterative sampler returns the reference's RuntimeError object when there is one."""
    import thejoker_b200 as tj
    from thejoker_b200 import units as u
    from thejoker_b200.synthetic import make_noisy_data

    torch = torch_cuda
    helper, spec, data, prior = make_helper((16, 1))
    n = 100_000
    ll = np.random.default_rng(0).normal(-50.0, 3.0, n)
    ll[[5, 70_000]] = -np.inf
    ll_dev = torch.from_numpy(ll).cuda()
    key = helper.new_llmax_key()
    helper.llmax_update(ll_dev, key)
    idx, tot, _ = helper.accept(ll_dev, key, rng=np.random.default_rng(1))
    assert helper.last_nonfinite == 2 and tot > 0
    ll[123] = np.inf
    ll[99_999] = np.nan
    helper.accept(torch.from_numpy(ll).cuda(), key, rng=np.random.default_rng(1))
    assert helper.last_nonfinite == 4
    helper.accept(torch.from_numpy(ll[1000:60_000].copy()).cuda(), key, rng=np.random.default_rng(1))
    assert helper.last_nonfinite == 0
    # a zero-weight data set whose ll is -inf for some samples is hard to build from the
    # public API; the sampler's rule is exercised through its stats instead
    joker = tj.TheJoker(prior, rng=np.random.default_rng(3), devices=[0])
    chunk = prior_chunk(4096)
    out = joker.iterative_rejection_sample(data, chunk, n_requested_samples=8, in_memory=True)
    assert not isinstance(out, Exception) and joker.last_stats["n_nonfinite"] == 0


@pytest.mark.parametrize("flat", [False, True])
def test_accept_bit_exact(torch_cuda, oracle_lib, flat):
    """where(exp(ll - max) > u)[0][:max_keep]: bit-exact index set vs numpy on the same
    ll and the same uniforms (host-supplied and device PCG64)."""
    torch = torch_cuda
    helper, spec, _, _ = make_helper((16, 1), K=1e-4 if flat else None)
    n = 300_000
    chunk = prior_chunk(n)
    ll = helper.batch_marginal_ln_likelihood(chunk)
    ll_dev = torch.from_numpy(ll).cuda()
    key = helper.new_llmax_key()
    helper.llmax_update(ll_dev, key)
    assert helper.llmax_value(key) == ll.max()
    rng = np.random.default_rng(7)
    uu = np.random.default_rng(7).uniform(size=n)
    for max_keep in (None, 1, 5, 100):
        want = oracle_lib.rejection_accept(ll, uu, max_keep)
        near = oracle_lib.near_threshold_count(ll, uu)
        idx, tot, n_near = helper.accept(ll_dev, key, uniforms=torch.from_numpy(uu).cuda(),
                                         max_keep=max_keep)
        idx2, tot2, _ = helper.accept(ll_dev, key, rng=rng, max_keep=max_keep)
        assert n_near == near
        # same ll, same uniforms: the sets are identical minus the near-threshold samples
        # (the comparison is always made; near-threshold indices are removed, not the test)
        assert accept_sets_match(idx.cpu().numpy(), ll, uu, max_keep) == (near, 0)
        assert accept_sets_match(idx2.cpu().numpy(), ll, uu, max_keep) == (near, 0)
        assert tot == tot2 and abs(tot - len(oracle_lib.rejection_accept(ll, uu))) <= near
        if near == 0:
            assert np.array_equal(idx.cpu().numpy(), want)
    if flat:
        assert tot > 10  # test_sampler.py:156: uninformative data keeps many samples
    # offsets: shard [lo, hi) of a larger stream reports global indices
    lo = 100_003
    idx, tot, _ = helper.accept(ll_dev[lo:], key, rng=rng, rng_offset=lo, index_base=lo)
    want = oracle_lib.rejection_accept(ll, uu)
    assert np.array_equal(idx.cpu().numpy(), want[want >= lo])
    # NaN / +inf propagate through the max like numpy.max (appendix B)
    bad = ll.copy()
    bad[17] = np.nan
    k2 = helper.new_llmax_key()
    helper.llmax_update(torch.from_numpy(bad).cuda(), k2)
    assert np.isnan(helper.llmax_value(k2))
    idx, tot, _ = helper.accept(torch.from_numpy(bad).cuda(), k2, rng=rng)
    assert tot == 0 and idx.numel() == 0


def test_rejection_sample_api(torch_cuda, oracle_lib):
    """TheJoker.rejection_sample end to end (thejoker/tests/test_sampler.py:128-173):
    count bounds, logprobs, determinism, and index-set identity with the oracle fed the
    same prior samples and the same generator."""
    import thejoker_b200 as tj
    from helpers import default_prior
    from thejoker_b200.synthetic import make_data

    for poly_trend in (1, 2):
        prior = default_prior(poly_trend, sigma_K0=25.0, P_min=5.0, P_max=500.0)
        data, _ = make_data(8, rng=np.random.default_rng(11))
        flat_data, _ = make_data(8, rng=np.random.default_rng(11), K=1e-4)
        prior_samples = prior.sample(size=16384, return_logprobs=True, rng=np.random.default_rng(1))
        joker = tj.TheJoker(prior, rng=np.random.default_rng(42))
        for in_memory in (True, False):
            samples = joker.rejection_sample(data, prior_samples, in_memory=in_memory)
            assert 0 < len(samples) < 10
            samples = joker.rejection_sample(data, prior_samples, return_logprobs=True,
                                             in_memory=in_memory)
            assert len(samples) > 0 and "ln_likelihood" in samples and "ln_prior" in samples
            samples = joker.rejection_sample(flat_data, prior_samples, in_memory=in_memory)
            assert len(samples) > 10
            assert np.isfinite(samples.ln_unmarginalized_likelihood(flat_data)).all()
        samples, lls = joker.rejection_sample(flat_data, prior_samples, return_all_logprobs=True)
        assert len(lls) == len(prior_samples)
        all_Ps, all_Ks = [], []
        for i in range(4):
            joker = tj.TheJoker(prior, rng=np.random.default_rng(42))
            s = joker.rejection_sample(flat_data, prior_samples)
            all_Ps.append(s["P"].value)
            all_Ks.append(s["K"].value)
        for i in range(1, 4):
            assert np.array_equal(all_Ps[0], all_Ps[i]) and np.array_equal(all_Ks[0], all_Ks[i])
        # identity with the reference algorithm on identical inputs
        joker = tj.TheJoker(prior, rng=np.random.default_rng(42))
        helper = joker._make_joker_helper(flat_data)
        chunk, _ = prior_samples.pack(units=helper.internal_units, names=helper.packed_order)
        got = joker.rejection_sample(flat_data, prior_samples, in_memory=True, max_posterior_samples=64)
        orc = oracle_lib.OracleHelper.from_spec(helper.spec)
        ref_ll = orc.batch_marginal_ln_likelihood(chunk, 0)
        uu = np.random.default_rng(42).uniform(size=len(chunk))
        good = oracle_lib.rejection_accept(ref_ll, uu, 64)
        # the accepted rows against the oracle's, minus the near-threshold samples
        n_near, _ = accept_sets_match(rows_to_idx(chunk[:, 0], got["P"].value), ref_ll, uu, 64)
        if n_near == 0:
            assert np.array_equal(got["P"].value, chunk[good, 0])
        assert joker.last_stats["n_near_threshold"] == oracle_lib.near_threshold_count(ref_ll, uu)
        with pytest.raises(ValueError):
            joker.rejection_sample(data, prior_samples, n_prior_samples=10**6)
        s_int = tj.TheJoker(prior, rng=np.random.default_rng(3)).rejection_sample(flat_data, 4096)
        assert len(s_int) > 0


def test_iterative_rejection_sample_api(torch_cuda, oracle_lib):
    """thejoker/tests/test_sampler.py:176-211 plus identity with the restated driver."""
    import thejoker_b200 as tj
    from helpers import default_prior
    from thejoker_b200.synthetic import make_data

    prior = default_prior(1, sigma_K0=25.0, P_min=5.0, P_max=500.0)
    data, _ = make_data(3, rng=np.random.default_rng(2))
    prior_samples = prior.sample(size=10_000, return_logprobs=True, rng=np.random.default_rng(1))
    for in_memory in (True, False):
        joker = tj.TheJoker(prior, rng=np.random.default_rng(42))
        s = joker.iterative_rejection_sample(data, prior_samples, n_requested_samples=4,
                                             in_memory=in_memory)
        assert len(s) > 1
        s = joker.iterative_rejection_sample(data, prior_samples, n_requested_samples=4,
                                             return_logprobs=True, in_memory=in_memory)
        assert len(s) > 1 and "ln_prior" in s
    runs = []
    for i in range(3):
        joker = tj.TheJoker(prior, rng=np.random.default_rng(42))
        s = joker.iterative_rejection_sample(data, prior_samples, n_requested_samples=4,
                                             randomize_prior_order=True)
        runs.append((s["P"].value, s["K"].value))
    assert all(np.array_equal(runs[0][0], r[0]) and np.array_equal(runs[0][1], r[1]) for r in runs)
    # identity of the accepted set with the restated in-memory driver
    joker = tj.TheJoker(prior, rng=np.random.default_rng(9))
    helper = joker._make_joker_helper(data)
    chunk, _ = prior_samples.pack(units=helper.internal_units, names=helper.packed_order)
    got = joker.iterative_rejection_sample(data, prior_samples, n_requested_samples=16,
                                           in_memory=True, growth_factor=8)
    orc = oracle_lib.OracleHelper.from_spec(helper.spec)
    idx, all_lls = oracle_lib.iterative_rejection_indices(
        lambda a, b: orc.batch_marginal_ln_likelihood(chunk[a:b], 0), len(chunk),
        np.random.default_rng(9), 16, growth_factor=8, safety_factor=1)
    assert np.array_equal(got["P"].value, chunk[idx, 0])
    assert joker.last_stats["n_ll_evaluated"] == len(all_lls)
    with pytest.raises(ValueError):
        joker.iterative_rejection_sample(data, prior_samples, n_requested_samples=1000)


def test_marginal_ln_likelihood_api(torch_cuda, oracle_lib):
    """thejoker/tests/test_sampler.py:81-108: JokerSamples / file / packed array in,
    len(ll) == len(prior_samples); values equal the helper's."""
    import thejoker_b200 as tj
    from helpers import default_prior
    from thejoker_b200.synthetic import make_data

    prior = default_prior(2, sigma_K0=25.0, P_min=1.0, P_max=365.0)
    data, _ = make_data(8, rng=np.random.default_rng(0))
    ps = prior.sample(size=100, rng=np.random.default_rng(0))
    joker = tj.TheJoker(prior)
    ll = joker.marginal_ln_likelihood(data, ps)
    assert len(ll) == len(ps)
    helper = joker._make_joker_helper(data)
    chunk, _ = ps.pack(units=helper.internal_units, names=helper.packed_order)
    assert np.array_equal(ll, helper.batch_marginal_ln_likelihood(np.ascontiguousarray(chunk)))
    assert np.array_equal(ll, joker.marginal_ln_likelihood(data, ps, in_memory=True))


def test_multi_survey_offsets(torch_cuda, oracle_lib):
    """v0_offsets: list of RVData with an indicator column per extra survey."""
    import thejoker_b200 as tj

    spec, data, prior = star_spec(24, 1, n_surveys=3)
    joker = tj.TheJoker(prior, rng=np.random.default_rng(0))
    helper = joker._make_joker_helper(data)
    assert helper.n_linear == 4
    chunk = prior_chunk(8192)
    ll = helper.batch_marginal_ln_likelihood(chunk)
    ref = oracle_lib.OracleHelper.from_spec(helper.spec).batch_marginal_ln_likelihood(chunk, 0)
    assert np.max(rel_err(ll, ref)) < 1e-10
    s = joker.rejection_sample(data, chunk, in_memory=True)
    assert list(s.keys())[:8] == ["P", "e", "omega", "M0", "s", "K", "v0", "dv0_1"]


@pytest.mark.parametrize("args,kw", [((24, 2), {"n_surveys": 3}), ((40, 1), {}), ((12, 3), {})])
def test_unmarginalized_likelihood_device(torch_cuda, oracle_lib, args, kw):
    """tjb_unmarginalized_ll vs the formula of samples.py:611-632 evaluated in numpy on the
    oracle's design column, and (no offsets) vs the host method of JokerSamples."""
    import thejoker_b200 as tj

    helper, spec, data, prior = make_helper(args, **kw)
    orc = oracle_lib.OracleHelper.from_spec(spec)
    chunk = prior_chunk(200, s_lognormal=(-1.0, 1.0))
    rows, _ = helper.batch_get_posterior_samples(chunk, 2, np.random.default_rng(3))
    got = helper.ln_unmarginalized_likelihood(rows)
    T = np.asarray(spec["trend_M"]).reshape(len(spec["t"]), -1)
    want = np.empty(len(rows))
    for i, r in enumerate(rows):
        model = r[5] * orc.design_column(r[:5]) + T @ r[6:]
        var = 1.0 / spec["ivar"] + r[4] ** 2
        want[i] = np.sum(-0.5 * ((spec["rv"] - model) ** 2 / var + np.log(2 * np.pi * var)))
    assert np.max(rel_err(got, want)) < 1e-10
    assert helper.ln_unmarginalized_likelihood(rows[:0]).shape == (0,)
    with pytest.raises(ValueError):
        helper.ln_unmarginalized_likelihood(rows[:, :-1])
    samples = tj.JokerSamples.unpack(rows, helper.internal_units, t_ref=helper.data.t_ref,
                                     poly_trend=prior.poly_trend, n_offsets=prior.n_offsets)
    dev = samples.ln_unmarginalized_likelihood(helper.data, helper=helper)
    assert np.array_equal(dev, got)
    if not kw:
        # without a helper: K z + polynomial trend about the samples' t_ref (no offsets)
        plain = samples.ln_unmarginalized_likelihood(data)
        assert np.max(rel_err(plain, got)) < 1e-12


def test_thompson_black_hole_example(torch_cuda):
    """End to end on real data: the reference's Thompson-black-hole notebook (11 TRES + 3
    APOGEE epochs, its priors, s ~ LogNormal, a survey offset) through rejection_sample
    with a device-drawn prior recovers the published orbit (Thompson et al. 2019:
    P = 83.205 d, K = 44.615 km/s, e = 0.005, f(M) = 0.766 Msun)."""
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(__file__)), "examples"))
    import thompson_black_hole as ex

    res = ex.run(log2_n=26)
    pub = ex.PUBLISHED
    for name in ("tres", "joint"):
        s = ex.summarize(res[name]["samples"])
        print(name, s, res[name]["stats"])
        assert s["n"] >= 1, name
        assert abs(s["P"][0] - pub["P"]) < 0.6 and s["P"][1] < 0.6, (name, s)
        assert abs(s["K"][0] - pub["K"]) < 0.8, (name, s)
        assert s["e_max"] < 0.05, (name, s)
        assert abs(s["fM"][0] - pub["fM"]) < 0.04, (name, s)
    # the joint fit (1900-day baseline) pins the period much better
    assert abs(ex.summarize(res["joint"]["samples"])["P"][0] - pub["P"]) < 0.25


def test_large_n_properties(torch_cuda):
    """Size-independent checks at BASELINE scale (2^24 samples, N=64): the fused max
    equals the max of the written ll, permutation equivariance, finite everywhere."""
    torch = torch_cuda
    helper, spec, _, _ = make_helper((64, 1))
    n = 1 << 24
    g = torch.Generator(device="cuda").manual_seed(123)
    U = torch.rand(4, n, dtype=torch.float64, device="cuda", generator=g)
    P = torch.exp(U[0] * np.log(512.0) + np.log(2.0))
    e = torch.distributions.Beta(torch.tensor(0.867, dtype=torch.float64, device="cuda"),
                                 torch.tensor(3.03, dtype=torch.float64, device="cuda")).sample((n,))
    e = e.contiguous()
    om, M0 = (U[2] * 2 - 1) * np.pi, (U[3] * 2 - 1) * np.pi
    key = helper.new_llmax_key()
    ll = helper.marginal_ll_soa(P, e, om, M0, llmax_key=key)
    assert torch.isfinite(ll).all()
    assert helper.llmax_value(key) == ll.max().item()
    perm = torch.randperm(n, device="cuda")[: 1 << 20]
    ll2 = helper.marginal_ll_soa(P[perm].contiguous(), e[perm].contiguous(),
                                 om[perm].contiguous(), M0[perm].contiguous())
    assert torch.equal(ll2, ll[perm])
    # omega -> omega + 2 pi and M0 -> M0 + 2 pi leave the model unchanged
    ll3 = helper.marginal_ll_soa(P[: 1 << 20].contiguous(), e[: 1 << 20].contiguous(),
                                 (om[: 1 << 20] + 2 * np.pi).contiguous(),
                                 (M0[: 1 << 20] - 2 * np.pi).contiguous())
    rel = ((ll3 - ll[: 1 << 20]).abs() / ll[: 1 << 20].abs()).max().item()
    assert rel < 1e-9


def test_device_prior_sampling(torch_cuda):
    """rejection_sample(data, prior_samples=<int>): the prior is drawn on the GPU by the
    library's counter-based sampler (csrc/prior_gen.cuh).  The
    CUDA sampler against its host build and against scipy's distributions; the likelihood
    kernel that generates the samples in registers against the same kernel reading the
    materialised columns (bit-equal); determinism; logprobs; a non-constant jitter prior."""
    from scipy import stats

    import thejoker_b200 as tj
    from helpers import default_prior, emu_prior_rows
    from thejoker_b200 import units as u
    from thejoker_b200.prior import LogNormal
    from thejoker_b200.synthetic import make_data

    torch = torch_cuda
    prior = default_prior(1, sigma_K0=25.0, P_min=5.0, P_max=500.0)
    n = 400_000
    cols, s, lp = prior.sample_device(n, "cuda:0", 1234, u.km / u.s, return_logprobs=True)
    dev = np.stack([c.cpu().numpy() for c in cols], axis=1)
    P, e = dev[:, 0], dev[:, 1]
    assert s == 0.0 and P.min() >= 5 and P.max() <= 500 and e.min() > 0 and e.max() < 1
    # the same (seed, index) -> the same sample as the host build of the sampler (libm vs
    # CUDA log / exp / sincospi differ by ulps; an accept decision of the gamma sampler
    # can flip on such a difference once in ~1e15 draws)
    host = emu_prior_rows(prior.device_generator(1234, u.km / u.s), 0, n)
    assert np.max(np.abs(dev - host[:, :4]) / np.maximum(np.abs(host[:, :4]), 1e-300)) < 1e-12
    assert stats.kstest(np.log(P), stats.uniform(np.log(5), np.log(100)).cdf).pvalue > 1e-3
    assert stats.kstest(e, stats.beta(0.867, 3.03).cdf).pvalue > 1e-3
    for j in (2, 3):
        assert stats.kstest(dev[:, j], stats.uniform(-np.pi, 2 * np.pi).cdf).pvalue > 1e-3
    # a window of the index range is the same samples
    cols2, _, _ = prior.sample_device(1000, "cuda:0", 1234, u.km / u.s, index0=77_000)
    assert np.array_equal(cols2[1].cpu().numpy(), e[77_000:78_000])
    lp_host = (prior.pars["P"].logp(P) + prior.pars["e"].logp(e) + prior.pars["omega"].logp(dev[:, 2])
               + prior.pars["M0"].logp(dev[:, 3]))
    assert np.allclose(lp, lp_host, rtol=1e-12, atol=1e-12)

    # generated-in-kernel ll == ll over the materialised columns, bit for bit; constant
    # jitter (folded into the table) and a LogNormal jitter in m/s (per-sample kernel)
    flat, _ = make_data(8, rng=np.random.default_rng(11), K=1e-4)
    prior_s = default_prior(1, sigma_K0=25.0, P_min=5.0, P_max=500.0,
                            s=LogNormal("s", np.log(200.0), 0.5, u.m / u.s))
    for pr in (prior, prior_s):
        helper = tj.TheJoker(pr, devices=[0])._make_joker_helper(flat)
        gen = pr.device_generator(99, helper.internal_units["s"])
        m, i0 = 100_003, 5_000_000_000  # a window beyond 2^32: the counter is 64-bit
        c, sv, _ = pr.sample_device(m, "cuda:0", 99, helper.internal_units["s"], index0=i0)
        key_a, key_b = helper.new_llmax_key(), helper.new_llmax_key()
        ll_a = helper.marginal_ll_generated(gen, i0, m, llmax_key=key_a)
        ll_b = helper.marginal_ll_soa(*c, s=None if not hasattr(sv, "shape") else sv,
                                      s_const=sv if not hasattr(sv, "shape") else 0.0,
                                      llmax_key=key_b)
        assert torch.equal(ll_a, ll_b) and torch.equal(key_a, key_b)
        idx = np.array([0, 17, m - 1], dtype=np.int64) + i0
        rows = helper.prior_rows(gen, idx)
        assert np.array_equal(rows[:, 1], c[1].cpu().numpy()[idx - i0])
        if hasattr(sv, "shape"):
            assert np.array_equal(rows[:, 4], sv.cpu().numpy()[idx - i0])

    runs = []
    for _ in range(2):
        joker = tj.TheJoker(prior, rng=np.random.default_rng(5))
        smp = joker.rejection_sample(flat, 1 << 18, max_posterior_samples=128,
                                     return_logprobs=True)
        runs.append(smp)
        assert 10 < len(smp) <= 128 and np.isfinite(smp["ln_prior"].value).all()
        assert np.isfinite(smp["ln_likelihood"].value).all()
    assert np.array_equal(runs[0]["P"].value, runs[1]["P"].value)
    assert np.array_equal(runs[0]["K"].value, runs[1]["K"].value)
    # the accepted set is what the materialised path accepts with the same uniforms
    joker = tj.TheJoker(prior, rng=np.random.default_rng(5))
    seed = int(np.random.default_rng(5).integers(0, 2**62))
    c, sv, _ = prior.sample_device(1 << 18, "cuda:0", seed, u.km / u.s)
    chunk = np.stack([t.cpu().numpy() for t in c] + [np.zeros(1 << 18)], axis=1)
    rng = np.random.default_rng(5)
    rng.integers(0, 2**62)
    smp_b = tj.TheJoker(prior, rng=rng).rejection_sample(flat, chunk, max_posterior_samples=128,
                                                         in_memory=True)
    assert np.array_equal(runs[0]["P"].value, smp_b["P"].value)
    # non-constant jitter prior, in m/s, through the per-sample-jitter kernel
    smp = tj.TheJoker(prior_s, rng=np.random.default_rng(5)).rejection_sample(flat, 1 << 16)
    sv = smp["s"].to_value(u.km / u.s)
    assert len(smp) > 10 and 0.02 < np.median(sv) < 2.0


def test_multistar_driver_matches_per_star(torch_cuda):
    """MultiStarJoker (BASELINE configs[4] in miniature): many
    stars, ragged epoch counts, two surveys per star, one shared prior cache.  Each star
    must reproduce TheJoker.rejection_sample run on it alone with the same child RNG."""
    import thejoker_b200 as tj
    from helpers import default_prior
    from thejoker_b200 import units as u
    from thejoker_b200.prior import Normal
    from thejoker_b200.synthetic import make_noisy_data

    rng = np.random.default_rng(3)
    prior = default_prior(1, sigma_K0=25.0, P_min=2.0, P_max=1024.0,
                          v0_offsets=[Normal("dv0_1", 0.0, 5.0, u.km / u.s)])
    ps = prior.sample(size=1 << 16, rng=np.random.default_rng(1))
    stars = []
    for i in range(6):
        n = int(np.clip(rng.poisson(20), 8, 40))
        full, _ = make_noisy_data(n, seed=100 + i, K=[None, 1e-4][i % 2])
        cut = int(rng.integers(3, n - 3))
        stars.append([tj.RVData(full._t_bmjd[:cut], full.rv[:cut], full.rv_err[:cut]),
                      tj.RVData(full._t_bmjd[cut:], (full.rv.value[cut:] + 3.0) * u.km / u.s,
                                full.rv_err[cut:])])
    ms = tj.MultiStarJoker(prior, ps, rng=np.random.default_rng(77), devices=[0])
    out = ms.rejection_sample(stars, max_posterior_samples=64, return_logprobs=True)
    assert len(out) == 6 and ms.engine == "native"
    # the Python-driven loop gives the same samples, likelihoods and statistics
    mp = tj.MultiStarJoker(prior, ps, rng=np.random.default_rng(77), devices=[0], engine="python")
    outp = mp.rejection_sample(stars, max_posterior_samples=64, return_logprobs=True)
    for a, b, sa, sb in zip(out, outp, ms.last_stats, mp.last_stats):
        assert sa == sb
        for k in ("P", "e", "omega", "M0", "s", "K", "v0", "dv0_1", "ln_likelihood"):
            va, vb = (np.asarray(getattr(x[k], "value", x[k])) for x in (a, b))
            assert np.array_equal(va, vb), k
    # several draws per accepted sample, one star in flight, more slots than stars
    for slots in (1, 16):
        m2 = tj.MultiStarJoker(prior, ps, rng=np.random.default_rng(5), devices=[0],
                               streams_per_device=slots)
        o2 = m2.rejection_sample(stars, max_posterior_samples=16, n_linear_samples=3)
        p2 = tj.MultiStarJoker(prior, ps, rng=np.random.default_rng(5), devices=[0],
                               engine="python").rejection_sample(
            stars, max_posterior_samples=16, n_linear_samples=3)
        for a, b in zip(o2, p2):
            assert len(a) == len(b) and len(a) % 3 == 0
            for k in ("P", "K", "v0", "dv0_1"):
                assert np.array_equal(a[k].value, b[k].value), (slots, k)
    seqs = np.random.default_rng(77).bit_generator._seed_seq.spawn(6)
    for i, star in enumerate(stars):
        child = np.random.Generator(np.random.PCG64(seqs[i]))
        ref = tj.TheJoker(prior, rng=child, draw="device").rejection_sample(
            star, ps, in_memory=True, max_posterior_samples=64)
        assert len(ref) == len(out[i]) > 0
        for k in ("P", "e", "K", "v0", "dv0_1"):
            assert np.array_equal(ref[k].value, out[i][k].value), (i, k)
        assert ms.last_stats[i]["n_accepted"] >= len(out[i])


def test_edge_cases(torch_cuda, oracle_lib):
    """many epochs (shared-memory opt-in above 48 KB), more epochs
    than shared memory holds (rows read from global memory: the reference has no limit on
    n_times), e = 0, tiny and huge periods, solver statistics."""
    import thejoker_b200 as tj
    from thejoker_b200 import _lib

    # N = 2000 epochs: 64 KB epoch table
    spec, _, _ = star_spec(2000, 1)
    helper = tj.CJokerHelper.from_spec(spec, device=0)
    chunk = prior_chunk(512)
    ll = helper.batch_marginal_ln_likelihood(chunk)
    truth, _ = oracle_lib.OracleHelper.from_spec(spec).truth_ll(chunk[:64])
    assert np.max(rel_err(ll[:64], truth)) < 1e-10
    # N = 8000: 256 KB > 227 KB of shared memory -> the global-memory-rows kernel
    spec_big, _, _ = star_spec(8000, 1)
    big = tj.CJokerHelper.from_spec(spec_big, device=0)
    ll_big = big.batch_marginal_ln_likelihood(chunk)
    truth_big, _ = oracle_lib.OracleHelper.from_spec(spec_big).truth_ll(chunk[:16])
    assert np.isfinite(ll_big).all() and np.max(rel_err(ll_big[:16], truth_big)) < 1e-10
    # circular orbits, extreme periods, extreme eccentricities
    helper, spec, _, _ = make_helper((32, 1))
    orc = oracle_lib.OracleHelper.from_spec(spec)
    odd = prior_chunk(64)
    odd[:16, 1] = 0.0
    odd[16:24, 0] = 0.05       # 3000 cycles over the baseline
    odd[24:32, 0] = 1e6        # a sliver of one orbit
    odd[32:48, 1] = np.linspace(0.95, 0.9995, 16)
    helper.solver_stats(reset=True)
    ll = helper.batch_marginal_ln_likelihood(odd)
    truth, _ = orc.truth_ll(odd)
    assert np.isfinite(ll).all()
    assert np.max(rel_err(ll, truth)) < 1e-9
    st = helper.solver_stats()
    assert st["not_converged"] == 0 and st["extra_fp64_passes"] > 0
    # invalid elements give NaN, like the reference's arithmetic would, never a hang
    bad = prior_chunk(40)
    bad[3, 1] = 1.2
    bad[5, 0] = -3.0
    bad[7, 0] = np.nan
    out = helper.batch_marginal_ln_likelihood(bad)
    assert np.isnan(out[3]) and np.isnan(out[7]) and np.isfinite(np.delete(out, [3, 5, 7])).all()


def test_prior_cache_path(torch_cuda, tmp_path):
    """rejection_sample(data, "<cache dir>"): the file path of the reference
    (thejoker.py:243-255) on the native SoA cache."""
    import thejoker_b200 as tj
    from helpers import default_prior
    from thejoker_b200 import units as u
    from thejoker_b200.cache import write_prior_cache
    from thejoker_b200.synthetic import make_data

    prior = default_prior(1, sigma_K0=25.0, P_min=5.0, P_max=500.0)
    flat, _ = make_data(8, rng=np.random.default_rng(11), K=1e-4)
    ps = prior.sample(size=30_000, return_logprobs=True, rng=np.random.default_rng(1))
    path = write_prior_cache(ps, str(tmp_path / "cache"))
    a = tj.TheJoker(prior, rng=np.random.default_rng(4)).rejection_sample(
        flat, ps, return_logprobs=True, n_prior_samples=20_000)
    b = tj.TheJoker(prior, rng=np.random.default_rng(4)).rejection_sample(
        flat, path, return_logprobs=True, n_prior_samples=20_000)
    assert len(a) == len(b) > 10
    for k in ("P", "K", "ln_prior", "ln_likelihood"):
        assert np.array_equal(a[k].value, b[k].value)
    # the file path under the reference's own file name: JokerSamples.write (this package's
    # container, whatever the extension) and a file in the reference's HDF5 layout
    # (tests/hdf5_writer.py), both told apart by their first bytes
    import json
    import warnings

    from hdf5_writer import write_reference_style_hdf5

    npz_path = str(tmp_path / "prior_samples.hdf5")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", UserWarning)
        ps.write(npz_path)
    hdr = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ref_table_headers.json")))
    rows = np.zeros(len(ps), dtype=[(c, "<f8") for c in ("P", "e", "omega", "M0", "s", "ln_prior")])
    for c in ("P", "e", "omega", "M0"):
        rows[c] = ps[c].value
    rows["s"] = ps["s"].to_value(u.m / u.s)  # the stored header declares s in m/s
    rows["ln_prior"] = ps["ln_prior"].value
    h5_path = write_reference_style_hdf5(str(tmp_path / "ref_layout.hdf5"), rows,
                                         [ln.encode() for ln in hdr["prior_samples"]],
                                         chunk_rows=4096, two_level=True)
    for pth in (npz_path, h5_path):
        c = tj.TheJoker(prior, rng=np.random.default_rng(4)).rejection_sample(
            flat, pth, return_logprobs=True, n_prior_samples=20_000)
        for k in ("P", "K", "ln_prior", "ln_likelihood"):
            assert np.array_equal(a[k].value, c[k].value), (pth, k)


def test_baseline_config3_jitter_trend(torch_cuda, oracle_lib):
    """BASELINE.json configs[2] in miniature: N=64, jitter as a nonlinear parameter
    (LogNormal) + poly_trend=2, rejection sampling; accepted set identical to the
    correct-jitter oracle fed the same samples and the same generator."""
    import thejoker_b200 as tj
    from helpers import default_prior
    from thejoker_b200 import units as u
    from thejoker_b200.prior import LogNormal
    from thejoker_b200.synthetic import make_noisy_data

    prior = default_prior(2, sigma_K0=30.0, s=LogNormal("s", -2.0, 1.0, u.km / u.s))
    data, _ = make_noisy_data(64, seed=42, K=1e-4)
    ps = prior.sample(size=1 << 16, rng=np.random.default_rng(1))
    assert not ps._uniform_s
    joker = tj.TheJoker(prior, rng=np.random.default_rng(42))
    helper = joker._make_joker_helper(data)
    chunk, _ = ps.pack(units=helper.internal_units, names=helper.packed_order)
    got = joker.rejection_sample(data, ps, in_memory=True, max_posterior_samples=256)
    orc = oracle_lib.OracleHelper.from_spec(helper.spec)
    assert helper.spec["jitter_mode"] == 1
    ref = orc.batch_marginal_ln_likelihood(chunk, 0)
    uu = np.random.default_rng(42).uniform(size=len(chunk))
    good = oracle_lib.rejection_accept(ref, uu, 256)
    got_idx = rows_to_idx(chunk[:, 0], got["P"].value)
    n_near, _ = accept_sets_match(got_idx, ref, uu, 256)
    assert np.array_equal(got["s"].value, chunk[got_idx, 4])
    if n_near == 0:
        assert np.array_equal(got["P"].value, chunk[good, 0])
    assert list(got.keys())[:8] == ["P", "e", "omega", "M0", "s", "K", "v0", "v1"]
    ll = joker.marginal_ln_likelihood(data, ps)
    truth, _ = orc.truth_ll(chunk[:4096])
    reference_gate(ll[:4096], ref[:4096], truth, label="config3 jitter + trend")


def test_baseline_config4_iterative_n256(torch_cuda, oracle_lib):
    """BASELINE.json configs[3] in miniature: iterative_rejection_sample, 256 epochs,
    flat data; identical to the restated file-path driver (safety_factor 4) run on the
    oracle's ll (small prior library: the O(N^3) oracle costs ~10 ms per sample here)."""
    import thejoker_b200 as tj
    from helpers import default_prior
    from thejoker_b200.synthetic import make_noisy_data

    prior = default_prior(1, sigma_K0=30.0)
    data, _ = make_noisy_data(256, seed=42, K=1e-4)
    ps = prior.sample(size=6000, rng=np.random.default_rng(1))
    joker = tj.TheJoker(prior, rng=np.random.default_rng(3))
    helper = joker._make_joker_helper(data)
    chunk, _ = ps.pack(units=helper.internal_units, names=helper.packed_order)
    got = joker.iterative_rejection_sample(data, ps, n_requested_samples=24, growth_factor=16,
                                           in_memory=False)
    orc = oracle_lib.OracleHelper.from_spec(helper.spec)
    idx, all_lls = oracle_lib.iterative_rejection_indices(
        lambda a, b: orc.batch_marginal_ln_likelihood(chunk[a:b], 0), len(chunk),
        np.random.default_rng(3), 24, growth_factor=16, safety_factor=4)
    assert 0 < len(got) == len(idx) <= 24
    assert np.array_equal(got["P"].value, chunk[idx, 0])
    assert joker.last_stats["n_ll_evaluated"] == len(all_lls)


def test_full_size_parity_config2(torch_cuda, oracle_lib):
    """BASELINE.md section 5 at full size, as a slice of BASELINE.json configs[1]: N = 64
    noisy epochs, default prior, a 2^24-sample prior drawn by the library's generator,
    marginal ll + accept on the GPU over all of it -- and the reference's own compiled
    operator (oracle/_ref, one process per host core like its pool.map; the bit-identical
    C restatement if it was not built) over the 2^20 samples [5 * 2^20, 6 * 2^20) of the
    same prior, with the same uniforms (global PCG64 offsets).  Reports max relative
    difference, the count beyond 1e-10 and the count where the reference itself is off the
    quad truth; the accepted set inside the slice must equal the reference's rule minus
    near-threshold samples (mirrors thejoker/src/tests/test_fast_likelihood.py:20-90 at the
    north star's size)."""
    import bench
    import thejoker_b200 as tj
    from thejoker_b200 import units as u
    from thejoker_b200.helper import extract_spec, prior_sample_device

    torch = torch_cuda
    n_total, w_lo, w = 1 << 24, 5 << 20, 1 << 20
    all_data, prior, trend_M = bench.make_star()
    helper = tj.CJokerHelper(all_data, prior, trend_M, device=0)
    gen = prior.device_generator(20261017, u.km / u.s)
    key = helper.new_llmax_key()
    ll = helper.marginal_ll_generated(gen, 0, n_total, llmax_key=key)
    rng = np.random.default_rng(7)
    idx, tot, near = helper.accept(ll, key, rng=rng, max_keep=n_total)
    idx = idx.cpu().numpy()
    llmax = helper.llmax_value(key)
    # the slice, materialised for the CPU arm
    cols = prior_sample_device(gen, w_lo, w, 0, with_s=False)
    chunk = np.ascontiguousarray(np.stack([c.cpu().numpy() for c in cols] + [np.zeros(w)], axis=1))
    arm = bench.cpu_arm(bench.host_cores())
    ref = arm.ll(chunk)
    arm.close()
    got = ll[w_lo:w_lo + w].cpu().numpy()
    r = rel_err(got, ref)
    sel = np.where(r > 2e-11)[0]
    orc = oracle_lib.OracleHelper.from_spec(extract_spec(all_data, prior, trend_M))
    truth = got.copy()
    if len(sel):  # quad truth only where it can matter (it is slow)
        truth[sel], _ = orc.truth_ll(np.ascontiguousarray(chunk[sel]))
    rep = reference_gate(got[sel], ref[sel], truth[sel], label=f"2^20 slice of 2^24 ({arm.kind})") \
        if len(sel) else dict(n_over_tol_vs_ref=0, n_reference_off_truth=0)
    print(f"2^20 parity vs {arm.kind}: max rel {r.max():.3e}, > 1e-10: {(r > 1e-10).sum()}, "
          f"reference off truth: {rep['n_reference_off_truth']}, accepted in 2^24: {tot}, near: {near}")
    assert (r > 1e-10).sum() == rep["n_over_tol_vs_ref"] <= 16
    # accept inside the slice: same uniforms (u of global sample g = g-th double of the stream)
    bg = np.random.PCG64()
    bg.state = rng.bit_generator.state
    bg.advance(w_lo)
    uu = np.random.Generator(bg).random(w)
    in_w = idx[(idx >= w_lo) & (idx < w_lo + w)]
    n_near, n_moved = accept_sets_match(in_w, ref, uu, ll_other=got, llmax=llmax, index_base=w_lo)
    print(f"accepted in slice: {len(in_w)}, near-threshold: {n_near}, moved by the ll difference: {n_moved}")
    assert n_moved <= 1
    # ... and the rows behind the accepted indices are the generator's
    rows = helper.prior_rows(gen, in_w[:64])
    assert np.array_equal(rows, chunk[in_w[:64] - w_lo])
