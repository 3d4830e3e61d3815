"""Small end-to-end pass over every kernel, for compute-sanitizer (memcheck / racecheck /
initcheck) on the GPU box:  compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
[multistar]   (with the argument: only the native multi-star loop)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import thejoker_b200 as tj  # noqa: E402
from helpers import default_prior, prior_chunk, star_spec  # noqa: E402
from thejoker_b200.synthetic import make_data  # noqa: E402

only_multistar = len(sys.argv) > 1 and sys.argv[1] == "multistar"
for N, pt, kw in (() if only_multistar else ((7, 1, {}), (16, 2, {"n_surveys": 2}), (33, 3, {}))):
    spec, data, prior = star_spec(N, pt, **kw)
    helper = tj.CJokerHelper.from_spec(spec, device=0)
    for n, sl in ((1, None), (1000, None), (4097, (-2.0, 1.0))):
        chunk = prior_chunk(n, s_lognormal=sl)
        ll = helper.batch_marginal_ln_likelihood(chunk)
        assert np.isfinite(ll).all()
        dev = torch.from_numpy(chunk).cuda()
        key = helper.new_llmax_key()
        out = helper.marginal_ll_aos(dev, uniform_s=sl is None, llmax_key=key)
        rng = np.random.default_rng(1)
        idx, tot, near = helper.accept(out, key, rng=rng, max_keep=50)
        uu = torch.rand(n, dtype=torch.float64, device="cuda")
        idx2, tot2, _ = helper.accept(out, key, uniforms=uu)
        helper.posterior_aA(chunk[:5])
        rows, _ = helper.batch_get_posterior_samples(chunk[:3], 2, rng)
        assert np.isfinite(helper.ln_unmarginalized_likelihood(rows)).all()
        helper.design_column(chunk[0])
    helper.pcg64_uniform(np.random.default_rng(0), 100_001, offset=17)
    if N == 7:
        # host-streaming paths over several ring slots (pageable source, staged)
        big = prior_chunk(2 * (1 << 18) + 1000)
        hc = [np.ascontiguousarray(big[:, i]) for i in range(4)]
        d_ll = helper.marginal_ll_host_columns(*hc, llmax_key=helper.new_llmax_key())
        h_ll = helper.marginal_ln_likelihood_columns(*hc)
        assert np.array_equal(d_ll.cpu().numpy(), h_ll)
        assert np.array_equal(helper.batch_marginal_ln_likelihood(big), h_ll)
if not only_multistar:
    # the other two sources of the epoch rows: staged in shared memory, read from global memory
    from thejoker_b200 import _lib

    lib = _lib.load()
    spec, data, prior = star_spec(33, 1)
    helper = tj.CJokerHelper.from_spec(spec, device=0)
    chunk = prior_chunk(700)
    a = helper.batch_marginal_ln_likelihood(chunk)
    _lib.check(lib.tjb_set_epoch_rows_mode(1))
    assert np.array_equal(a, helper.batch_marginal_ln_likelihood(chunk))
    _lib.check(lib.tjb_set_epoch_rows_mode(0))
    spec, data, prior = star_spec(8000, 1)
    big_helper = tj.CJokerHelper.from_spec(spec, device=0)
    assert np.isfinite(big_helper.batch_marginal_ln_likelihood(prior_chunk(96))).all()
if not only_multistar:
    prior = default_prior(1, sigma_K0=25.0, P_min=5.0, P_max=500.0)
    flat, _ = make_data(8, rng=np.random.default_rng(11), K=1e-4)
    ps = prior.sample(size=20_000, return_logprobs=True, rng=np.random.default_rng(1))
    joker = tj.TheJoker(prior, rng=np.random.default_rng(42))
    joker.rejection_sample(flat, ps, in_memory=True)
    joker.iterative_rejection_sample(flat, ps, n_requested_samples=8, in_memory=True,
                                     growth_factor=16)
    joker.rejection_sample(flat, 30_000)
# native multi-star loop: several slot threads, ragged stars, two surveys each
from thejoker_b200 import units as u  # noqa: E402
from thejoker_b200.prior import Normal  # noqa: E402
from thejoker_b200.synthetic import make_noisy_data  # noqa: E402

prior2 = default_prior(1, sigma_K0=25.0, v0_offsets=[Normal("dv0_1", 0.0, 5.0, u.km / u.s)])
ps2 = prior2.sample(size=9_001, rng=np.random.default_rng(1))
stars = []
for i in range(7):
    full, _ = make_noisy_data(9 + 3 * i, seed=100 + i, K=[None, 1e-4][i % 2])
    stars.append([tj.RVData(full._t_bmjd[:4], full.rv[:4], full.rv_err[:4]),
                  tj.RVData(full._t_bmjd[4:], full.rv[4:], full.rv_err[4:])])
out = tj.MultiStarJoker(prior2, ps2, rng=np.random.default_rng(3), devices=[0],
                        streams_per_device=3).rejection_sample(stars, max_posterior_samples=32,
                                                               n_linear_samples=2)
assert len(out) == 7 and all(len(o) % 2 == 0 for o in out)
print("sanitize smoke ok")
