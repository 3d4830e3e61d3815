"""Generates tests/golden/*.npz: small input / output vectors for the hot path.

Source of the numbers: the CPU oracle (oracle/joker_oracle.c, the literal double
restatement of thejoker/src/fast_likelihood.pyx with scipy's LAPACK -- itself pinned
bit-for-bit to the reference's compiled Cython, see make_ref_golden.py and
tests/test_ref_pinning.py) with the jitter APPLIED (the reference ignores it), and the
quad-precision truth (oracle/joker_truth.c).  These fixtures are self-minted; the
reference-minted ones are ref_*.npz.  What they pin is (i) the oracle against
regressions, (ii) the CUDA path against oracle + truth on machines where only the
fixtures travel.

Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import prior_chunk, star_spec  # noqa: E402
from oracle.oracle import OracleHelper, rejection_accept  # noqa: E402

CASES = {
    # name: (n_times, poly_trend, jitter (mu, sigma) or None, kwargs)
    "n16_l2": (16, 1, None, {}),
    "n64_l2": (64, 1, None, {}),
    "n64_l3_jitter": (64, 2, (-2.0, 1.0), {}),
    "n3_normalK": (3, 1, None, {"normal_K": 10.0}),
    "n20_l3_offsets": (20, 1, None, {"n_surveys": 2}),
    "n12_l4": (12, 3, None, {}),
    "n64_l2_flat": (64, 1, None, {"K": 1e-4}),
}
SPEC_KEYS = ("t", "rv", "ivar", "t0", "trend_M", "mu", "Lambda", "K_prior_kind", "sigma_K0", "P0",
             "max_K", "jitter_mode")


def main(n=2048):
    here = os.path.dirname(os.path.abspath(__file__))
    for name, (N, pt, sl, kw) in CASES.items():
        spec, _, _ = star_spec(N, pt, **kw)
        chunk = prior_chunk(n, seed=123, s_lognormal=sl)
        orc = OracleHelper.from_spec(spec)
        ll = orc.batch_marginal_ln_likelihood(chunk)
        ll_truth, kappa = orc.truth_ll(chunk)
        uu = np.random.default_rng(7).uniform(size=n)
        good = rejection_accept(ll, uu)
        lls_p, a, Ainv = orc.posterior_aAinv(chunk[:16])
        A = np.linalg.inv(Ainv)
        out = {k: np.asarray(spec[k]) for k in SPEC_KEYS}
        out["max_K"] = np.asarray(spec["max_K"] if np.isfinite(spec["max_K"]) else 1e300)
        out.update(chunk=chunk, ll=ll, ll_truth=ll_truth, kappa=kappa, uniforms=uu, good=good,
                   post_ll=lls_p, post_a=a, post_A=A)
        np.savez_compressed(os.path.join(here, f"{name}.npz"), **out)
        print(name, "n_good", len(good), "max|ll-truth|/|truth| %.2e" %
              np.max(np.abs(ll - ll_truth) / np.abs(ll_truth)))


if __name__ == "__main__":
    main()
