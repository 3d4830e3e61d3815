"""Generates tests/golden/ref_*.npz from the REFERENCE'S OWN compiled Cython operator.

Source of the numbers: thejoker/src/fast_likelihood.pyx, translated by Cython and compiled
unmodified from /root/reference by oracle/ref_build/build_ref.py, driven through its real
CJokerHelper.__init__ and its public methods (oracle/ref_cython.py).  The one function
that is not the reference's is twobody's c_rv_from_elements (third party, absent), which
the oracle's restatement supplies -- so these vectors pin everything downstream of the
Kepler solve (jitter handling, design matrix, A/Ainv, b/B/Binv, LAPACK path, ll, the
posterior draw and its rng consumption) to the reference binary, and the Kepler solve to
the restated published algorithm.

ref_rejection_*.npz additionally run the reference's own in-memory drivers,
thejoker/likelihood_helpers.py:91-229 (rejection_sample_inmem, iterative_rejection_inmem),
executed from where the file lies on that compiled helper: rng.uniform accept against
lls.max(), truncation, the posterior draws that follow on the same Generator, and the
batch-growth schedule of the iterative sampler.

Can only run in the build container (needs /root/reference):
    python tests/golden/make_ref_golden.py
Inputs are the same seeded stars / prior chunks as make_golden.py.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import prior_chunk, star_spec  # noqa: E402
from make_golden import CASES, SPEC_KEYS  # noqa: E402
from oracle.ref_cython import RefCythonHelper  # noqa: E402

N_LL = 512       # rows of the prior chunk through batch_marginal_ln_likelihood
N_POST = 8       # rows through test_likelihood_worker / batch_get_posterior_samples
N_DRAW = 3       # linear draws per row


REJECTION_CASES = ("n3_normalK", "n64_l2_flat", "n16_l2", "n20_l3_offsets")
N_REJ = 4096     # prior rows handed to the reference's drivers
SEED_REJ, SEED_ITER = 21, 5


def rejection_cases(here):
    for name in REJECTION_CASES:
        N, pt, sl, kw = CASES[name]
        spec, _, _ = star_spec(N, pt, **kw)
        chunk = prior_chunk(N_REJ, seed=123, s_lognormal=sl)
        ln_prior = np.random.default_rng(1).normal(size=N_REJ)
        ref = RefCythonHelper(spec, poly_trend=spec["n_poly"], n_offsets=spec["n_offsets"])
        out = {k: np.asarray(spec[k]) for k in SPEC_KEYS}
        out["max_K"] = np.asarray(spec["max_K"] if np.isfinite(spec["max_K"]) else 1e300)
        out["n_poly"], out["n_offsets"] = np.asarray(spec["n_poly"]), np.asarray(spec["n_offsets"])
        out.update(chunk=chunk, ln_prior=ln_prior, seed_rej=SEED_REJ, seed_iter=SEED_ITER)
        # (1) plain rejection, every accepted row, 2 linear draws per row
        smp, lls = ref.rejection_sample_inmem(chunk, np.random.default_rng(SEED_REJ),
                                              ln_prior=ln_prior, n_linear_samples=2,
                                              return_all_logprobs=True)
        out.update(rej_raw=np.array(smp["raw"]), rej_lls=lls, rej_ln_prior=smp["ln_prior"],
                   rej_ln_likelihood=smp["ln_likelihood"])
        # (2) truncated at max_posterior_samples = 3, 1 draw per row
        smp = ref.rejection_sample_inmem(chunk, np.random.default_rng(SEED_REJ),
                                         max_posterior_samples=3)
        out["rej3_raw"] = np.array(smp["raw"])
        # (3) iterative: 4 requested, first batch 256 rows
        smp = ref.iterative_rejection_inmem(chunk, np.random.default_rng(SEED_ITER), 4,
                                            ln_prior=ln_prior, init_batch_size=256)
        out.update(iter_raw=np.array(smp["raw"]), iter_ln_prior=smp["ln_prior"],
                   iter_ln_likelihood=smp["ln_likelihood"], iter_n_requested=4,
                   iter_init_batch_size=256)
        np.savez_compressed(os.path.join(here, f"ref_rejection_{name}.npz"), **out)
        print("rejection", name, "accepted", len(out["rej_ln_prior"]), "truncated",
              len(out["rej3_raw"]), "iterative", len(out["iter_raw"]))


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    rejection_cases(here)
    for name, (N, pt, sl, kw) in CASES.items():
        spec, _, _ = star_spec(N, pt, **kw)
        chunk = prior_chunk(N_LL, seed=123, s_lognormal=sl)
        ref = RefCythonHelper(spec, poly_trend=spec["n_poly"], n_offsets=spec["n_offsets"])
        out = {k: np.asarray(spec[k]) for k in SPEC_KEYS}
        out["max_K"] = np.asarray(spec["max_K"] if np.isfinite(spec["max_K"]) else 1e300)
        out["n_poly"], out["n_offsets"] = np.asarray(spec["n_poly"]), np.asarray(spec["n_offsets"])
        out["chunk"] = chunk
        out["ref_ll"] = ref.batch_marginal_ln_likelihood(chunk)
        wk = {k: [] for k in ("ll", "a", "A", "Ainv", "b", "B", "Binv")}
        for row in chunk[:N_POST]:
            ll, mats = ref.test_likelihood_worker(row)
            wk["ll"].append(ll)
            for k, v in mats.items():
                wk[k].append(v)
        for k, v in wk.items():
            out["ref_worker_" + k] = np.array(v)
        samples, lls = ref.batch_get_posterior_samples(chunk[:N_POST], N_DRAW,
                                                       np.random.default_rng(11))
        out["ref_samples"], out["ref_samples_ll"] = np.array(samples), np.array(lls)
        np.savez_compressed(os.path.join(here, f"ref_{name}.npz"), **out)
        print(name, "ll[:2]", out["ref_ll"][:2], "worker ll[:2]", out["ref_worker_ll"][:2])


if __name__ == "__main__":
    main()
