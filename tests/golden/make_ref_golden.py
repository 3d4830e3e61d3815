"""Generates tests/golden/ref_*.npz from the REFERENCE'S OWN compiled Cython operator.

Source of the numbers: thejoker/src/fast_likelihood.pyx, translated by Cython and compiled
unmodified from /root/reference by oracle/ref_build/build_ref.py, driven through its real
CJokerHelper.__init__ and its public methods (oracle/ref_cython.py).  The one function
that is not the reference's is twobody's c_rv_from_elements (third party, absent), which
the oracle's restatement supplies -- so these vectors pin everything downstream of the
Kepler solve (jitter handling, design matrix, A/Ainv, b/B/Binv, LAPACK path, ll, the
posterior draw and its rng consumption) to the reference binary, and the Kepler solve to
the restated published algorithm.

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


def main():
    here = os.path.dirname(os.path.abspath(__file__))
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
