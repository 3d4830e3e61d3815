"""Generates tests/golden/ref_truth.npz: the quad-precision value (oracle/joker_truth.c,
__float128, residual form) of every ll stored in the reference-minted fixtures ref_n*.npz
and ref_rejection_*.npz, with the reference's jitter semantics (s ignored, pyx:458).

The reference's own double-precision formula loses digits where chi2 = d^T B^-1 d cancels
(flat, high-S/N data: it does not centre y); this file is what lets the GPU tests gate the
CUDA path at 1e-10 of the exact value on those fixtures and *measure* how far the
reference itself is from it, instead of loosening the gate against the reference.

    python tests/golden/make_ref_truth.py        (needs only the oracle, not /root/reference)
"""
import glob
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from make_golden import SPEC_KEYS  # noqa: E402
from oracle.oracle import OracleHelper  # noqa: E402


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    out = {}
    for path in sorted(glob.glob(os.path.join(here, "ref_n*.npz"))
                       + glob.glob(os.path.join(here, "ref_rejection_*.npz"))):
        name = os.path.basename(path)[:-4]
        z = np.load(path)
        spec = {k: (z[k] if z[k].ndim else z[k].item()) for k in SPEC_KEYS}
        spec["jitter_mode"] = 0  # the reference ignores the jitter column
        orc = OracleHelper.from_spec(spec)
        truth, kappa = orc.truth_ll(np.ascontiguousarray(z["chunk"]))
        ref = z["ref_ll"] if "ref_ll" in z else z["rej_lls"]
        out[name] = truth
        out[name + "_kappa"] = kappa
        r = np.abs(ref - truth) / np.abs(truth)
        print(f"{name}: n={len(truth)}  reference vs truth max {r.max():.2e}  "
              f"(> 1e-10: {(r > 1e-10).sum()})  kappa max {kappa.max():.2e}")
    np.savez_compressed(os.path.join(here, "ref_truth.npz"), **out)


if __name__ == "__main__":
    main()
