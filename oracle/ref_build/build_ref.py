"""Compile the reference's own Cython operator, thejoker/src/fast_likelihood.pyx, from
where it lies under /root/reference into oracle/_ref/ (git-ignored).  TEST INFRASTRUCTURE.

What is the reference's and what is not:
  * fast_likelihood.pyx is translated by Cython and compiled unmodified: get_ivar,
    CJokerHelper.__init__, make_AAinv, make_bBBinv, likelihood_worker,
    batch_marginal_ln_likelihood, batch_get_posterior_samples, test_likelihood_worker,
    with scipy's cython_lapack exactly as upstream.
  * twobody (third party, absent) contributes one C function, c_rv_from_elements; it is
    supplied by the oracle's restatement (twobody_shim.c -> joker_oracle.c).
  * astropy / pymc / pytensor are absent; the module-level imports of the pyx are
    satisfied by the duck-typed stand-ins in oracle/ref_build/shim/ (oracle/ref_cython.py
    installs them in a private interpreter only).

No reference source is copied into the repository: the generated C lives in
oracle/_ref/build/ during the build and is deleted afterwards.

    python oracle/ref_build/build_ref.py          # -> oracle/_ref/fast_likelihood.<abi>.so
"""
import os
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE = os.path.dirname(HERE)
REF_PYX = "/root/reference/thejoker/src/fast_likelihood.pyx"
OUT_DIR = os.path.join(ORACLE, "_ref")


def ext_path():
    return os.path.join(OUT_DIR, "fast_likelihood" + sysconfig.get_config_var("EXT_SUFFIX"))


def build(force=False, verbose=False):
    """Returns the path of the built extension, or None when /root/reference is absent
    (the GPU box: only a prebuilt oracle/_ref travels there)."""
    out = ext_path()
    if not os.path.exists(REF_PYX):
        return out if os.path.exists(out) else None
    srcs = [REF_PYX, os.path.join(HERE, "twobody_shim.c"), os.path.join(ORACLE, "joker_oracle.c"),
            os.path.join(HERE, "src", "twobody.h"), os.path.abspath(__file__)]
    if (not force and os.path.exists(out)
            and all(os.path.getmtime(out) >= os.path.getmtime(s) for s in srcs)):
        return out
    import numpy as np

    bdir = os.path.join(OUT_DIR, "build")
    os.makedirs(bdir, exist_ok=True)
    c_file = os.path.join(bdir, "fast_likelihood.c")
    run = lambda cmd: subprocess.run(cmd, check=True, cwd=bdir,
                                     stdout=None if verbose else subprocess.DEVNULL,
                                     stderr=None if verbose else subprocess.PIPE)
    try:
        run([sys.executable, "-m", "cython", "-3", REF_PYX, "-o", c_file])
        cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        run([cc, "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-w",
             "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION",
             "-I" + sysconfig.get_paths()["include"], "-I" + np.get_include(), "-I" + HERE,
             c_file, os.path.join(HERE, "twobody_shim.c"), os.path.join(ORACLE, "joker_oracle.c"),
             "-o", out, "-lm"])
    finally:
        shutil.rmtree(bdir, ignore_errors=True)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
