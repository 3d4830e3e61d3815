"""Driver for the reference's own compiled Cython operator (oracle/_ref/, built by
oracle/ref_build/build_ref.py from /root/reference/thejoker/src/fast_likelihood.pyx).

TEST INFRASTRUCTURE ONLY.  Used to pin the oracle: tests/golden/make_ref_golden.py runs
it in this container to produce tests/golden/ref_*.npz, tests compare the C restatement
(oracle/joker_oracle.c) and the CUDA path against those vectors, and -- when the built
extension travelled to the GPU box -- bench.py --impl reference times it.

`RefCythonHelper(spec, poly_trend, n_offsets)` constructs the reference's CJokerHelper
through its real __init__ from duck-typed (data, prior) objects that carry the same plain
arrays an `OracleHelper` takes, so both see identical inputs.  Only twobody's
c_rv_from_elements (absent third party) is not the reference's code; see build_ref.py.
"""
from __future__ import annotations

import contextlib
import importlib.util
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIM = os.path.join(_HERE, "ref_build", "shim")
_SHIM_NAMES = ("astropy", "astropy.units", "thejoker", "thejoker.units", "thejoker.utils",
               "thejoker.distributions", "thejoker.logging", "thejoker.samples",
               "thejoker.likelihood_helpers", "thejoker.src", "thejoker.src.fast_likelihood")
REF_LIKELIHOOD_HELPERS = "/root/reference/thejoker/likelihood_helpers.py"
_module = None
_lh_module = None
_shim_modules = {}


def ext_path():
    sys.path.insert(0, os.path.join(_HERE, "ref_build"))
    try:
        import build_ref
    finally:
        sys.path.pop(0)
    return build_ref.build()


def available() -> bool:
    """True when the compiled reference operator can be loaded (built here if the
    reference sources are present; a failed build only means "not available")."""
    try:
        p = ext_path()
    except Exception:
        return False
    return p is not None and os.path.exists(p)


@contextlib.contextmanager
def _shims():
    """Temporarily expose the stand-in packages (and the loaded extension) in sys.modules;
    whatever was there before is restored, so the rest of the process never sees them."""
    saved = {n: sys.modules.get(n) for n in _SHIM_NAMES}
    sys.path.insert(0, _SHIM)
    try:
        for n in _SHIM_NAMES:
            sys.modules.pop(n, None)
        sys.modules.update(_shim_modules)
        yield
    finally:
        for n in _SHIM_NAMES:
            if n in sys.modules:
                _shim_modules[n] = sys.modules.pop(n)
            if saved[n] is not None:
                sys.modules[n] = saved[n]
        sys.path.remove(_SHIM)


def load():
    """The compiled reference module thejoker.src.fast_likelihood."""
    global _module
    if _module is None:
        path = ext_path()
        if path is None or not os.path.exists(path):
            raise RuntimeError("oracle/_ref is not built (needs /root/reference; run "
                               "python oracle/ref_build/build_ref.py in the build container)")
        with _shims():
            import thejoker.src  # noqa: F401  (the stand-in parent packages)

            name = "thejoker.src.fast_likelihood"
            spec = importlib.util.spec_from_file_location(name, path)
            mod = importlib.util.module_from_spec(spec)
            sys.modules[name] = mod
            spec.loader.exec_module(mod)
            _module = mod
    return _module


def load_likelihood_helpers():
    """The reference's thejoker/likelihood_helpers.py, executed from where it lies (pure
    Python: numpy + its package logger; build container only)."""
    global _lh_module
    if _lh_module is None:
        if not os.path.exists(REF_LIKELIHOOD_HELPERS):
            raise RuntimeError("needs /root/reference (build container only)")
        with _shims():
            import thejoker  # noqa: F401

            name = "thejoker.likelihood_helpers"
            spec = importlib.util.spec_from_file_location(name, REF_LIKELIHOOD_HELPERS)
            mod = importlib.util.module_from_spec(spec)
            sys.modules[name] = mod
            spec.loader.exec_module(mod)
            _lh_module = mod
    return _lh_module


class _Named:
    def __init__(self, name):
        self.name = name


class _Dist:
    """What the pyx reads from prior.model[name] / prior.pars[name]."""

    def __init__(self, unit_attr, unit, mu=0.0, sigma=1.0, print_name="Normal"):
        setattr(self, unit_attr, unit)
        self.mu, self.sigma = mu, sigma
        self.owner = _Named(None)
        self.owner.op = _Named(None)
        self.owner.op._print_name = (print_name, "\\operatorname{%s}" % print_name)


class RefCythonHelper:
    """The reference's CJokerHelper, constructed from plain arrays.

    spec : dict with t, rv, ivar, t0, trend_M, mu, Lambda (order K, v0, offsets, v1, ...;
           pyx:204-252), K_prior_kind (0 FixedCompanionMass / 1 Normal), sigma_K0, P0, max_K.
    """

    def __init__(self, spec, poly_trend, n_offsets=0):
        mod = load()
        with _shims():
            import astropy.units as u
            from thejoker.units import UNIT_ATTR_NAME as UA

            rv_unit = u.Unit({"rv": 1})
            t = np.ascontiguousarray(spec["t"], dtype="f8")
            L = 1 + poly_trend + n_offsets
            mu, Lam = np.asarray(spec["mu"], "f8"), np.asarray(spec["Lambda"], "f8")
            trend_M = np.asarray(spec["trend_M"], "f8")
            if trend_M.ndim != 2:
                trend_M = trend_M.reshape(len(t), L - 1)
            trend_M = np.ascontiguousarray(trend_M)

            data = _Named("data")
            data.rv = u.Quantity(np.asarray(spec["rv"], "f8"), rv_unit)
            data.ivar = u.Quantity(np.asarray(spec["ivar"], "f8"), 1 / rv_unit ** 2)
            data._t_bmjd = t
            data._t_ref_bmjd = float(spec["t0"])
            data.t_ref = None
            data_cls = type("ShimData", (), {"__len__": lambda s: len(t)})
            d = data_cls()
            d.__dict__.update(data.__dict__)

            prior = _Named("prior")
            prior.poly_trend, prior.n_offsets = poly_trend, n_offsets
            prior._v_trend_names = ["v%d" % i for i in range(poly_trend)]
            prior.v0_offsets = [_Named("dv0_%d" % (i + 1)) for i in range(n_offsets)]
            lin_names = ["K"] + prior._v_trend_names
            prior._linear_equiv_units = {n: None for n in lin_names}
            prior.par_names = (["P", "e", "omega", "M0", "s"] + lin_names
                               + [o.name for o in prior.v0_offsets])
            model = {}
            fixed = int(spec.get("K_prior_kind", 0)) == 0
            K = _Dist(UA, rv_unit, mu[0], np.sqrt(Lam[0]) if not fixed else 1.0,
                      "FixedCompanionMass" if fixed else "Normal")
            if fixed:
                K._sigma_K0 = u.Quantity(float(spec["sigma_K0"]), rv_unit)
                K._P0 = u.Quantity(float(spec["P0"]), u.day)
                K._max_K = u.Quantity(float(spec["max_K"]), rv_unit)
            model["K"] = K
            model["v0"] = _Dist(UA, rv_unit, mu[1], np.sqrt(Lam[1])) if poly_trend >= 1 else None
            for i in range(n_offsets):
                model["dv0_%d" % (i + 1)] = _Dist(UA, rv_unit, mu[2 + i], np.sqrt(Lam[2 + i]))
            for i in range(1, poly_trend):
                j = 1 + n_offsets + i
                model["v%d" % i] = _Dist(UA, rv_unit / u.day ** i, mu[j], np.sqrt(Lam[j]))
            prior.model = model
            prior.pars = {"K": K, "P": _Dist(UA, u.day)}
            self.helper = mod.CJokerHelper(d, prior, trend_M)
        self.n_times, self.n_linear = len(t), L

    def batch_marginal_ln_likelihood(self, chunk):
        return np.asarray(self.helper.batch_marginal_ln_likelihood(
            np.ascontiguousarray(chunk, dtype="f8")))

    def test_likelihood_worker(self, row):
        """ll plus the (a, A, Ainv, b, B, Binv) the reference's worker leaves behind."""
        h = self.helper
        ll = h.test_likelihood_worker(np.ascontiguousarray(row, dtype="f8"))
        return ll, {k: np.array(getattr(h, k)) for k in ("a", "A", "Ainv", "b", "B", "Binv")}

    def batch_get_posterior_samples(self, chunk, n_linear_samples_per, rng):
        return self.helper.batch_get_posterior_samples(
            np.ascontiguousarray(chunk, dtype="f8"), int(n_linear_samples_per), rng)

    # -- the reference's in-memory drivers (thejoker/likelihood_helpers.py:91-229), run on
    #    this helper: rng.uniform accept, truncation, posterior draws, iteration schedule
    def rejection_sample_inmem(self, chunk, rng, **kw):
        lh = load_likelihood_helpers()
        with _shims():
            return lh.rejection_sample_inmem(self.helper, np.ascontiguousarray(chunk, "f8"), rng,
                                             **kw)

    def iterative_rejection_inmem(self, chunk, rng, n_requested_samples, **kw):
        lh = load_likelihood_helpers()
        with _shims():
            return lh.iterative_rejection_inmem(self.helper, np.ascontiguousarray(chunk, "f8"),
                                                rng, n_requested_samples, **kw)


def reference_function(rel_path, name, namespace=None):
    """One top-level function of a reference module, compiled from the file where it lies
    (ast-extracted, so that module-level imports of absent packages are not executed).
    Build container only."""
    import ast

    path = os.path.join("/root/reference", rel_path)
    tree = ast.parse(open(path).read(), filename=path)
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == name:
            ns = {"np": np}
            ns.update(namespace or {})
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
            return ns[name]
    raise KeyError(name)
