"""The Python surface of the hot path keeps the reference's names and argument order.

tests/golden/ref_api_signatures.json holds the public signatures of the reference's
TheJoker / JokerPrior / RVData / JokerSamples and of the host helpers on the path, read
from its sources with `ast` (tests/golden/make_ref_api_golden.py).  For every one that is
in scope the product must offer the same callable with the reference's parameters as a
prefix, in order, with the same defaults (the product may append keyword arguments such
as ``devices=``).  What is out of scope is listed here with the reason.
"""
import inspect
import json
import os

import pytest

import thejoker_b200 as tj
from thejoker_b200 import cache, data_helpers, likelihood_helpers, sharding
from thejoker_b200 import prior as prior_helpers  # the reference keeps these in prior_helpers.py

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "ref_api_signatures.json")) as f:
    REF = json.load(f)

CLASSES = {"TheJoker": tj.TheJoker, "JokerPrior": tj.JokerPrior, "RVData": tj.RVData,
           "JokerSamples": tj.JokerSamples}
FUNCTIONS = {"batch_tasks": sharding.batch_tasks,
             "read_batch": cache.read_batch, "read_batch_slice": cache.read_batch_slice,
             "read_batch_idx": cache.read_batch_idx, "read_random_batch": cache.read_random_batch,
             "get_constant_term_design_matrix": likelihood_helpers.get_constant_term_design_matrix,
             "get_trend_design_matrix": likelihood_helpers.get_trend_design_matrix,
             "ln_normal": likelihood_helpers.ln_normal,
             "validate_prepare_data": data_helpers.validate_prepare_data,
             "get_nonlinear_equiv_units": prior_helpers.get_nonlinear_equiv_units,
             "get_linear_equiv_units": prior_helpers.get_linear_equiv_units,
             "validate_poly_trend": prior_helpers.validate_poly_trend,
             "validate_n_offsets": prior_helpers.validate_n_offsets}

# (class, method) -> why it is not part of this build (SURVEY.md section 8 / DESIGN.md section 2)
OUT_OF_SCOPE = {
    ("TheJoker", "setup_mcmc"): "pymc MCMC hand-off (SURVEY section 2 rows 14-19)",
    ("JokerPrior", "__init__"): "takes pymc random variables; the product's prior carries "
                                "plain distribution objects (pymc absent) -- default() and "
                                "sample() keep the reference's signatures",
    ("RVData", "plot"): "matplotlib plotting",
    ("RVData", "guess_from_table"): "astropy Table column guessing",
    ("RVData", "from_timeseries"): "astropy TimeSeries I/O",
    ("RVData", "to_timeseries"): "astropy TimeSeries I/O",
    ("RVData", "t"): "astropy Time object; the path uses _t_bmjd / t_ref",
    ("JokerSamples", "from_inference_data"): "arviz / pymc trace conversion",
    ("JokerSamples", "get_orbit"): "twobody KeplerOrbit objects",
    ("JokerSamples", "orbits"): "twobody KeplerOrbit objects",
    ("JokerSamples", "get_t0"): "astropy Time arithmetic on top of the samples",
    ("JokerSamples", "get_time_with_phase"): "astropy Time arithmetic on top of the samples",
}


def _params(obj):
    fn = obj.fget if isinstance(obj, property) else obj
    fn = getattr(fn, "__func__", fn)
    return list(inspect.signature(fn).parameters.values())


def _check_prefix(where, ref_args, params):
    ref_pos = [a for a in ref_args if not a["name"].startswith("*") and not a.get("kwonly")]
    names = [p.name for p in params]
    want = [a["name"] for a in ref_pos]
    assert names[:len(want)] == want, f"{where}: {names} does not start with {want}"
    for a, p in zip(ref_pos, params):
        if a["default"] is None:
            continue
        assert p.default is not inspect.Parameter.empty, f"{where}: {p.name} lost its default"
        if repr(p.default) == a["default"] or str(p.default) == a["default"]:
            continue
        # an expression such as `1 * u.year`: evaluate it in the product's unit system
        from thejoker_b200 import units as u

        want_value = eval(a["default"], {"u": u, "np": __import__("numpy")})
        assert u.to_value(p.default, want_value.unit, want_value.unit) == want_value.value, \
            f"{where}: default of {p.name} is {p.default!r}, reference has {a['default']}"


CASES = [(c, m) for c, v in sorted(REF["classes"].items()) for m in sorted(v["methods"])]


@pytest.mark.parametrize("cls,method", CASES)
def test_method_signature(cls, method):
    if (cls, method) in OUT_OF_SCOPE:
        pytest.skip(OUT_OF_SCOPE[(cls, method)])
    ref = REF["classes"][cls]["methods"][method]
    obj = inspect.getattr_static(CLASSES[cls], method)
    is_prop = any(d in ("property", "cached_property") for d in ref["decorators"])
    if is_prop:
        assert isinstance(obj, property), f"{cls}.{method} is a property in the reference"
        return
    _check_prefix(f"{cls}.{method}", ref["args"], _params(obj))


@pytest.mark.parametrize("name", sorted(REF["functions"]))
def test_function_signature(name):
    _check_prefix(name, REF["functions"][name]["args"], _params(FUNCTIONS[name]))


def test_every_out_of_scope_entry_is_real():
    for cls, method in OUT_OF_SCOPE:
        assert method in REF["classes"][cls]["methods"], (cls, method)


def test_cjokerhelper_surface():
    """The operator boundary itself: the cpdef methods of the reference's CJokerHelper with
    their argument names, its public attributes, and the two module-level names
    samples.py imports from the extension (pyx:41-45)."""
    from thejoker_b200 import helper as helper_mod

    ref = REF["CJokerHelper"]
    for name, args in ref["methods"].items():
        fn = getattr(tj.CJokerHelper, name)
        names = [p.name for p in inspect.signature(fn).parameters.values()]
        assert names[:len(args)] == args, (name, names, args)
    # B / Binv (N x N) are never formed on the GPU path; the oracle has them
    never_formed = {"B", "Binv"}
    src = inspect.getsource(tj.CJokerHelper)
    for attr in ref["public_attributes"]:
        if attr in never_formed:
            continue
        assert f"self.{attr}" in src, attr
    for name in ref["module_names"]:
        assert hasattr(helper_mod, name), name
