"""tests/golden/ref_api_signatures.json: the public call signatures of the reference's
Python surface for this path, read from the reference's source files with `ast` (build
container only; nothing is imported or copied).

    python tests/golden/make_ref_api_golden.py
"""
import ast
import json
import os

REF = "/root/reference/thejoker"
TARGETS = {
    "thejoker.py": ["TheJoker"],
    "prior.py": ["JokerPrior"],
    "data.py": ["RVData"],
    "samples.py": ["JokerSamples"],
}
FUNCTIONS = {
    "utils.py": ["batch_tasks", "read_batch", "read_batch_slice", "read_batch_idx",
                 "read_random_batch"],
    "likelihood_helpers.py": ["get_constant_term_design_matrix", "get_trend_design_matrix",
                              "ln_normal"],
    "data_helpers.py": ["validate_prepare_data"],
    "prior_helpers.py": ["get_nonlinear_equiv_units", "get_linear_equiv_units",
                         "validate_poly_trend", "validate_n_offsets"],
}


def signature(fn):
    a = fn.args
    names = [x.arg for x in a.posonlyargs + a.args]
    defaults = [None] * (len(names) - len(a.defaults)) + [ast.unparse(d) for d in a.defaults]
    out = [{"name": n, "default": d} for n, d in zip(names, defaults)]
    if a.vararg:
        out.append({"name": "*" + a.vararg.arg, "default": None})
    out += [{"name": k.arg, "default": None if d is None else ast.unparse(d), "kwonly": True}
            for k, d in zip(a.kwonlyargs, a.kw_defaults)]
    if a.kwarg:
        out.append({"name": "**" + a.kwarg.arg, "default": None})
    return out


def main():
    rec = {"classes": {}, "functions": {}}
    for fname, classes in TARGETS.items():
        tree = ast.parse(open(os.path.join(REF, fname)).read())
        for node in tree.body:
            if isinstance(node, ast.ClassDef) and node.name in classes:
                methods = {}
                for item in node.body:
                    if isinstance(item, ast.FunctionDef) and (not item.name.startswith("_")
                                                              or item.name == "__init__"):
                        deco = [ast.unparse(d) for d in item.decorator_list]
                        methods[item.name] = {"args": signature(item), "decorators": deco,
                                              "line": item.lineno}
                rec["classes"][node.name] = {"file": fname, "methods": methods}
    for fname, funcs in FUNCTIONS.items():
        tree = ast.parse(open(os.path.join(REF, fname)).read())
        for node in tree.body:
            if isinstance(node, ast.FunctionDef) and node.name in funcs:
                rec["functions"][node.name] = {"file": fname, "args": signature(node),
                                               "line": node.lineno}
    # the Cython operator: cpdef methods and public attributes of CJokerHelper
    import re

    pyx = open(os.path.join(REF, "src", "fast_likelihood.pyx")).read()
    helper = {"methods": {}, "public_attributes": []}
    for m in re.finditer(r"cpdef\s+(\w+)\(([^)]*)\)", pyx, re.S):
        flat = re.sub(r"\[[^\]]*\]", "", m.group(2).replace("\n", " "))  # drop memoryview types
        args = [a.split("=")[0].split()[-1] for a in flat.split(",")]
        helper["methods"][m.group(1)] = args
    m = re.search(r"def __init__\(([^)]*)\)", pyx, re.S)
    helper["methods"]["__init__"] = [a.split()[-1]
                                     for a in re.sub(r"\[[^\]]*\]", "", m.group(1)).split(",")]
    helper["public_attributes"] = sorted(set(re.findall(r"public\s+[\w\[\]:, ]+?\s(\w+)\s*$",
                                                        pyx, re.M)))
    helper["module_names"] = sorted(set(re.findall(r"^(_nonlinear_\w+)\s*=", pyx, re.M)))
    rec["CJokerHelper"] = helper
    print("CJokerHelper", helper)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_api_signatures.json")
    with open(path, "w") as f:
        json.dump(rec, f, indent=1, sort_keys=True)
    for c, v in rec["classes"].items():
        print(c, sorted(v["methods"]))
    print(sorted(rec["functions"]))


if __name__ == "__main__":
    main()
