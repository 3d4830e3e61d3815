"""tests/golden/ref_host_logic.json: outputs of the reference's own host-side helpers on
the path -- utils.py::batch_tasks (22-72) and likelihood_helpers.py::
get_constant_term_design_matrix / get_trend_design_matrix / ln_normal (8-37, 232-233) --
executed from the files where they lie under /root/reference (build container only).

    python tests/golden/make_ref_host_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.ref_cython import load_likelihood_helpers, reference_function  # noqa: E402


class FakeData:
    def __init__(self, t, t_ref):
        self._t_bmjd, self._t_ref_bmjd = np.asarray(t, float), float(t_ref)

    def __len__(self):
        return len(self._t_bmjd)


def design_cases():
    rng = np.random.default_rng(17)
    out = []
    for n, ids_kind, poly in [(5, None, 1), (7, "two", 1), (9, "three", 2), (6, None, 3),
                              (8, "gap", 4), (4, "two", 5)]:
        t = np.sort(rng.uniform(55000.0, 56000.0, n))
        ids = {None: None, "two": rng.integers(0, 2, n), "three": rng.integers(0, 3, n),
               "gap": rng.choice([2, 5, 9], n)}[ids_kind]
        if ids is not None:
            ids[0] = ids.min()  # keep every case's first id present
        out.append((t, float(t.min()), None if ids is None else ids.tolist(), poly))
    return out


class _Q:
    """What the reference's samples_analysis functions read from sample['P']."""

    def __init__(self, value):
        self.value = np.asarray(value, float)

    def to(self, unit):
        return self

    def to_value(self, unit):
        return self.value


class _T:
    def __init__(self, mjd):
        self.mjd = self.jd = np.asarray(mjd, float)
        self.tcb = self


class _Data:
    def __init__(self, t, t_ref):
        self._t, self._t_ref = np.asarray(t, float), float(t_ref)
        self.t = _T(self._t)

    def phase(self, P):
        return ((self._t - self._t_ref) / np.asarray(getattr(P, "value", P), float)) % 1.0


def samples_analysis_cases():
    """samples_analysis.py:35-135 on plain arrays (P in days, t in BMJD)."""
    ns = {"u": type("U", (), {"day": "day"})}
    fns = {n: reference_function("thejoker/samples_analysis.py", n, ns)
           for n in ("is_P_unimodal", "max_phase_gap", "phase_coverage", "periods_spanned")}
    rng = np.random.default_rng(5)
    out = []
    for k in range(8):
        n = int(rng.integers(4, 30))
        t = np.sort(56000.0 + rng.uniform(0, 400, n))
        t_ref = float(t.min())
        P = float(np.exp(rng.uniform(np.log(2), np.log(300))))
        Ps = P * (1 + 10 ** rng.uniform(-6, -1) * rng.normal(size=6))
        data = _Data(t, t_ref)
        one = {"P": _Q(P)}
        out.append({"t": t.tolist(), "t_ref": t_ref, "P": P, "P_samples": Ps.tolist(),
                    "is_P_unimodal": bool(fns["is_P_unimodal"]({"P": _Q(Ps)}, data)),
                    "max_phase_gap": float(fns["max_phase_gap"](one, data)),
                    "phase_coverage": float(fns["phase_coverage"](one, data)),
                    "periods_spanned": float(fns["periods_spanned"](one, data))})
    return out


def main():
    batch_tasks = reference_function("thejoker/utils.py", "batch_tasks")
    lh = load_likelihood_helpers()
    rec = {"batch_tasks": [], "design": [], "ln_normal": []}
    for n_tasks, n_batches, start in [(10, 3, 0), (10, 3, 7), (3, 8, 0), (0, 4, 0), (16, 4, 2),
                                      (17, 16, 0), (1 << 28, 8, 0), (1000, 0, 5), (5, 5, 1),
                                      (268435456, 7, 1024)]:
        tasks = batch_tasks(n_tasks, n_batches, start_idx=start, args=["x"])
        rec["batch_tasks"].append({"n_tasks": n_tasks, "n_batches": n_batches, "start_idx": start,
                                   "tasks": [[list(t[0]), t[1], t[2]] for t in tasks]})
    arr = np.arange(23) * 10
    tasks = batch_tasks(11, 4, arr=arr, start_idx=3)
    rec["batch_tasks_arr"] = {"arr": arr.tolist(), "n_tasks": 11, "n_batches": 4, "start_idx": 3,
                              "tasks": [[t[0].tolist(), t[1]] for t in tasks]}
    for t, t_ref, ids, poly in design_cases():
        data = FakeData(t, t_ref)
        M = lh.get_trend_design_matrix(data, ids, poly)
        C = lh.get_constant_term_design_matrix(data, ids)
        rec["design"].append({"t": t.tolist(), "t_ref": t_ref, "ids": ids, "poly_trend": poly,
                              "trend_M": M.tolist(), "const_M": C.tolist()})
    for x, mu, var in [(0.3, 0.1, 2.0), (-5.0, 1.0, 0.01), (1e3, 0.0, 1e4)]:
        rec["ln_normal"].append({"x": x, "mu": mu, "var": var, "value": float(lh.ln_normal(x, mu, var))})
    rec["samples_analysis"] = samples_analysis_cases()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_host_logic.json")
    with open(path, "w") as f:
        json.dump(rec, f)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
