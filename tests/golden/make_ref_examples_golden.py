"""tests/golden/ref_examples.npz: radial velocities computed by the third-party `twobody`
package itself, as stored in the reference repository.

docs/examples/make-data.ipynb (cells 2-13) seeds numpy's default_rng(123), draws the true
orbital elements, evaluates `twobody.KeplerOrbit.radial_velocity(t)` and writes the result
-- WITHOUT adding noise; `rv_err` is only stored alongside -- to docs/examples/data.ecsv,
data-triple.ecsv and data-survey{1,2}.ecsv.  The truths are pickled astropy objects
(unreadable here), but the notebook's draw order is reproduced below from the same seed,
and the stored `rv_err` columns confirm that the stream is aligned (they are the
10**uniform draws that follow the epochs in the same stream).

So these files are known-answer vectors for the one function of the hot path that lives
outside /root/reference: twobody's c_rv_from_elements (M = 2 pi (t - t0)/P - M0, Kepler
solve, rv = K (cos(f + omega) + e cos omega) + v0).  Times are stored as float64 TCB
Julian dates (resolution 4.7e-10 d), which limits the comparison to a few 1e-10 km/s.

Build container only (reads /root/reference/docs/examples):
    python tests/golden/make_ref_examples_golden.py
"""
import os
import re

import numpy as np

EX = "/root/reference/docs/examples"


def read_ecsv(name):
    txt = open(os.path.join(EX, name)).read()
    m = re.search(r"jd1:\s*([-\d.eE+]+),\s*jd2:\s*([-\d.eE+]+)", txt)
    t_ref = (float(m.group(1)), float(m.group(2))) if m else None
    lines = [ln for ln in txt.splitlines() if not ln.startswith("#")]
    assert lines[0].split() == ["bjd", "rv", "rv_err"]
    arr = np.array([[float(x) for x in ln.split()] for ln in lines[1:] if ln.strip()])
    return arr, t_ref


def draw_truth(rnd, e):
    """cell 4 / 8 / 11: t0 offset, P, M0, omega, K, v0 in this order."""
    t0_off = rnd.uniform(0.0, 40)
    return dict(t0_off=t0_off, P=rnd.uniform(40, 80), M0=rnd.uniform(0.0, 2 * np.pi),
                omega=rnd.uniform(0.0, 2 * np.pi), e=e, K=rnd.uniform(5, 15),
                v0=rnd.uniform(-50, 50))


def main():
    rnd = np.random.default_rng(seed=123)  # cell 2
    out = {}

    # cell 4 -> data.ecsv
    tr = draw_truth(rnd, 0.1)
    x = np.concatenate(([0], np.sort(rnd.uniform(0, 3.0, 256))))
    err = 10 ** rnd.uniform(-1, 0.5, size=257)
    arr, t_ref = read_ecsv("data.ecsv")
    assert np.allclose(err, arr[:, 2], rtol=1e-14, atol=0), "rng stream not aligned (data)"
    dt = (arr[:, 0] - t_ref[0]) - t_ref[1]
    assert np.max(np.abs(dt - tr["P"] * x)) < 1.6e-8 * dt.max() + 1e-9  # UTC vs TCB day length
    out.update(single_dt=dt, single_rv=arr[:, 1], single_rv_err=arr[:, 2],
               single_truth=np.array([tr[k] for k in ("P", "e", "omega", "M0", "K", "v0")]))

    # cell 8 -> data-triple.ecsv (two Keplerian components, no t_ref in the file's meta:
    # the first epoch is t0 itself)
    tr1 = draw_truth(rnd, 0.25)
    tr2 = dict(P=10 * rnd.uniform(40, 80), M0=rnd.uniform(0.0, 2 * np.pi),
               omega=rnd.uniform(0.0, 2 * np.pi), e=0.1, K=13.0)
    x = np.concatenate(([0], np.sort(rnd.uniform(0, 5.0, 256))))
    err = 10 ** rnd.uniform(-1, 0.5, size=257)
    arr, _ = read_ecsv("data-triple.ecsv")
    assert np.allclose(err, arr[:, 2], rtol=1e-14, atol=0), "rng stream not aligned (triple)"
    dt = arr[:, 0] - arr[0, 0]
    assert np.max(np.abs(dt - tr1["P"] * x)) < 1.6e-8 * dt.max() + 1e-9
    out.update(triple_dt=dt, triple_rv=arr[:, 1], triple_rv_err=arr[:, 2],
               triple_truth1=np.array([tr1[k] for k in ("P", "e", "omega", "M0", "K", "v0")]),
               triple_truth2=np.array([tr2[k] for k in ("P", "e", "omega", "M0", "K")]))

    # cell 11-13 -> data-survey1.ecsv (10 epochs) + data-survey2.ecsv (6 epochs, +4.8 km/s)
    tr = draw_truth(rnd, 0.13)
    x = np.concatenate(([0], np.sort(rnd.uniform(0, 3.0, 16))))
    err = 10 ** rnd.uniform(-1, 0.5, size=17)
    a1, t_ref = read_ecsv("data-survey1.ecsv")
    a2, _ = read_ecsv("data-survey2.ecsv")
    arr = np.concatenate([a1, a2])
    assert np.allclose(err, arr[:, 2], rtol=1e-14, atol=0), "rng stream not aligned (survey)"
    dt = (arr[:, 0] - t_ref[0]) - t_ref[1]
    assert np.max(np.abs(dt - tr["P"] * x)) < 1.6e-8 * dt.max() + 1e-9
    out.update(survey_dt=dt, survey_rv=arr[:, 1], survey_rv_err=arr[:, 2], survey_n1=len(a1),
               survey_offset=4.8,
               survey_truth=np.array([tr[k] for k in ("P", "e", "omega", "M0", "K", "v0")]))

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_examples.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;",
          {k: v.shape for k, v in out.items() if hasattr(v, "shape") and v.ndim})


if __name__ == "__main__":
    main()
