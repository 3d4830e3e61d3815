"""The reference's docs/examples/Thompson-black-hole.ipynb on the B200 path.

Radial velocities of 2MASS J05215658+4359220 (Thompson et al. 2019, Science 366, 637),
11 TRES + 3 APOGEE epochs as tabulated in the reference's notebook (cells 5-6), the
notebook's priors (cells 13, 29), and `TheJoker.rejection_sample` with the prior drawn on
the GPU.  The published orbit: P = 83.205 +- 0.064 d, K = 44.615 +- 0.123 km/s,
e = 0.00476, f(M) = 0.766 +- 0.006 Msun.

    python examples/thompson_black_hole.py [log2_prior_samples]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import thejoker_b200 as tj  # noqa: E402
from thejoker_b200 import units as u  # noqa: E402
from thejoker_b200.prior import LogNormal, Normal  # noqa: E402

TRES = np.array([
    [8006.97517, 0.000, 0.075], [8023.98151, -43.313, 0.075], [8039.89955, -27.963, 0.045],
    [8051.98423, 10.928, 0.118], [8070.99556, 43.782, 0.075], [8099.80651, -30.033, 0.054],
    [8106.91698, -42.872, 0.135], [8112.81800, -44.863, 0.088], [8123.79627, -25.810, 0.115],
    [8136.59960, 15.691, 0.146], [8143.78352, 34.281, 0.087]])
APOGEE = np.array([[6204.95544, -37.417, 0.011], [6229.92499, 34.846, 0.010],
                   [6233.87715, 42.567, 0.010]])
PUBLISHED = dict(P=83.205, P_err=0.064, K=44.615, K_err=0.123, e=0.00476, fM=0.766, fM_err=0.00637)
G_MSUN_DAY_KMS = 2 * np.pi * 1.32712440018e11 / 86400.0  # 2 pi G Msun / (1 day) [km^3 s^-3]


def rvdata(tbl):
    # HJD - 2450000 in the table; BMJD = JD - 2400000.5
    return tj.RVData(tbl[:, 0] + 2450000.0 - 2400000.5, tbl[:, 1] * u.km / u.s,
                     tbl[:, 2] * u.km / u.s)


def mass_function(P_day, K_kms, e):
    """f(M) = P K^3 (1 - e^2)^(3/2) / (2 pi G)  [Msun]"""
    return P_day * K_kms**3 * (1 - e**2) ** 1.5 / G_MSUN_DAY_KMS


def run(log2_n=24, devices=(0,), seed=42):
    n = 1 << log2_n
    tres, apogee = rvdata(TRES), rvdata(APOGEE)
    out = {}
    # notebook cell 13-16: TRES alone, periods 16-128 d, extra jitter s ~ LogNormal(-2, 1)
    prior = tj.JokerPrior.default(P_min=16 * u.day, P_max=128 * u.day, sigma_K0=30 * u.km / u.s,
                                  P0=1 * u.year, sigma_v=25 * u.km / u.s,
                                  s=LogNormal("s", -2.0, 1.0, u.km / u.s))
    joker = tj.TheJoker(prior, rng=np.random.default_rng(seed), devices=list(devices))
    t0 = time.perf_counter()
    samples = joker.rejection_sample(tres, n, max_posterior_samples=256).wrap_K()
    out["tres"] = dict(seconds=time.perf_counter() - t0, samples=samples, stats=dict(joker.last_stats))
    # cell 25-30: APOGEE + TRES with a velocity offset between the surveys, P in 75-90 d
    prior_joint = tj.JokerPrior.default(
        P_min=75 * u.day, P_max=90 * u.day, sigma_K0=30 * u.km / u.s, P0=1 * u.year,
        sigma_v=25 * u.km / u.s, v0_offsets=[Normal("dv0_1", 0.0, 5.0, u.km / u.s)],
        s=LogNormal("s", -2.0, 1.0, u.km / u.s))
    joker = tj.TheJoker(prior_joint, rng=np.random.default_rng(seed), devices=list(devices))
    t0 = time.perf_counter()
    samples = joker.rejection_sample([apogee, tres], n, max_posterior_samples=256).wrap_K()
    out["joint"] = dict(seconds=time.perf_counter() - t0, samples=samples,
                        stats=dict(joker.last_stats))
    return out


def summarize(samples):
    P, K, e = samples["P"].to_value(u.day), samples["K"].to_value(u.km / u.s), samples["e"].value
    fM = mass_function(P, K, e)
    return dict(n=len(P), P=(float(np.mean(P)), float(np.std(P))),
                K=(float(np.mean(K)), float(np.std(K))), e_max=float(np.max(e)),
                fM=(float(np.mean(fM)), float(np.std(fM))))


if __name__ == "__main__":
    res = run(int(sys.argv[1]) if len(sys.argv) > 1 else 24)
    for name, r in res.items():
        s = summarize(r["samples"])
        print(f"{name:6s} {r['seconds']:.3f} s  accepted {r['stats']['n_accepted']:>7d} -> kept {s['n']:3d}"
              f"   P = {s['P'][0]:.3f} +- {s['P'][1]:.3f} d   K = {s['K'][0]:.3f} +- {s['K'][1]:.3f} km/s"
              f"   e < {s['e_max']:.4f}   f(M) = {s['fM'][0]:.4f} +- {s['fM'][1]:.4f} Msun")
    print("published: P = 83.205 +- 0.064 d, K = 44.615 +- 0.123 km/s, e = 0.00476, "
          "f(M) = 0.766 +- 0.006 Msun")
