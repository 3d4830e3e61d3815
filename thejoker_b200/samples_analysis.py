"""Post-accept utilities on a handful of posterior samples (host side; SURVEY.md
section 8 f4).  Same names and semantics as thejoker/samples_analysis.py; periods are
taken in days, times in BMJD."""
from __future__ import annotations

import numpy as np

from . import units as u

__all__ = ["MAP_sample", "is_P_unimodal", "is_P_Kmodal", "max_phase_gap", "phase_coverage",
           "periods_spanned", "phase_coverage_per_period"]


def _P_days(samples):
    return np.atleast_1d(samples["P"].to_value(u.day))


def MAP_sample(samples, return_index=False):
    """The maximum a posteriori sample (samples_analysis.py:12-34)."""
    if "ln_prior" not in samples or "ln_likelihood" not in samples:
        raise ValueError("You must pass in samples that have prior and likelihood information "
                         "stored; use return_logprobs=True when generating the samples.")
    idx = int(np.argmax(samples["ln_prior"].value + samples["ln_likelihood"].value))
    return (samples[idx], idx) if return_index else samples[idx]


def is_P_unimodal(samples, data):
    """True when the period samples span less than one period-mode width
    4 P_min^2 / (2 pi T) (samples_analysis.py:37-57)."""
    P = _P_days(samples)
    T = np.ptp(data._t_bmjd)
    return np.ptp(P) < 4 * P.min() ** 2 / (2 * np.pi * T)


def is_P_Kmodal(samples, data, n_clusters=2, n_iter=50):
    """Experimental (samples_analysis.py:60-97): cluster ln P into n_clusters modes (1-D
    Lloyd iterations from quantile seeds) and test each for unimodality.  Returns
    (all modes unimodal, representative period per mode [day], samples per mode)."""
    lnP = np.log(_P_days(samples))
    centres = np.quantile(lnP, (np.arange(n_clusters) + 0.5) / n_clusters)
    for _ in range(n_iter):
        lab = np.argmin(np.abs(lnP[:, None] - centres[None, :]), axis=1)
        new = np.array([lnP[lab == j].mean() if np.any(lab == j) else centres[j]
                        for j in range(n_clusters)])
        if np.allclose(new, centres):
            break
        centres = new
    unimodal, reps, counts = [], [], []
    has_post = "ln_prior" in samples and "ln_likelihood" in samples
    for j in np.unique(lab):
        sub = samples[np.where(lab == j)[0]]
        unimodal.append(True if len(sub) == 1 else bool(is_P_unimodal(sub, data)))
        rep = MAP_sample(sub) if (has_post and len(sub) > 1) else sub[0]
        reps.append(float(_P_days(rep)[0]))
        counts.append(int(np.sum(lab == j)))
    return all(unimodal), np.array(reps) * u.day, np.array(counts)


def max_phase_gap(sample, data):
    """Largest gap in orbital phase between consecutive observations
    (samples_analysis.py:96-107).  As in the reference the sorted phases are concatenated
    with themselves unshifted, so the gap across phase 1 -> 0 is not counted."""
    phase = np.sort(data.phase(_P_days(sample)[0]))
    phase = np.concatenate((phase, phase))
    return float((phase[1:] - phase[:-1]).max())


def phase_coverage(sample, data, n_bins=10):
    """Fraction of n_bins phase bins that contain an observation (:114-126)."""
    H, _ = np.histogram(data.phase(_P_days(sample)[0]), bins=np.linspace(0, 1, n_bins + 1))
    return float((H > 0).sum() / n_bins)


def periods_spanned(sample, data):
    """Number of periods the data baseline covers (:129-139)."""
    return float(np.ptp(data._t_bmjd) / _P_days(sample)[0])


def phase_coverage_per_period(sample, data):
    """Largest number of observations inside any single period window (:142-151)."""
    cycles = (data._t_bmjd - data._t_ref_bmjd) / _P_days(sample)[0]
    H1, _ = np.histogram(cycles, bins=np.arange(0, cycles.max() + 1, 1))
    H2, _ = np.histogram(cycles, bins=np.arange(-0.5, cycles.max() + 1, 1))
    return int(max(H1.max(), H2.max()))
