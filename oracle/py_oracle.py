"""numpy restatement of the reference's own pure-Python test oracle
(thejoker/src/tests/py_likelihood.py), with the *intended* jitter semantics.

TEST INFRASTRUCTURE ONLY.  Small cases only (Python loop per sample).
Used to cross-check oracle/joker_oracle.c independently of LAPACK call order.
"""
import numpy as np


def kepler_E(M, e, tol=1e-13, maxiter=128):
    """Newton solve, vectorised over epochs (restates twobody; see joker_oracle.c)."""
    M = np.asarray(M, dtype=float)
    E = M + e * np.sin(M) + 0.5 * e * e * np.sin(2 * M)
    for _ in range(maxiter):
        dM = M - (E - e * np.sin(E))
        E = E + dM / (1 - e * np.cos(E))
        if np.all(np.abs(dM) < tol):
            break
    return E


def design_matrix(row, t, t0, trend_M):
    """py_likelihood.py:110-131 (generalised to offsets through trend_M)."""
    P, e, om, M0 = row[:4]
    M = 2 * np.pi * (t - t0) / P - M0
    E = kepler_E(M, e)
    f = 2 * np.arctan2(np.sqrt(1 + e) * np.sin(E / 2), np.sqrt(1 - e) * np.cos(E / 2))
    z = np.cos(om + f) + e * np.cos(om)
    return np.hstack((z[:, None], trend_M))


def log_multivariate_gaussian(x, mu, V, Vinv):
    """astroML.utils.log_multivariate_gaussian as used at py_likelihood.py:98."""
    dx = x - mu
    Vchol = np.linalg.cholesky(V)
    logdet = 2 * np.sum(np.log(np.diag(Vchol)))
    chi2 = dx @ Vinv @ dx
    return -0.5 * logdet - 0.5 * chi2 - 0.5 * len(x) * np.log(2 * np.pi)


def likelihood_worker(y, ivar, M, mu, Lambda, make_aA=False):
    """py_likelihood.py:32-107."""
    Lam = np.diag(Lambda)
    Laminv = np.diag(1 / Lambda)
    Cinv = np.diag(ivar)
    C = np.diag(1 / ivar)
    b = M @ mu
    B = C + M @ Lam @ M.T
    Ainv = Laminv + M.T @ Cinv @ M
    A = np.linalg.inv(Ainv)
    Binv = Cinv - Cinv @ M @ A @ M.T @ Cinv
    ll = log_multivariate_gaussian(y, b, B, Binv)
    if make_aA:
        a = np.linalg.solve(Ainv, Laminv @ mu + M.T @ Cinv @ y)
        return ll, b, B, a, A
    return ll, b, B


def lambda_K(P, e, sigma_K0, P0, max_K, clamp=True):
    """py_likelihood.py:165-166."""
    lam = sigma_K0**2 / (1 - e**2) * (P / P0) ** (-2 / 3)
    return min(max_K**2, lam) if clamp else lam


def marginal_ln_likelihood(chunk, spec, apply_jitter=True):
    """py_likelihood.py:175-205 on a plain spec dict (see OracleHelper.from_spec)."""
    t, y, ivar0 = spec["t"], spec["rv"], spec["ivar"]
    mu = np.asarray(spec["mu"], dtype=float)
    out = np.zeros(len(chunk))
    for n, row in enumerate(chunk):
        Lam = np.array(spec["Lambda"], dtype=float)
        if spec["K_prior_kind"] == 0:
            Lam[0] = lambda_K(row[0], row[1], spec["sigma_K0"], spec["P0"], spec["max_K"])
        s = row[4] if apply_jitter else 0.0
        ivar = ivar0 / (1 + s**2 * ivar0)
        M = design_matrix(row, t, spec["t0"], spec["trend_M"])
        out[n] = likelihood_worker(y, ivar, M, mu, Lam)[0]
    return out
