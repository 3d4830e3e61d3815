"""Design study for the fixed-work Kepler solver used by the CUDA kernel.

Evaluates starter + fixed number of Householder steps (angle-addition updates of
(sinE, cosE), one reciprocal per step) against a long-double converged Newton
truth over a dense (e, M) grid.  Run: python tools/kepler_solver_study.py
"""
import numpy as np

LD = np.float64  # long double is ~100x slower in numpy; double Newton+bisection is exact to ~1 ulp


def truth(M, e):
    M = M.astype(LD); e = e.astype(LD)
    E = np.where(e < 0.8, M, LD(np.pi) * np.ones_like(M))
    # bisection-safe Newton in long double
    lo = np.zeros_like(M); hi = LD(np.pi) * np.ones_like(M)
    for _ in range(60):
        f = E - e * np.sin(E) - M
        lo = np.where(f < 0, E, lo); hi = np.where(f > 0, E, hi)
        En = E - f / (1 - e * np.cos(E))
        bad = (En <= lo) | (En >= hi)
        En = np.where(bad, 0.5 * (lo + hi), En)
        E = En
    return E


def starter(kind, M, e, sM, cM):
    if kind == "S1":
        return e * sM * (1 + e * cM)
    if kind == "S2":
        return e * sM / np.sqrt(1 - 2 * e * cM + e * e)
    if kind == "S3":
        se, ce = np.sin(e), np.cos(e)
        return e * sM / (1 - (sM * ce + cM * se) + sM)
    if kind == "S0":
        return np.zeros_like(M)
    raise ValueError


def hstep(D, s, c, e, order=3):
    es = e * s; ec = e * c
    f = D - es
    r = 1.0 / (1.0 - ec)
    u = -f * r
    a = 0.5 * es * r
    b = ec * r / 6.0
    if order == 1:
        d = u
    elif order == 2:
        d = u * (1 - a * u)
    elif order == 3:
        d = u * (1 + u * (-a + u * (2 * a * a - b)))
    elif order == 4:
        # series reversion of u = d + a d^2 + b d^3 + g d^4, g = -es r/24
        g = -es * r / 24.0
        d = u * (1 + u * (-a + u * ((2 * a * a - b) + u * (-5 * a ** 3 + 5 * a * b - g))))
    sd, cd = np.sin(d), np.cos(d)
    return D + d, s * cd + c * sd, c * cd - s * sd, d


def run(kind, orders, emax, n_e=400, n_M=4000):
    e = np.linspace(0, emax, n_e)
    Mlin = np.linspace(0, np.pi, n_M)
    Mlog = np.logspace(-8, np.log10(np.pi), n_M)
    M = np.concatenate([Mlin, Mlog])
    ee, MM = np.meshgrid(e, M, indexing="ij")
    Et = truth(MM, ee)
    st, ct = np.sin(Et).astype(float), np.cos(Et).astype(float)
    sM, cM = np.sin(MM), np.cos(MM)
    D = starter(kind, MM, ee, sM, cM)
    err0 = np.abs((MM + D) - Et.astype(float)).max()
    s, c = np.sin(MM + D), np.cos(MM + D)
    dmax = []
    for o in orders:
        D, s, c, d = hstep(D, s, c, ee, o)
        dmax.append(np.abs(d).max())
    err = np.maximum(np.abs(s - st), np.abs(c - ct))
    worst = np.unravel_index(err.argmax(), err.shape)
    return err0, err.max(), dmax, (ee[worst], MM[worst])


if __name__ == "__main__":
    for emax in (0.8, 0.9, 0.95, 0.99):
        for kind in ("S2", "S3"):
            for orders in ((3, 3), (4, 3), (3, 3, 3)):
                e0, err, dmax, w = run(kind, orders, emax, 100, 1500)
                print(f"emax={emax} {kind} {orders}: start err {e0:.2e} final {err:.2e} "
                      f"dmax {[f'{d:.1e}' for d in dmax]} worst e={w[0]:.3f} M={w[1]:.2e}")
