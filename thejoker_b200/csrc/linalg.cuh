// linalg.cuh -- per-sample small linear algebra of the marginal likelihood, in
// registers, for L = n_linear columns (L <= 8).
//
// Replaces make_AAinv / make_bBBinv / likelihood_worker of
// thejoker/src/fast_likelihood.pyx:255-425.  The reference forms the N x N matrix
// B = C + M Lambda M^T, LU-factors it for log det(2 pi B) and builds B^-1 by
// Woodbury (O(N^3) per sample).  With the matrix-determinant lemma and the
// Woodbury identity the same value needs only the (L+1)(L+2)/2 Gram sums over
// epochs:
//    Ainv   = Lambda^-1 + M^T C^-1 M                      (pyx:266-278)
//    h      = M^T C^-1 y + Lambda^-1 mu                   (pyx:396-407)
//    chi2   = y^T C^-1 y + mu^T Lambda^-1 mu - h^T Ainv^-1 h
//    logdet = N log 2pi - sum log ivar + sum log Lambda + log det Ainv
//    ll     = -(chi2 + logdet) / 2                        (pyx:425)
// Ainv is SPD, so an unrolled LDL^T replaces dgetrf/dgetri/dsysv.
#pragma once

#include "kepler.cuh"

namespace tjb {

constexpr int kMaxLinear = 8;

// packed upper-triangle index of a symmetric L x L matrix, i <= j
template <int L>
TJB_HD constexpr int tri(int i, int j) { return i * L - (i * (i - 1)) / 2 + (j - i); }
template <int L>
constexpr int kTri = L * (L + 1) / 2;

// reciprocal of a non-zero normal double (either sign): MUFU seed + third-order step
TJB_HD double rcp_nz(double x) {
#if defined(__CUDA_ARCH__)
  return rcp_pos(x);
#else
  return 1.0 / x;
#endif
}

// LDL^T of the packed SPD matrix G (overwritten: D on the diagonal, the rows of
// L^T above it); rD receives 1/D.  Returns false if a pivot is exactly zero (the
// analogue of dgetrf info != 0, pyx:283-284).
template <int L>
TJB_HD bool ldlt(double *G, double *rD) {
  bool ok = true;
#pragma unroll
  for (int j = 0; j < L; j++) {
    double d = G[tri<L>(j, j)];
#pragma unroll
    for (int k = 0; k < j; k++) d = fma(-G[tri<L>(k, j)] * G[tri<L>(k, j)], G[tri<L>(k, k)], d);
    G[tri<L>(j, j)] = d;
    ok = ok && (d != 0.0);
    const double rd = rcp_nz(d);
    rD[j] = rd;
#pragma unroll
    for (int i = j + 1; i < L; i++) {
      double v = G[tri<L>(j, i)];
#pragma unroll
      for (int k = 0; k < j; k++)
        v = fma(-G[tri<L>(k, j)] * G[tri<L>(k, i)], G[tri<L>(k, k)], v);
      G[tri<L>(j, i)] = v * rd;  // L_ij
    }
  }
  return ok;
}

// given the factorisation from ldlt(): quad = h^T G^-1 h and prod = det G
template <int L>
TJB_HD void ldlt_quad(const double *G, const double *rD, const double *h, double &quad,
                      double &detG) {
  double y[L];
  quad = 0.0;
  detG = 1.0;
#pragma unroll
  for (int j = 0; j < L; j++) {
    double v = h[j];
#pragma unroll
    for (int k = 0; k < j; k++) v = fma(-G[tri<L>(k, j)], y[k], v);
    y[j] = v;
    quad = fma(v * v, rD[j], quad);
    detG *= G[tri<L>(j, j)];
  }
}

// solve G x = h in place (x returned in h) from the factorisation
template <int L>
TJB_HD void ldlt_solve(const double *G, const double *rD, double *h) {
#pragma unroll
  for (int j = 0; j < L; j++) {
#pragma unroll
    for (int k = 0; k < j; k++) h[j] = fma(-G[tri<L>(k, j)], h[k], h[j]);
  }
#pragma unroll
  for (int j = 0; j < L; j++) h[j] = h[j] * rD[j];
#pragma unroll
  for (int j = L - 1; j >= 0; j--) {
#pragma unroll
    for (int k = j + 1; k < L; k++) h[j] = fma(-G[tri<L>(j, k)], h[k], h[j]);
  }
}

// prior variance of K for the FixedCompanionMass prior (pyx:461-464,
// distributions.py:143-147): min(max_K^2, sigma_K0^2 / (1 - e^2) (P/P0)^(-2/3)).
TJB_HD double lambda_K_fixed_mass(double P, double e, double sigma_K0_sq, double inv_P0,
                                  double max_K_sq, bool clamp) {
  const double x = P * inv_P0;
#if defined(__CUDA_ARCH__)
  const double rc = rcbrt(x);
#else
  const double rc = 1.0 / cbrt(x);
#endif
  const double lam = sigma_K0_sq * rcp_nz(fma(-e, e, 1.0)) * (rc * rc);
  return clamp ? fmin(max_K_sq, lam) : lam;
}

}  // namespace tjb
