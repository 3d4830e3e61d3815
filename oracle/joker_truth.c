/*
 * joker_truth.c -- quad-precision (__float128) evaluation of the marginal
 * likelihood model of The Joker, used to judge which of two double-precision
 * implementations is closer to the exact value on ill-conditioned inputs.
 *
 * TEST INFRASTRUCTURE ONLY (same rules as joker_oracle.c).
 *
 * Model (thejoker/src/fast_likelihood.pyx:266, 306-339, 344-357, 425 and
 * src/tests/py_likelihood.py:70-107):
 *   y = M x + eps, eps ~ N(0, C), C = diag(1/ivar), x ~ N(mu, Lambda),
 *   ll = log N(y | M mu, C + M Lambda M^T).
 * Evaluated here in the cancellation-free residual form
 *   chi2   = sum_n ivar_n (y_n - (M a)_n)^2 + sum_i (a_i - mu_i)^2 / Lambda_i
 *   logdet = N log 2pi - sum log ivar_n + sum log Lambda_i + log det(Ainv)
 * with Ainv = Lambda^-1 + M^T C^-1 M and a = Ainv^-1 (M^T C^-1 y + Lambda^-1 mu).
 */
#include <quadmath.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "joker_oracle.h"

typedef __float128 q_t;
#define MAXL 16

static q_t kepler_q(q_t M, q_t e) {
  /* reduce to [-pi, pi], bisection-safeguarded Newton to full quad precision */
  q_t twopi = 2 * M_PIq;
  q_t Mr = M - twopi * rintq(M / twopi);
  q_t sgn = Mr < 0 ? -1 : 1;
  q_t Ma = fabsq(Mr);
  q_t lo = 0, hi = M_PIq, E = Ma + e * sinq(Ma);
  if (E > hi) E = hi;
  for (int it = 0; it < 200; it++) {
    q_t f = E - e * sinq(E) - Ma;
    if (f > 0) hi = E; else lo = E;
    q_t En = E - f / (1 - e * cosq(E));
    if (!(En > lo && En < hi)) En = (lo + hi) / 2;
    if (fabsq(En - E) < 1e-32q * (1 + fabsq(En))) { E = En; break; }
    E = En;
  }
  return sgn * E + (M - Mr);
}

/* solve S x = r for symmetric positive-definite S (L<=MAXL) by Gaussian
 * elimination with partial pivoting; returns log|det S| */
static q_t solve_q(int L, q_t S[MAXL][MAXL], q_t *r, q_t inv[MAXL][MAXL]) {
  q_t Aug[MAXL][2 * MAXL + 1];
  q_t logdet = 0;
  for (int i = 0; i < L; i++) {
    for (int j = 0; j < L; j++) { Aug[i][j] = S[i][j]; Aug[i][L + 1 + j] = (i == j); }
    Aug[i][L] = r[i];
  }
  int W = 2 * L + 1;
  for (int c = 0; c < L; c++) {
    int p = c;
    for (int i = c + 1; i < L; i++)
      if (fabsq(Aug[i][c]) > fabsq(Aug[p][c])) p = i;
    if (p != c)
      for (int k = 0; k < W; k++) { q_t t = Aug[c][k]; Aug[c][k] = Aug[p][k]; Aug[p][k] = t; }
    logdet += logq(fabsq(Aug[c][c]));
    for (int i = 0; i < L; i++) {
      if (i == c) continue;
      q_t f = Aug[i][c] / Aug[c][c];
      for (int k = c; k < W; k++) Aug[i][k] -= f * Aug[c][k];
    }
  }
  for (int i = 0; i < L; i++) {
    r[i] = Aug[i][L] / Aug[i][i];
    if (inv)
      for (int j = 0; j < L; j++) inv[i][j] = Aug[i][L + 1 + j] / Aug[i][i];
  }
  return logdet;
}

static q_t one_sample(const OrcSpec *sp, const double *row, int clamp, q_t *a_out,
                      q_t inv_out[MAXL][MAXL], q_t *kappa) {
  int N = sp->n_times, L = sp->n_linear;
  q_t P = row[0], e = row[1], om = row[2], M0 = row[3], s = row[4];
  q_t Lam[MAXL], mu[MAXL], S[MAXL][MAXL], rhs[MAXL], rhs0[MAXL];
  q_t *Mq = (q_t *)malloc(sizeof(q_t) * (size_t)N * L);
  q_t *w = (q_t *)malloc(sizeof(q_t) * N);
  for (int i = 0; i < L; i++) { Lam[i] = sp->Lambda[i]; mu[i] = sp->mu[i]; }
  if (sp->K_prior_kind == 0) {
    q_t sk = sp->sigma_K0, P0 = sp->P0, mk = sp->max_K;
    Lam[0] = sk * sk / (1 - e * e) * powq(P / P0, -2 / 3.q);
    if (clamp && Lam[0] > mk * mk) Lam[0] = mk * mk;
  }
  q_t sumlogw = 0;
  for (int n = 0; n < N; n++) {
    q_t iv = sp->ivar[n];
    w[n] = sp->jitter_mode ? iv / (1 + s * s * iv) : iv;
    sumlogw += logq(w[n]);
    q_t M = 2 * M_PIq * ((q_t)sp->t[n] - (q_t)sp->t0) / P - M0;
    q_t E = kepler_q(M, e);
    q_t f = 2 * atan2q(sqrtq(1 + e) * sinq(E / 2), sqrtq(1 - e) * cosq(E / 2));
    Mq[n * L + 0] = cosq(om + f) + e * cosq(om);
    for (int i = 1; i < L; i++) Mq[n * L + i] = sp->trend_M[n * (L - 1) + (i - 1)];
  }
  q_t dCd = 0;
  for (int i = 0; i < L; i++) {
    rhs[i] = mu[i] / Lam[i];
    for (int j = 0; j < L; j++) S[i][j] = (i == j) ? 1 / Lam[i] : 0;
  }
  for (int n = 0; n < N; n++) {
    q_t y = sp->rv[n], d = y;
    for (int i = 0; i < L; i++) {
      d -= Mq[n * L + i] * mu[i];
      rhs[i] += Mq[n * L + i] * w[n] * y;
      for (int j = 0; j < L; j++) S[i][j] += Mq[n * L + i] * w[n] * Mq[n * L + j];
    }
    dCd += w[n] * d * d;
  }
  memcpy(rhs0, rhs, sizeof(rhs));
  q_t logdetS = solve_q(L, S, rhs, inv_out);
  q_t chi2 = 0, sumlogLam = 0;
  for (int n = 0; n < N; n++) {
    q_t r = sp->rv[n];
    for (int i = 0; i < L; i++) r -= Mq[n * L + i] * rhs[i];
    chi2 += w[n] * r * r;
  }
  for (int i = 0; i < L; i++) {
    chi2 += (rhs[i] - mu[i]) * (rhs[i] - mu[i]) / Lam[i];
    sumlogLam += logq(Lam[i]);
  }
  q_t logdet = N * logq(2 * M_PIq) - sumlogw + sumlogLam + logdetS;
  if (a_out)
    for (int i = 0; i < L; i++) a_out[i] = rhs[i];
  if (kappa) *kappa = dCd / chi2;
  free(Mq); free(w);
  return -(chi2 + logdet) / 2;
}

int orc_truth_marginal_ln_likelihood(const OrcSpec *sp, const double *chunk, long n_samples,
                                     double *ll, double *kappa, int n_threads) {
  if (sp->n_linear > MAXL) return -1;
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#pragma omp parallel for schedule(dynamic, 16)
#endif
  for (long n = 0; n < n_samples; n++) {
    q_t k;
    ll[n] = (double)one_sample(sp, chunk + 5 * n, 1, 0, 0, &k);
    if (kappa) kappa[n] = (double)k;
  }
  return 0;
}

int orc_truth_posterior_aA(const OrcSpec *sp, const double *row, int clamp, double *a,
                           double *A) {
  int L = sp->n_linear;
  if (L > MAXL) return -1;
  q_t aq[MAXL], inv[MAXL][MAXL];
  one_sample(sp, row, clamp, aq, inv, 0);
  for (int i = 0; i < L; i++) {
    a[i] = (double)aq[i];
    for (int j = 0; j < L; j++) A[i * L + j] = (double)inv[i][j];
  }
  return 0;
}
