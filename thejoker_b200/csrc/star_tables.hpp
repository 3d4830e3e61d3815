// star_tables.hpp -- host-side (no CUDA) construction of the per-star constants the
// kernels consume: the centred data vector, the epoch tables and the StarParams
// blocks.  All sums that do not depend on the prior sample are evaluated here once,
// in long double.  Replaces the data-extraction half of CJokerHelper.__init__
// (thejoker/src/fast_likelihood.pyx:161-205).
#pragma once

#include <cmath>
#include <cstring>
#include <vector>

#include "marginal_ll.cuh"

namespace tjb {

struct StarHost {
  int N = 0, L = 0;
  double t_ref = 0;
  std::vector<double> t, rv, ivar, trend;  // trend [N, L-1] row-major
  double mu[kMaxLinear] = {0}, Lambda[kMaxLinear] = {0};
  int K_prior_kind = 0, jitter_mode = 0;
  double sigma_K0 = 0, P0 = 1, max_K = 0;
  // derived
  bool centred = false;
  long double centre = 0;
  std::vector<double> yc;  // y - centre
};

// Centring: y' = y - c and mu_v0' = mu_v0 - c leave the marginal likelihood unchanged
// when the first trend column is the constant term (M x = M x' + c 1), and remove the
// part of the cancellation in chi2 = y^T C^-1 y - h^T A h that a large systemic
// velocity causes.  c = ivar-weighted mean of y, rounded to double.
inline void star_prepare(StarHost &st) {
  const int N = st.N, L = st.L;
  st.centred = (L > 1);
  for (int n = 0; n < N && st.centred; n++)
    if (st.trend[(size_t)n * (L - 1)] != 1.0) st.centred = false;
  st.centre = 0;
  if (st.centred) {
    long double sw = 0, swy = 0;
    for (int n = 0; n < N; n++) {
      sw += (long double)st.ivar[n];
      swy += (long double)st.ivar[n] * (long double)st.rv[n];
    }
    st.centre = (sw > 0) ? (long double)(double)(swy / sw) : 0.0L;
  }
  st.yc.resize(N);
  for (int n = 0; n < N; n++) st.yc[n] = (double)((long double)st.rv[n] - st.centre);
}

inline void star_fill_common(const StarHost &st, StarParams &sp) {
  memset(&sp, 0, sizeof(sp));
  sp.n_times = st.N;
  sp.mu_K = st.mu[0];
  sp.Lambda_K = st.Lambda[0];
  sp.K_prior_kind = st.K_prior_kind;
  sp.sigma_K0_sq = st.sigma_K0 * st.sigma_K0;
  sp.inv_P0 = 1.0 / st.P0;
  sp.max_K_sq = st.max_K * st.max_K;
  sp.apply_jitter = st.jitter_mode;
  sp.zero = 0.0;
  for (int i = 1; i < st.L; i++) sp.inv_Lambda[i] = (double)(1.0L / (long double)st.Lambda[i]);
}

// sin / cos at the kTrigTableSize nodes of the trig table, correctly rounded from long double
// (TJB_TRIG2: followed by the 2^kFineLog2 fine nodes, angles j 2^-kFineLog2 of a node spacing)
inline std::vector<SinCos> make_trig_table() {
  std::vector<SinCos> t(kTrigNodes > 0 ? kTrigNodes : 1);
  const long double two_pi = 6.283185307179586476925286766559005768L;
  for (int j = 0; j < kTrigTableSize; j++) {
    // exact octant symmetry: evaluate in the first octant-ish range for best accuracy
    const long double ang = two_pi * (long double)j / (long double)kTrigTableSize;
    t[j].s = (double)sinl(ang);
    t[j].c = (double)cosl(ang);
  }
  if (kTrigTableSize >= 4) {  // the four axis nodes exactly
    const int q = kTrigTableSize / 4;
    t[0] = {0.0, 1.0}; t[q] = {1.0, 0.0}; t[2 * q] = {0.0, -1.0}; t[3 * q] = {-1.0, 0.0};
  }
  for (int j = kTrigTableSize; j < kTrigNodes; j++) {
    const long double ang = two_pi * (long double)(j - kTrigTableSize) /
                            ((long double)kTrigTableSize * (long double)(1 << kFineLog2));
    t[j].s = (double)sinl(ang);
    t[j].c = (double)cosl(ang);
  }
  return t;
}

constexpr long double kLog2PiL = 1.8378770664093454835606594728112353L;

// constant-jitter table for jitter s: rows [dt, w, w y, w T_1 .. w T_{L-1}]
inline void star_build_const(const StarHost &st, double s, StarParams &sp, std::vector<double> &tab) {
  const int N = st.N, L = st.L, RS = row_stride(L);
  const long double s2 = st.jitter_mode ? (long double)s * (long double)s : 0.0L;
  tab.assign((size_t)N * RS, 0.0);
  long double mu_c[kMaxLinear];
  for (int i = 0; i < L; i++) mu_c[i] = st.mu[i];
  if (st.centred) mu_c[1] -= st.centre;
  long double G[kMaxLinear][kMaxLinear] = {{0}}, hy[kMaxLinear] = {0}, sumlogw = 0, Syy = 0;
  for (int n = 0; n < N; n++) {
    const long double iv = st.ivar[n];
    const long double w = iv / (1.0L + s2 * iv);  // pyx:48-67
    sumlogw += logl(w);
    const long double y = st.yc[n];
    Syy += w * y * y;
    double *row = &tab[(size_t)n * RS];
    row[0] = st.t[n] - st.t_ref;
    row[1] = (double)w;
    row[2] = (double)(w * y);
    for (int i = 1; i < L; i++) {
      const long double Ti = st.trend[(size_t)n * (L - 1) + (i - 1)];
      row[2 + i] = (double)(w * Ti);
      hy[i] += w * Ti * y;
      for (int j = i; j < L; j++)
        G[i][j] += w * Ti * (long double)st.trend[(size_t)n * (L - 1) + (j - 1)];
    }
  }
  star_fill_common(st, sp);
  long double quad0 = Syy, c0 = N * kLog2PiL - sumlogw;
  for (int i = 1; i < L; i++) {
    const long double il = 1.0L / (long double)st.Lambda[i];
    G[i][i] += il;
    sp.hc[i] = (double)(hy[i] + mu_c[i] * il);
    quad0 += mu_c[i] * mu_c[i] * il;
    c0 += logl((long double)st.Lambda[i]);
  }
  for (int i = 0; i < L; i++)  // packed upper triangle, same indexing as tri<L>()
    for (int j = i; j < L; j++) sp.Gc[i * L - (i * (i - 1)) / 2 + (j - i)] = (double)G[i][j];
  sp.quad0 = (double)quad0;
  sp.c0 = (double)c0;
}

// per-sample-jitter table: rows [dt, 1/ivar, y, T_1 .. T_{L-1}]
inline void star_build_jit(const StarHost &st, StarParams &sp, std::vector<double> &tab) {
  const int N = st.N, L = st.L, RS = row_stride(L);
  tab.assign((size_t)N * RS, 0.0);
  long double mu_c[kMaxLinear];
  for (int i = 0; i < L; i++) mu_c[i] = st.mu[i];
  if (st.centred) mu_c[1] -= st.centre;
  for (int n = 0; n < N; n++) {
    double *row = &tab[(size_t)n * RS];
    row[0] = st.t[n] - st.t_ref;
    row[1] = (double)(1.0L / (long double)st.ivar[n]);
    row[2] = st.yc[n];
    for (int i = 1; i < L; i++) row[2 + i] = st.trend[(size_t)n * (L - 1) + (i - 1)];
  }
  star_fill_common(st, sp);
  long double quad0 = 0, c0 = N * kLog2PiL;
  for (int i = 1; i < L; i++) {
    const long double il = 1.0L / (long double)st.Lambda[i];
    sp.hc[i] = (double)(mu_c[i] * il);
    quad0 += mu_c[i] * mu_c[i] * il;
    c0 += logl((long double)st.Lambda[i]);
  }
  sp.quad0 = (double)quad0;
  sp.c0 = (double)c0;
}

}  // namespace tjb
