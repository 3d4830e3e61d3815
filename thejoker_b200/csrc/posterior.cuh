// posterior.cuh -- linear-parameter posterior (a, A) and draws for accepted
// samples; design-column probe; FP64 peak probe.
//
// Replaces the per-sample body of CJokerHelper.batch_get_posterior_samples and
// test_likelihood_worker (thejoker/src/fast_likelihood.pyx:471-576):
//   a = Ainv^-1 (M^T C^-1 y + Lambda^-1 mu)   (pyx:394-420, dsysv)
//   A = Ainv^-1                               (pyx:530, np.linalg.inv(self.Ainv))
//   x ~ N(a, A)                               (pyx:529-530)
// The table is the per-sample-jitter table of marginal_ll.cuh ([dt, var, y, T..]);
// y is centred there, `center` is added back to a[v0].
#pragma once

#include "marginal_ll.cuh"

namespace tjb {

#if defined(__CUDACC__)

template <int L>
__global__ void __launch_bounds__(128)
posterior_kernel(const StarParams sp, const double *__restrict__ rows, const int k,
                 const int clamp_K, const double center, const int center_col,
                 const int n_per, const double *__restrict__ normals,
                 double *__restrict__ ll_out, double *__restrict__ a_out,
                 double *__restrict__ A_out, double *__restrict__ draws_out) {
  constexpr int RS = row_stride(L);
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = gid < k;
  const int r = valid ? gid : k - 1;
  const double P = rows[5 * r + 0], e = rows[5 * r + 1], om = rows[5 * r + 2],
               M0 = rows[5 * r + 3], s = rows[5 * r + 4];
  TrigCoef tc;
  tc.load(sp.zero, sp.trig_table);  // a few rows only: the table is read from global memory
  const OrbitConsts oc = make_orbit_consts(tc, P, e, om, M0);
  const double s2 = sp.apply_jitter ? s * s : 0.0;

  double G[kTri<L>], h[L], Syy = 0.0;
#pragma unroll
  for (int i = 0; i < kTri<L>; i++) G[i] = 0.0;
#pragma unroll
  for (int i = 0; i < L; i++) h[i] = 0.0;
  LogProduct lp;
  lp.init();
  for (int n = 0; n < sp.n_times; n++) {
    const double *row = sp.table + n * RS;
    const double z = rv_unit_column<false>(oc, tc, row[0], nullptr);
    const double var = row[1] + s2;
    const double w = 1.0 / var;
    lp.mul(var);
    lp.renorm();
    const double y = row[2];
    double m[L];
    m[0] = z;
#pragma unroll
    for (int j = 1; j < L; j++) m[j] = row[2 + j];
    const double wy = w * y;
    Syy = fma(wy, y, Syy);
#pragma unroll
    for (int i = 0; i < L; i++) {
      const double wm = w * m[i];
      h[i] = fma(wy, m[i], h[i]);
#pragma unroll
      for (int j = i; j < L; j++) G[tri<L>(i, j)] = fma(wm, m[j], G[tri<L>(i, j)]);
    }
  }
  const double lamK = sp.K_prior_kind == 0
                          ? lambda_K_fixed_mass(P, e, sp.sigma_K0_sq, sp.inv_P0, sp.max_K_sq,
                                                clamp_K != 0)
                          : sp.Lambda_K;
  const double ilamK = 1.0 / lamK;
  G[0] += ilamK;
  h[0] = fma(sp.mu_K, ilamK, h[0]);
  double quad0 = fma(sp.mu_K * sp.mu_K, ilamK, Syy + sp.quad0);
#pragma unroll
  for (int i = 1; i < L; i++) {
    G[tri<L>(i, i)] += sp.inv_Lambda[i];
    h[i] += sp.hc[i];
  }
  double rD[L];
  const bool ok = ldlt<L>(G, rD);
  double quad, detG;
  ldlt_quad<L>(G, rD, h, quad, detG);
  double ll = -0.5 * ((quad0 - quad) + (sp.c0 + lp.log_value() + log(lamK * detG)));
  if (!ok) ll = INFINITY;

  // posterior mean
  double a[L];
#pragma unroll
  for (int i = 0; i < L; i++) a[i] = h[i];
  ldlt_solve<L>(G, rD, a);
  if (!valid) return;
  if (ll_out) ll_out[r] = ll;
  if (a_out) {
#pragma unroll
    for (int i = 0; i < L; i++) a_out[r * L + i] = a[i] + (i == center_col ? center : 0.0);
  }
  if (A_out) {
#pragma unroll
    for (int c = 0; c < L; c++) {
      double col[L];
#pragma unroll
      for (int i = 0; i < L; i++) col[i] = (i == c) ? 1.0 : 0.0;
      ldlt_solve<L>(G, rD, col);
#pragma unroll
      for (int i = 0; i < L; i++) A_out[(r * L + i) * L + c] = col[i];
    }
  }
  if (draws_out) {
    // x = a + Lfac^-T D^-1/2 z  (covariance (Lfac D Lfac^T)^-1 = A)
    for (int d = 0; d < n_per; d++) {
      double x[L];
      const double *zn = normals + ((long long)r * n_per + d) * L;
#pragma unroll
      for (int i = 0; i < L; i++) x[i] = zn[i] * sqrt(rD[i]);
#pragma unroll
      for (int j = L - 1; j >= 0; j--) {
#pragma unroll
        for (int q = j + 1; q < L; q++) x[j] = fma(-G[tri<L>(j, q)], x[q], x[j]);
      }
      double *o = draws_out + ((long long)r * n_per + d) * (5 + L);
      o[0] = P; o[1] = e; o[2] = om; o[3] = M0; o[4] = s;
#pragma unroll
      for (int i = 0; i < L; i++) o[5 + i] = a[i] + x[i] + (i == center_col ? center : 0.0);
    }
  }
}

// rows[j] = [P, e, omega, M0, s][idx[j]]: the packed rows of the accepted samples, taken
// from the device-resident prior columns (multiproc_helpers.py:261-263 does this on the
// host with read_batch_idx).  s == null: every sample has jitter s_const.
__global__ void gather_rows_kernel(const double *__restrict__ P, const double *__restrict__ e,
                                   const double *__restrict__ omega, const double *__restrict__ M0,
                                   const double *__restrict__ s, const double s_const,
                                   const long long *__restrict__ idx, const int k,
                                   double *__restrict__ rows) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= k) return;
  const long long i = idx[j];
  double *o = rows + 5 * (long long)j;
  o[0] = P[i];
  o[1] = e[i];
  o[2] = omega[i];
  o[3] = M0[i];
  o[4] = s ? s[i] : s_const;
}

// z[n] for one sample, computed by every lane of one warp (lane 0 writes)
__global__ void design_column_kernel(const double *__restrict__ dt, const int N, const double P,
                                     const double e, const double om, const double M0,
                                     const double zero, const SinCos *__restrict__ trig,
                                     double *__restrict__ z, int *__restrict__ stats) {
  TrigCoef tc;
  tc.load(zero, trig);
  const OrbitConsts oc = make_orbit_consts(tc, P, e, om, M0);
  SolveStats st = {0, 0, 0};
  for (int n = 0; n < N; n++) {
    const double v = rv_unit_column<true>(oc, tc, dt[n], &st);
    if (threadIdx.x == 0) z[n] = v;
  }
  if (threadIdx.x == 0) {
    stats[0] = st.extra_f32;
    stats[1] = st.extra_f64;
    stats[2] = st.not_converged;
  }
}

// ln of the UN-marginalised likelihood of full posterior samples (samples.py:611-632):
// rows [P, e, omega, M0, s, x_0 .. x_{L-1}] with x in design-column order (K, v0,
// offsets, v1, ...); tab = the per-sample-jitter epoch table [dt, 1/ivar, y - centre,
// T_1 .. T_{L-1}].  ll = sum_n ln N(y_n | x_0 z_n + sum_k x_k T_nk, 1/ivar_n + s^2).
// One thread per sample; a handful to a few million samples, so the table is read
// through the read-only cache rather than staged.
__global__ void __launch_bounds__(128)
unmarginalized_ll_kernel(const double *__restrict__ tab, const int N, const int L, const int RS,
                         const double centre, const double *__restrict__ rows, const long long n,
                         const double zero, const SinCos *__restrict__ trig,
                         double *__restrict__ ll) {
  TrigCoef tc;
  tc.load(zero, trig);
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long n_pad = (n + 31) & ~31LL;  // whole warps: the solver votes across lanes
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += stride) {
    const bool live = i < n;
    const double *row = rows + (live ? i : 0) * (5 + L);
    const OrbitConsts oc = make_orbit_consts(tc, row[0], row[1], row[2], row[3]);
    const double s2 = row[4] * row[4];
    double acc = 0.0;
    for (int m = 0; m < N; m++) {
      const double *tr = tab + (size_t)m * RS;
      const double z = rv_unit_column<false>(oc, tc, __ldg(tr), nullptr);
      double model = row[5] * z;
      if (L > 1) model = fma(row[6] - centre, __ldg(tr + 3), model);
      for (int k = 2; k < L; k++) model = fma(row[5 + k], __ldg(tr + 2 + k), model);
      const double var = __ldg(tr + 1) + s2;
      const double r = __ldg(tr + 2) - model;
      acc += r * r / var + log(6.283185307179586 * var);
    }
    if (live) ll[i] = -0.5 * acc;
  }
}

// dependent-FMA chains, 8 independent accumulators per thread
__global__ void __launch_bounds__(256) fp64_peak_kernel(const int iters, double *out) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5,
         a6 = a0 + 6, a7 = a0 + 7;
  const double m = 0.999999, c = 1e-7;
  for (int i = 0; i < iters; i++) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if (s == 12345.678) out[0] = s;  // keep the chains live
}

#endif  // __CUDACC__

}  // namespace tjb
