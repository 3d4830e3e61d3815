/*
 * joker_oracle.c -- CPU restatement of The Joker's marginal-likelihood hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in thejoker_b200/ (the product) may import,
 * link or call this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker / the
 * timed CPU baseline.
 *
 * PARITY STATUS: pinned.  Everything the reference itself implements is pinned bit for
 * bit to its own compiled Cython; the one function it takes from a third party is pinned
 * to that third party's own outputs as stored in the reference repository.
 *   - thejoker/src/fast_likelihood.pyx is translated by Cython and compiled unmodified
 *     from /root/reference (oracle/ref_build/build_ref.py -> oracle/_ref/, git-ignored)
 *     and driven through its real CJokerHelper.__init__ and public methods
 *     (oracle/ref_cython.py; astropy / pymc / pytensor are absent, so duck-typed
 *     stand-ins satisfy its imports).  On every vector of tests/golden/ref_*.npz
 *     (tests/golden/make_ref_golden.py) and on fresh seeded inputs
 *     (tests/test_ref_pinning.py) this file reproduces the reference's ll, a, A, Ainv,
 *     b, B, Binv and posterior samples BIT FOR BIT.
 *   - The Kepler part restates the published algorithm of the third-party dependency
 *     `twobody` (>=0.9.1, unpinned; twobody/src/twobody.c :: c_rv_from_elements), which
 *     is not under /root/reference and not installed, so the compiled reference above
 *     links THIS file's orc_rv_from_elements in its place: Newton iteration on
 *     E - e sin E = M from a second-order series starter, tolerance and maxiter as
 *     passed by fast_likelihood.pyx:35-36, true anomaly by the half-angle atan2 formula,
 *     rv = K (cos(f + omega) + e cos omega).  It is pinned to twobody's OWN OUTPUTS: the
 *     reference's docs/examples/{data,data-triple,data-survey1,data-survey2}.ecsv hold
 *     noiseless `twobody.KeplerOrbit.radial_velocity(t)` values (make-data.ipynb adds no
 *     noise to rv) whose true elements follow from the notebook's seed; on all 531
 *     epochs (e = 0.1, 0.13, 0.25, a two-orbit sum, a survey offset) this function
 *     reproduces them to 2.8e-10 km/s = 4e-11 K, which is the resolution of the files'
 *     float64 Julian dates (tests/golden/make_ref_examples_golden.py,
 *     tests/test_ref_pinning.py::test_oracle_kepler_reproduces_twobody_outputs).  What
 *     stays unpinned is only the last-ulp behaviour of twobody's loop exit; the two
 *     plausible readings differ by <= 3e-11 relative in ll, and joker_truth.c (quad
 *     precision) bounds both.
 * The linear-algebra part follows the pyx statement by statement and calls the same
 * LAPACK entry points (scipy.linalg.cython_lapack dgetrf/dgetri/dsysv, bound at run time
 * from Python through orc_set_lapack).
 *
 * Each function cites the reference lines it follows.
 */
#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "joker_oracle.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------------- */
/* LAPACK binding (fast_likelihood.pyx:19 cimports scipy.linalg.cython_lapack) */

typedef void (*dgetrf_t)(int *, int *, double *, int *, int *, int *);
typedef void (*dgetri_t)(int *, double *, int *, int *, double *, int *, int *);
typedef void (*dsysv_t)(char *, int *, int *, double *, int *, int *, double *,
                        int *, double *, int *, int *);

static dgetrf_t p_dgetrf = 0;
static dgetri_t p_dgetri = 0;
static dsysv_t p_dsysv = 0;

void orc_set_lapack(void *dgetrf, void *dgetri, void *dsysv) {
  p_dgetrf = (dgetrf_t)dgetrf;
  p_dgetri = (dgetri_t)dgetri;
  p_dsysv = (dsysv_t)dsysv;
}

int orc_has_lapack(void) { return p_dgetrf && p_dgetri && p_dsysv; }

/* Built-in stand-ins (column-major, unblocked LU with partial pivoting, the
 * algorithm of LAPACK dgetf2) used only when no LAPACK was bound, so the oracle
 * stays usable from plain C.  tests/test_oracle.py checks them against LAPACK. */
static void own_dgetrf(int *m, int *n_, double *a, int *lda, int *ipiv, int *info) {
  int n = *n_, ld = *lda;
  (void)m;
  *info = 0;
  for (int j = 0; j < n; j++) {
    int p = j;
    double big = fabs(a[j + j * ld]);
    for (int i = j + 1; i < n; i++) {
      double v = fabs(a[i + j * ld]);
      if (v > big) { big = v; p = i; }
    }
    ipiv[j] = p + 1;
    if (a[p + j * ld] != 0.0) {
      if (p != j)
        for (int k = 0; k < n; k++) {
          double tmp = a[j + k * ld];
          a[j + k * ld] = a[p + k * ld];
          a[p + k * ld] = tmp;
        }
      double r = 1.0 / a[j + j * ld];
      for (int i = j + 1; i < n; i++) a[i + j * ld] *= r;
    } else if (*info == 0) {
      *info = j + 1;
    }
    for (int k = j + 1; k < n; k++) {
      double akj = a[j + k * ld];
      for (int i = j + 1; i < n; i++) a[i + k * ld] -= a[i + j * ld] * akj;
    }
  }
}

static void own_lu_solve(int n, const double *lu, int ld, const int *ipiv, double *x) {
  for (int i = 0; i < n; i++) {
    int p = ipiv[i] - 1;
    if (p != i) { double t = x[i]; x[i] = x[p]; x[p] = t; }
  }
  for (int i = 0; i < n; i++)
    for (int k = 0; k < i; k++) x[i] -= lu[i + k * ld] * x[k];
  for (int i = n - 1; i >= 0; i--) {
    for (int k = i + 1; k < n; k++) x[i] -= lu[i + k * ld] * x[k];
    x[i] /= lu[i + i * ld];
  }
}

static void own_dgetri(int *n_, double *a, int *lda, int *ipiv, double *work,
                       int *lwork, int *info) {
  int n = *n_, ld = *lda;
  (void)work; (void)lwork;
  *info = 0;
  for (int i = 0; i < n; i++)
    if (a[i + i * ld] == 0.0) { *info = i + 1; return; }
  double *inv = (double *)malloc(sizeof(double) * n * n);
  for (int j = 0; j < n; j++) {
    double *col = inv + j * n;
    for (int i = 0; i < n; i++) col[i] = (i == j) ? 1.0 : 0.0;
    own_lu_solve(n, a, ld, ipiv, col);
  }
  for (int j = 0; j < n; j++)
    for (int i = 0; i < n; i++) a[i + j * ld] = inv[i + j * n];
  free(inv);
}

static void own_dsysv(char *uplo, int *n_, int *nrhs, double *a, int *lda, int *ipiv,
                      double *b, int *ldb, double *work, int *lwork, int *info) {
  /* symmetric solve: reconstruct the full matrix from the stored triangle, LU it */
  int n = *n_, ld = *lda;
  (void)nrhs; (void)ldb; (void)work; (void)lwork;
  for (int j = 0; j < n; j++)
    for (int i = 0; i < j; i++) {
      if (*uplo == 'U') a[j + i * ld] = a[i + j * ld];
      else a[i + j * ld] = a[j + i * ld];
    }
  int m = n;
  own_dgetrf(&m, &m, a, lda, ipiv, info);
  if (*info != 0) return;
  own_lu_solve(n, a, ld, ipiv, b);
}

#define CALL_DGETRF (p_dgetrf ? p_dgetrf : own_dgetrf)
#define CALL_DGETRI (p_dgetri ? p_dgetri : own_dgetri)
#define CALL_DSYSV (p_dsysv ? p_dsysv : own_dsysv)

/* ------------------------------------------------------------------------- */
/* Kepler: restatement of twobody.c (third-party; see header note)            */

/* E - e sin E */
static double mean_from_ecc_anomaly(double E, double e) { return E - e * sin(E); }

/* Newton solve of Kepler's equation.  variant 0: apply the Newton update, then
 * stop if the mean-anomaly residual that produced it was below tol (returned E
 * is one quadratic step past the tolerance); variant 1: stop before applying
 * the update.  Both are provided because the upstream source is not available
 * here; tests report the sensitivity of ll to the choice. */
double orc_eccentric_anomaly(double M, double e, double tol, int maxiter, int variant) {
  double E, dM;
  if (e == 0.0) return M;
  E = M + e * sin(M) + 0.5 * e * e * sin(2.0 * M);
  for (int it = 0; it < maxiter; it++) {
    dM = M - mean_from_ecc_anomaly(E, e);
    if (variant == 1 && fabs(dM) < tol) break;
    E = E + dM / (1.0 - e * cos(E));
    if (variant == 0 && fabs(dM) < tol) break;
  }
  return E;
}

static double true_from_ecc_anomaly(double E, double e) {
  return 2.0 * atan2(sqrt(1.0 + e) * sin(0.5 * E), sqrt(1.0 - e) * cos(0.5 * E));
}

/* contract of the extern declared at fast_likelihood.pyx:27-30; phase convention
 * M(t) = 2 pi (t - t0) / P - phi0 (samples.py:228-229, thejoker.py:441);
 * rv = K (cos(f + omega) + e cos omega) (_keplerian_orbit.py:642-656). */
void orc_rv_from_elements(const double *t, double *rv, int N_t, double P, double K,
                          double e, double omega, double phi0, double t0, double tol,
                          int maxiter, int variant) {
  for (int n = 0; n < N_t; n++) {
    double M = 2.0 * M_PI * (t[n] - t0) / P - phi0;
    double E = orc_eccentric_anomaly(M, e, tol, maxiter, variant);
    double f = true_from_ecc_anomaly(E, e);
    rv[n] = K * (cos(omega + f) + e * cos(omega));
  }
}

/* ------------------------------------------------------------------------- */
/* Workspace = the per-object scratch CJokerHelper allocates (pyx:181-205)     */

typedef struct {
  int N, L;
  double *M_T;     /* [L,N] */
  double *s_ivar;  /* [N]  jitter-inflated inverse variance (pyx:48-67) */
  double *use_ivar;/* [N]  what the algebra reads: ivar (reference) or s_ivar */
  double *Lambda;  /* [L]  per-sample copy (Lambda[0] is rewritten, pyx:461-464) */
  double *B, *Binv, *Btmp; /* [N,N] */
  double *A, *Ainv, *Atmp; /* [L,L] */
  double *b;       /* [N] */
  double *a;       /* [L] */
  int *npar_ipiv, *ntime_ipiv;
  double *npar_work, *ntime_work;
} Work;

static Work *work_new(const OrcSpec *sp) {
  int N = sp->n_times, L = sp->n_linear;
  Work *w = (Work *)calloc(1, sizeof(Work));
  w->N = N; w->L = L;
  w->M_T = (double *)calloc((size_t)L * N, sizeof(double));
  w->s_ivar = (double *)calloc(N, sizeof(double));
  w->use_ivar = (double *)calloc(N, sizeof(double));
  w->Lambda = (double *)calloc(L, sizeof(double));
  w->B = (double *)calloc((size_t)N * N, sizeof(double));
  w->Binv = (double *)calloc((size_t)N * N, sizeof(double));
  w->Btmp = (double *)calloc((size_t)N * N, sizeof(double));
  w->A = (double *)calloc((size_t)L * L, sizeof(double));
  w->Ainv = (double *)calloc((size_t)L * L, sizeof(double));
  w->Atmp = (double *)calloc((size_t)L * L, sizeof(double));
  w->b = (double *)calloc(N, sizeof(double));
  w->a = (double *)calloc(L, sizeof(double));
  w->npar_ipiv = (int *)calloc(L, sizeof(int));
  w->ntime_ipiv = (int *)calloc(N, sizeof(int));
  w->npar_work = (double *)calloc(L > N ? L : N, sizeof(double));
  w->ntime_work = (double *)calloc(N, sizeof(double));
  /* rows 1.. of M_T are the transposed trend matrix (pyx:181-184) */
  for (int n = 0; n < N; n++)
    for (int i = 1; i < L; i++) w->M_T[i * N + n] = sp->trend_M[n * (L - 1) + (i - 1)];
  for (int i = 0; i < L; i++) w->Lambda[i] = sp->Lambda[i];
  return w;
}

static void work_free(Work *w) {
  free(w->M_T); free(w->s_ivar); free(w->use_ivar); free(w->Lambda);
  free(w->B); free(w->Binv); free(w->Btmp);
  free(w->A); free(w->Ainv); free(w->Atmp); free(w->b); free(w->a);
  free(w->npar_ipiv); free(w->ntime_ipiv); free(w->npar_work); free(w->ntime_work);
  free(w);
}

/* pyx:48-67 */
static void get_ivar(const double *ivar, double s, double *new_ivar, int N) {
  for (int i = 0; i < N; i++) new_ivar[i] = ivar[i] / (1 + s * s * ivar[i]);
}

/* pyx:255-297 */
static int make_AAinv(Work *w) {
  int L = w->L, N = w->N, info = 0, lwork = L;
  for (int i = 0; i < L; i++)
    for (int j = 0; j < L; j++) w->Ainv[i * L + j] = 0.;
  for (int i = 0; i < L; i++) {
    w->Ainv[i * L + i] = 1 / w->Lambda[i];
    for (int j = 0; j < L; j++) {
      for (int n = 0; n < N; n++)
        w->Ainv[i * L + j] += (w->M_T[j * N + n] * w->use_ivar[n] * w->M_T[i * N + n]);
      w->Atmp[i * L + j] = w->Ainv[i * L + j];
    }
  }
  CALL_DGETRF(&L, &L, w->Atmp, &L, w->npar_ipiv, &info);
  if (info != 0) return -1;
  CALL_DGETRI(&L, w->Atmp, &L, w->npar_ipiv, w->npar_work, &lwork, &info);
  if (info != 0) return -1;
  for (int i = 0; i < L; i++)
    for (int j = 0; j < L; j++) w->A[i * L + j] = w->Atmp[i * L + j];
  return 0;
}

/* pyx:299-357 */
static double make_bBBinv(Work *w, const OrcSpec *sp) {
  int L = w->L, N = w->N, info = 0;
  double log_det_val;
  for (int n = 0; n < N; n++) {
    w->b[n] = 0.;
    for (int i = 0; i < L; i++) w->b[n] += w->M_T[i * N + n] * sp->mu[i];
    for (int m = 0; m < N; m++) w->B[n * N + m] = 0.;
  }
  for (int n = 0; n < N; n++) {
    w->B[n * N + n] = 1 / w->use_ivar[n];
    for (int m = 0; m < N; m++) {
      w->Binv[n * N + m] = 0.;
      for (int i = 0; i < L; i++)
        w->B[n * N + m] += (w->M_T[i * N + n] * w->Lambda[i] * w->M_T[i * N + m]);
      w->Btmp[n * N + m] = w->B[n * N + m];
    }
  }
  for (int n = 0; n < N; n++) {
    w->Binv[n * N + n] = w->use_ivar[n];
    for (int i = 0; i < L; i++)
      for (int m = 0; m < N; m++)
        for (int j = 0; j < L; j++)
          w->Binv[n * N + m] -= (w->use_ivar[n] * w->M_T[i * N + n] * w->A[i * L + j] *
                                 w->M_T[j * N + m] * w->use_ivar[m]);
  }
  CALL_DGETRF(&N, &N, w->Btmp, &N, w->ntime_ipiv, &info);
  if (info != 0) return INFINITY;
  log_det_val = 0.;
  for (int i = 0; i < N; i++) log_det_val += log(2 * M_PI * fabs(w->Btmp[i * N + i]));
  return log_det_val;
}

/* pyx:359-425 */
static double likelihood_worker(Work *w, const OrcSpec *sp, int make_aAinv) {
  int L = w->L, N = w->N, info = 0, nrhs = 1, lwork = N;
  char uplo = 'U';
  double chi2, log_det_val;
  if (make_AAinv(w) < 0) return INFINITY;
  log_det_val = make_bBBinv(w, sp);
  chi2 = 0.;
  for (int n = 0; n < N; n++)
    for (int m = 0; m < N; m++)
      chi2 += ((w->b[m] - sp->rv[m]) * w->Binv[n * N + m] * (w->b[n] - sp->rv[n]));
  if (make_aAinv == 1) {
    for (int i = 0; i < L; i++) w->a[i] = 0.;
    for (int n = 0; n < N; n++)
      for (int i = 0; i < L; i++) w->a[i] += w->M_T[i * N + n] * w->use_ivar[n] * sp->rv[n];
    for (int i = 0; i < L; i++) w->a[i] += sp->mu[i] / w->Lambda[i];
    for (int i = 0; i < L; i++)
      for (int j = 0; j < L; j++) w->Atmp[i * L + j] = w->Ainv[i * L + j];
    CALL_DSYSV(&uplo, &L, &nrhs, w->Atmp, &L, w->npar_ipiv, w->a, &L, w->npar_work, &lwork,
               &info);
    if (info != 0) return INFINITY;
  }
  return -0.5 * (chi2 + log_det_val);
}

/* per-sample preamble shared by pyx:445-467, 503-524, 555-574.
 * clamp: pyx:464 clamps Lambda[0] at max_K^2 in batch_marginal_ln_likelihood only. */
static void prepare_sample(Work *w, const OrcSpec *sp, const double *row, int clamp) {
  double P = row[0], e = row[1], om = row[2], M0 = row[3];
  orc_rv_from_elements(sp->t, w->M_T, w->N, P, 1., e, om, M0, sp->t0, sp->kepler_tol,
                       sp->kepler_maxiter, sp->kepler_variant);
  get_ivar(sp->ivar, row[4], w->s_ivar, w->N);
  /* jitter_mode 0 == the reference as written: s_ivar is a dead store and the
   * algebra reads self.ivar (pyx:274, 317, 333-339, 401).  jitter_mode 1 == the
   * intended semantics (src/tests/py_likelihood.py:17-29, 168-170). */
  memcpy(w->use_ivar, sp->jitter_mode ? w->s_ivar : sp->ivar, sizeof(double) * w->N);
  if (sp->K_prior_kind == 0) {
    /* Cython 3 (unpinned in the reference's pyproject.toml:7) types `double ** (-2/3.)` as
     * possibly complex and emits C99 cpow on (x + 0i); glibc's cpow and pow differ in the
     * last ulp for some x.  Following the generated code keeps Lambda[0] bit-identical to
     * the compiled reference (oracle/_ref). */
    w->Lambda[0] = (sp->sigma_K0 * sp->sigma_K0 / (1 - e * e)) *
                   creal(cpow(P / sp->P0 + 0.0 * I, -2 / 3. + 0.0 * I));
    if (clamp) w->Lambda[0] = fmin(sp->max_K * sp->max_K, w->Lambda[0]);
  }
}

/* pyx:428-469 */
int orc_batch_marginal_ln_likelihood(const OrcSpec *sp, const double *chunk, long n_samples,
                                     double *ll, int n_threads) {
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#pragma omp parallel
#endif
  {
    Work *w = work_new(sp);
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
    for (long n = 0; n < n_samples; n++) {
      prepare_sample(w, sp, chunk + 5 * n, 1);
      ll[n] = likelihood_worker(w, sp, 0);
    }
    work_free(w);
  }
  return 0;
}

/* pyx:547-576 (test_likelihood_worker) and the per-sample body of
 * batch_get_posterior_samples (pyx:503-530).  Exposes a, A, b, B, Binv, Ainv.
 * clamp_override: -1 = reference behaviour for this entry point (no clamp),
 * 0/1 = force. */
double orc_likelihood_worker_full(const OrcSpec *sp, const double *row, int clamp_override,
                                  double *a, double *A, double *Ainv, double *b, double *B,
                                  double *Binv) {
  Work *w = work_new(sp);
  int clamp = clamp_override < 0 ? 0 : clamp_override;
  prepare_sample(w, sp, row, clamp);
  double ll = likelihood_worker(w, sp, 1);
  int N = w->N, L = w->L;
  if (a) memcpy(a, w->a, sizeof(double) * L);
  if (A) memcpy(A, w->A, sizeof(double) * L * L);
  if (Ainv) memcpy(Ainv, w->Ainv, sizeof(double) * L * L);
  if (b) memcpy(b, w->b, sizeof(double) * N);
  if (B) memcpy(B, w->B, sizeof(double) * N * N);
  if (Binv) memcpy(Binv, w->Binv, sizeof(double) * N * N);
  work_free(w);
  return ll;
}

/* pyx:471-545 without the Python-level multivariate_normal call: returns, per
 * input row, ll, a[L] and Ainv[L,L]; the caller (oracle/oracle.py) draws
 * rng.multivariate_normal(a, inv(Ainv)) exactly as pyx:529-530 does. */
int orc_batch_posterior_aAinv(const OrcSpec *sp, const double *chunk, long n_samples,
                              int clamp_override, double *ll, double *a, double *Ainv) {
  Work *w = work_new(sp);
  int L = w->L;
  int clamp = clamp_override < 0 ? 0 : clamp_override;
  for (long n = 0; n < n_samples; n++) {
    prepare_sample(w, sp, chunk + 5 * n, clamp);
    ll[n] = likelihood_worker(w, sp, 1);
    memcpy(a + n * L, w->a, sizeof(double) * L);
    memcpy(Ainv + n * L * L, w->Ainv, sizeof(double) * L * L);
  }
  work_free(w);
  return 0;
}

/* the unit-amplitude RV column for one sample (row 0 of M_T), for tests */
void orc_design_column(const OrcSpec *sp, const double *row, double *z) {
  orc_rv_from_elements(sp->t, z, sp->n_times, row[0], 1., row[1], row[2], row[3], sp->t0,
                       sp->kepler_tol, sp->kepler_maxiter, sp->kepler_variant);
}

int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
