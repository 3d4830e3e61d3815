/* joker_oracle.h -- TEST INFRASTRUCTURE ONLY (see joker_oracle.c header). */
#ifndef JOKER_ORACLE_H
#define JOKER_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

/* What CJokerHelper.__init__ extracts from (data, prior, trend_M)
 * (thejoker/src/fast_likelihood.pyx:125-253). */
typedef struct {
  int n_times;          /* N */
  int n_linear;         /* L = 1 + poly_trend + n_offsets (pyx:158) */
  double t0;            /* data._t_ref_bmjd (pyx:162) */
  const double *t;      /* [N] BMJD (pyx:163) */
  const double *rv;     /* [N] (pyx:164) */
  const double *ivar;   /* [N] 1/rv_err^2 (pyx:165-166) */
  const double *trend_M;/* [N, L-1] row-major (pyx:167, 174-184) */
  const double *mu;     /* [L] prior means of [K, v0, dv0_*, v1, ...] (pyx:204-252) */
  const double *Lambda; /* [L] prior variances, same order; Lambda[0] unused when K_prior_kind==0 */
  int K_prior_kind;     /* 0 = FixedCompanionMass (pyx:225-226), 1 = plain Normal */
  double sigma_K0, P0, max_K; /* pyx:239-242 */
  int jitter_mode;      /* 0 = reference as written (s ignored), 1 = s enters the covariance */
  double kepler_tol;    /* pyx:35 */
  int kepler_maxiter;   /* pyx:36 */
  int kepler_variant;   /* 0 update-then-test (default), 1 test-then-update */
} OrcSpec;

void orc_set_lapack(void *dgetrf, void *dgetri, void *dsysv);
int orc_has_lapack(void);
int orc_max_threads(void);

double orc_eccentric_anomaly(double M, double e, double tol, int maxiter, int variant);
void orc_rv_from_elements(const double *t, double *rv, int N_t, double P, double K, double e,
                          double omega, double phi0, double t0, double tol, int maxiter,
                          int variant);
void orc_design_column(const OrcSpec *sp, const double *row, double *z);

int orc_batch_marginal_ln_likelihood(const OrcSpec *sp, const double *chunk, long n_samples,
                                     double *ll, int n_threads);
double orc_likelihood_worker_full(const OrcSpec *sp, const double *row, int clamp_override,
                                  double *a, double *A, double *Ainv, double *b, double *B,
                                  double *Binv);
int orc_batch_posterior_aAinv(const OrcSpec *sp, const double *chunk, long n_samples,
                              int clamp_override, double *ll, double *a, double *Ainv);

/* joker_truth.c: quad-precision evaluation of the same model (residual form) */
int orc_truth_marginal_ln_likelihood(const OrcSpec *sp, const double *chunk, long n_samples,
                                     double *ll, double *kappa, int n_threads);
int orc_truth_posterior_aA(const OrcSpec *sp, const double *row, int clamp, double *a,
                           double *A);

#ifdef __cplusplus
}
#endif
#endif
