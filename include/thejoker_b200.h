/*
 * thejoker_b200.h -- C ABI of libthejoker_b200.so, the sm_100a implementation of
 * The Joker's hot path (prior stream -> Kepler solve per epoch -> design matrix ->
 * Gaussian-marginal log-likelihood -> rejection accept -> linear-parameter draw).
 *
 * The entry points are what a binding of the reference's operator boundary,
 * `cdef class CJokerHelper` (thejoker/src/fast_likelihood.pyx:70-576), needs:
 * plain pointers and sizes, no torch / numpy / Python types.  All functions
 * return 0 on success and a negative TJB_E_* code on failure; the message of the
 * last failure on the calling thread is available from tjb_last_error().
 *
 * Pointer arguments named d_* are DEVICE pointers on the handle's GPU (the host
 * layer allocates them as torch tensors and passes data_ptr()); h_* are HOST
 * pointers.  All work is enqueued on the stream set with tjb_set_stream()
 * (default: the legacy default stream).  Functions taking only device pointers are
 * asynchronous with respect to the host; functions with h_* outputs synchronise
 * the stream before returning.  A handle may be used by one host thread at a time.
 *
 * There is no CPU path: every entry point fails with TJB_E_CUDA when no sm_100
 * device is usable.
 */
#ifndef THEJOKER_B200_H
#define THEJOKER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TJB_VERSION 100
#define TJB_MAX_LINEAR 8

enum {
  TJB_OK = 0,
  TJB_E_INVALID = -1, /* bad argument / unsupported shape */
  TJB_E_CUDA = -2,    /* CUDA runtime error, or no usable device */
  TJB_E_NOMEM = -3
};

/* What CJokerHelper.__init__ extracts from (data, prior, trend_M)
 * (fast_likelihood.pyx:125-253).  All arrays are HOST pointers, copied by
 * tjb_create.  Linear-parameter order: [K, v0, dv0_1.., v1, v2, ..]
 * (pyx:143-148, 204-252). */
typedef struct TjbSpec {
  int32_t n_times;       /* N = len(data) (pyx:155) */
  int32_t n_linear;      /* L = 1 + poly_trend + n_offsets (pyx:158), 1..TJB_MAX_LINEAR */
  double t_ref;          /* data._t_ref_bmjd (pyx:162) */
  const double *t;       /* [N] BMJD (pyx:163) */
  const double *rv;      /* [N] (pyx:164) */
  const double *ivar;    /* [N] inverse variance, 1/rv_unit^2 (pyx:165-166) */
  const double *trend_M; /* [N, L-1] row-major (pyx:167, 174-184) */
  double mu[TJB_MAX_LINEAR];     /* prior means (pyx:204, 219, 243-252) */
  double Lambda[TJB_MAX_LINEAR]; /* prior variances (pyx:205, 220, 246-251); [0] unused if K_prior_kind==0 */
  int32_t K_prior_kind;  /* 0 = FixedCompanionMass (pyx:225-226), 1 = Normal */
  double sigma_K0;       /* pyx:239 */
  double P0;             /* pyx:240-241, in days */
  double max_K;          /* pyx:242 */
  int32_t jitter_mode;   /* 0 = as the reference is written: s ignored (pyx:458 is a dead store);
                            1 = s enters the covariance, ivar/(1 + s^2 ivar) (pyx:48-67) */
} TjbSpec;

typedef struct TjbHandle TjbHandle;

/* ---- lifecycle -------------------------------------------------------- */
/* replaces CJokerHelper.__init__ (pyx:125-253) */
int tjb_create(const TjbSpec *spec, int device, TjbHandle **out);
/* Replace the star (data + linear prior) of an existing handle; device buffers are
 * reused.  For drivers that run many stars against one shared prior cache (the
 * reference re-creates a CJokerHelper per star, thejoker.py:87-91). */
int tjb_update_star(TjbHandle *h, const TjbSpec *spec);
void tjb_destroy(TjbHandle *h);
int tjb_set_stream(TjbHandle *h, void *cuda_stream);
const char *tjb_last_error(void);
int tjb_version(void);
/* number of SMs and the kernel's resident CTAs per SM on the handle's device */
int tjb_device_info(TjbHandle *h, int *n_sm, int *ctas_per_sm, int *cc_major, int *cc_minor);

/* ---- batch_marginal_ln_likelihood (pyx:428-469) --------------------- */
/* Prior columns as separate device arrays (SoA).  d_s may be NULL: then every
 * sample has jitter s_const.  d_ll[n] receives the log-likelihoods.  If
 * d_llmax_key is not NULL the int64 it points to is max-updated with the
 * order-preserving key of every ll written (see tjb_key_to_double); initialise it
 * with tjb_llmax_reset.  Keys are plain int64, so shards on several GPUs combine
 * with an integer MAX all-reduce. */
int tjb_marginal_ll_soa(TjbHandle *h, const double *d_P, const double *d_e,
                        const double *d_omega, const double *d_M0, const double *d_s,
                        double s_const, int64_t n, double *d_ll, int64_t *d_llmax_key);
/* Fused max exchange for one process driving several GPUs: besides the key passed per
 * call, every likelihood launch of this handle also max-updates the n given keys, which
 * live on the listed peer devices (NVLink peer access is enabled here).  After all
 * shards' kernels have completed, each GPU's key holds the max over all shards -- the
 * master-side ``lls.max()`` of multiproc_helpers.py:256-258 without a collective call.
 * n = 0 clears the list. */
int tjb_set_peer_keys(TjbHandle *h, int64_t *const *d_peer_keys, const int *peer_devices, int n);
/* Prior rows packed as the reference packs them: d_chunk[n,5] row-major
 * [P, e, omega, M0, s] (pyx:41, 448-451).  uniform_s != 0 promises that column 4
 * is the same for all rows (the host layer checks), enabling the constant-jitter
 * kernel. */
int tjb_marginal_ll_aos(TjbHandle *h, const double *d_chunk, int uniform_s, int64_t n,
                        double *d_ll, int64_t *d_llmax_key);
/* Literal drop-in for batch_marginal_ln_likelihood on HOST buffers: copies the
 * chunk to the device in slices, runs the kernel and copies ll back, overlapping
 * the three on two streams.  Pinned host buffers give full PCIe bandwidth. */
int tjb_marginal_ll_host(TjbHandle *h, const double *h_chunk, int64_t n, double *h_ll);
/* The same for prior samples held as separate HOST columns (what a JokerSamples is:
 * thejoker.py:129-134 packs them into the (n, 5) chunk first; here they are sent as
 * they are).  h_s may be NULL: every sample then has jitter s_const and only 32 B per
 * sample cross PCIe. */
int tjb_marginal_ll_host_soa(TjbHandle *h, const double *h_P, const double *h_e,
                             const double *h_omega, const double *h_M0, const double *h_s,
                             double s_const, int64_t n, double *h_ll);
/* The same host columns, but the ll values stay on the device (d_ll[n]) with their running
 * max in *d_llmax_key (may be NULL), ready for tjb_accept: the prior cache is streamed
 * through the GPU slice by slice and never becomes resident.  This is how
 * rejection_sample_inmem / iterative_rejection_inmem (likelihood_helpers.py:91-229) and
 * the read_batch loop of the pool workers (multiproc_helpers.py:63-98) are served when the
 * prior samples live in host memory or in a memory-mapped cache file.  Pageable host
 * memory is staged through a page-locked ring by host threads (both host entry points). */
int tjb_marginal_ll_host_soa_resident(TjbHandle *h, const double *h_P, const double *h_e,
                                      const double *h_omega, const double *h_M0,
                                      const double *h_s, double s_const, int64_t n, double *d_ll,
                                      int64_t *d_llmax_key);

/* ---- drawn priors: JokerPrior.sample for the nonlinear parameters (prior.py:297-407) ----
 * rejection_sample(data, <int>) draws its prior samples first (thejoker.py:215-218).  Here
 * sample i of the prior is a pure function of (seed, i): a Philox4x32-10 stream per sample,
 * consumed in the order P, e, omega, M0, s.  The likelihood kernel can therefore generate
 * the samples in registers (tjb_marginal_ll_generated: the prior never exists in HBM), the
 * accepted rows are re-generated from their indices (tjb_prior_rows), and any sharding of
 * the index range over GPUs / ranks yields the same prior.  Distribution kinds: the
 * families of the reference's default prior and its documented variants --
 * UniformLog (distributions.py:17-51), Beta (Kipping13*, distributions.py:155-176),
 * uniform angles (prior.py:437, 469-472), constant / LogNormal / Normal jitter. */
enum {
  TJB_PRIOR_CONSTANT = 0,   /* p0 */
  TJB_PRIOR_UNIFORM = 1,    /* U(p0, p1) */
  TJB_PRIOR_UNIFORMLOG = 2, /* exp(U(ln p0, ln p1)) */
  TJB_PRIOR_BETA = 3,       /* Beta(p0, p1) */
  TJB_PRIOR_LOGNORMAL = 4,  /* exp(N(p0, p1)) */
  TJB_PRIOR_NORMAL = 5      /* N(p0, p1) */
};
typedef struct TjbPriorDist {
  int32_t kind;
  int32_t reserved;
  double p0, p1;
  double scale; /* multiplies the draw: unit conversion into [day, -, rad, rad, rv unit] */
} TjbPriorDist;
typedef struct TjbPriorGen {
  TjbPriorDist par[5]; /* P, e, omega, M0, s */
  uint64_t seed;
} TjbPriorGen;
/* columns of the samples with global indices [index0, index0 + n) into device arrays on
 * `device` (any may be NULL); asynchronous on `cuda_stream`; needs no handle */
int tjb_prior_sample(int device, void *cuda_stream, const TjbPriorGen *gen, int64_t index0,
                     int64_t n, double *d_P, double *d_e, double *d_omega, double *d_M0,
                     double *d_s);
/* packed rows h_rows[k, 5] = [P, e, omega, M0, s] of the samples with the given global
 * indices (host in, host out) */
int tjb_prior_rows(TjbHandle *h, const TjbPriorGen *gen, const int64_t *h_idx, int64_t k,
                   double *h_rows);
/* batch_marginal_ln_likelihood (pyx:428-469) over the generated samples
 * [index0, index0 + n): as tjb_marginal_ll_soa, without any prior array */
int tjb_marginal_ll_generated(TjbHandle *h, const TjbPriorGen *gen, int64_t index0, int64_t n,
                              double *d_ll, int64_t *d_llmax_key);

/* ---- accept step (likelihood_helpers.py:107-109; multiproc_helpers.py:256-258) */
int tjb_llmax_reset(TjbHandle *h, int64_t *d_llmax_key);
/* max-update *d_llmax_key with d_ll[0..n) (for ll arrays not produced above) */
int tjb_llmax_update(TjbHandle *h, const double *d_ll, int64_t n, int64_t *d_llmax_key);
int tjb_llmax_get(TjbHandle *h, const int64_t *d_llmax_key, double *h_max);
double tjb_key_to_double(int64_t key);
int64_t tjb_double_to_key(double x);

/* good = where(exp(ll - max) > u)[0][:max_keep], ascending.  Exactly one of
 * d_uniforms / pcg != NULL.  With pcg, u[i] is the (pcg_offset + i)-th double a
 * numpy Generator(PCG64) in state (state, inc) would return from .random(), bit
 * for bit (so rng.uniform(size=n) never has to exist on the host).  index_base is
 * added to the written indices (shard offset).  Outputs: d_idx[max_keep] device;
 * h_counts[0] = number accepted in [0,n) before truncation, h_counts[1] = number
 * written, h_counts[2] = number of samples with |exp(ll-max) - u| <= near_tol. */
typedef struct TjbPcg64 {
  uint64_t state_hi, state_lo, inc_hi, inc_lo;
} TjbPcg64;
int tjb_accept(TjbHandle *h, const double *d_ll, int64_t n, const int64_t *d_llmax_key,
               const double *d_uniforms, const TjbPcg64 *pcg, int64_t pcg_offset,
               int64_t index_base, int64_t max_keep, double near_tol, int64_t *d_idx,
               int64_t *h_counts);
/* Number of non-finite lls (NaN, +-inf) the last tjb_accept / tjb_accept_dist call of this
 * handle saw in its range (summed over the ranks for tjb_accept_dist): what
 * iterative_rejection_inmem tests with np.isfinite over all lls before it accepts
 * (likelihood_helpers.py:173-176).  No synchronisation: the value was fetched with the
 * counts. */
int tjb_accept_nonfinite(TjbHandle *h, int64_t *h_count);

/* ---- the same accept step over shards that live in different processes (one rank per
 * GPU): NCCL inside the library, no torch / MPI types in the interface.  Replaces the
 * master-side gather + max + compare + where of multiproc_helpers.py:256-263, 373-381.
 * NCCL is bound at run time (dlopen of libnccl.so.2; inside a PyTorch process that is
 * torch's own copy); without it these entry points fail with TJB_E_CUDA and everything
 * else works.  Bootstrap: rank 0 calls tjb_comm_unique_id and distributes the
 * TJB_COMM_ID_BYTES it gets by any host-side means (a file, MPI_Bcast, a TCP store);
 * every rank then calls tjb_comm_create (collective: ncclCommInitRank). */
#define TJB_COMM_ID_BYTES 128
typedef struct TjbComm TjbComm;
int tjb_comm_unique_id(void *out_id);
int tjb_comm_create(const void *id_bytes, int n_ranks, int rank, int device, TjbComm **out);
void tjb_comm_destroy(TjbComm *comm);
/* integer MAX all-reduce of *d_llmax_key over the ranks, on the handle's stream */
int tjb_comm_allreduce_max_key(TjbHandle *h, TjbComm *comm, int64_t *d_llmax_key);
/* Collective accept.  This rank owns the global samples [global_offset, global_offset +
 * n_local) with lls d_ll and the running max of its shard in *d_llmax_key, which is
 * replaced by the max over all ranks (ncclAllReduce MAX on the int64 key).  Uniforms: as
 * tjb_accept, addressed globally (u of global sample g is the g-th double of the PCG64
 * stream; d_uniforms, if given, is this rank's slice).  Every rank receives in
 * d_idx[max_keep] the same ascending global indices -- the rank-ordered concatenation
 * truncated to max_keep -- and in h_counts the global [accepted, written, near].  Two
 * small all-gathers (counts, then indices); synchronises the stream. */
int tjb_accept_dist(TjbHandle *h, TjbComm *comm, const double *d_ll, int64_t n_local,
                    int64_t *d_llmax_key, const double *d_uniforms, const TjbPcg64 *pcg,
                    int64_t global_offset, int64_t max_keep, double near_tol, int64_t *d_idx,
                    int64_t *h_counts);

/* the uniforms themselves (tests; parity with numpy) */
int tjb_pcg64_uniform(TjbHandle *h, const TjbPcg64 *pcg, int64_t offset, int64_t n,
                      double *d_out);

/* ---- batch_get_posterior_samples / test_likelihood_worker (pyx:471-576) */
/* For k rows h_rows[k,5]: ll, posterior mean a[k,L] and covariance A[k,L,L] of the
 * linear parameters (pyx:394-423, 530).  clamp_K != 0 applies the max_K clamp to
 * Lambda_K; the reference does not in these two entry points (pyx:519-521,
 * 569-571). Any output pointer may be NULL. */
int tjb_posterior_aA(TjbHandle *h, const double *h_rows, int64_t k, int clamp_K,
                     double *h_ll, double *h_a, double *h_A);
/* Draw n_per linear-parameter vectors per row: x = a + chol(A) z with
 * h_normals[k, n_per, L] standard normals supplied by the host RNG, and pack
 * rows [P, e, omega, M0, s, x...] (pyx:532-542).  h_out[k*n_per, 5+L]. */
int tjb_posterior_draw(TjbHandle *h, const double *h_rows, int64_t k, int n_per, int clamp_K,
                       const double *h_normals, double *h_out, double *h_ll);
/* row 0 of M_T for one sample (pyx:453-455): z[N]; h_stats[3] (optional) returns
 * extra FP32 steps, extra FP64 steps, non-converged epochs of the solver. */
int tjb_design_column(TjbHandle *h, const double *h_row, double *h_z, int32_t *h_stats);
/* ln of the un-marginalised likelihood of k full posterior samples
 * (JokerSamples.ln_unmarginalized_likelihood, thejoker/samples.py:611-632): h_rows[k, 5+L]
 * row-major [P, e, omega, M0, s, K, v0, (offsets), v1, ...] -- the layout
 * batch_get_posterior_samples returns (pyx:531-542) -- in internal units;
 * h_ll[k] = sum_n ln N(rv_n | M_n . x, 1/ivar_n + s^2).  The jitter always enters here,
 * whatever jitter_mode says (the reference's samples.py does apply it). */
int tjb_unmarginalized_ll(TjbHandle *h, const double *h_rows, int64_t k, double *h_ll);

/* ---- many stars against one shared, device-resident prior cache ------------------
 * The reference has no batched entry point: users loop over stars and every
 * TheJoker.rejection_sample call re-creates the CJokerHelper and re-reads the prior
 * cache (thejoker/thejoker.py:87-91, 213-257; rejection_sample_inmem,
 * likelihood_helpers.py:91-127).  This runs that loop natively -- per star: point a
 * handle at the star (tjb_update_star), marginal ll over the whole prior, accept with the
 * star's own PCG64 stream, gather the accepted rows from the device columns, draw the
 * linear parameters (batch_get_posterior_samples, pyx:471-545) -- on n_slots host
 * threads, each with its own handle, CUDA stream and ll buffer, so the host side of one
 * star overlaps the kernels of the others.  Stars are independent, so the results do not
 * depend on n_slots or on which slot ran a star. */
typedef struct TjbMultiStarJob {
  int64_t n_stars;
  const TjbSpec *specs;     /* [n_stars], all with the same n_linear L */
  const TjbPcg64 *pcg;      /* [n_stars] generator state of each star's uniforms (see tjb_accept) */
  /* shared prior cache: SoA device columns on `device`; d_s may be NULL (then s_const) */
  const double *d_P, *d_e, *d_omega, *d_M0, *d_s;
  double s_const;
  int64_t n_prior;
  int64_t max_keep;         /* accepted samples kept per star (max_posterior_samples) */
  double near_tol;          /* as tjb_accept */
  int32_t n_per;            /* linear-parameter draws per accepted sample; 0 = indices only */
  int32_t clamp_K;          /* as tjb_posterior_draw */
  int32_t n_slots;          /* stars in flight (host threads), 1..64 */
  int32_t reserved;
  const double *h_normals;  /* [n_stars, max_keep, n_per, L] standard normals (n_per > 0) */
  /* outputs, HOST; star i writes only the first h_counts[3i+1] entries of its slices */
  int64_t *h_idx;           /* [n_stars, max_keep] accepted prior indices, ascending */
  int64_t *h_counts;        /* [n_stars, 3] as tjb_accept */
  double *h_llmax;          /* [n_stars] max ll over the prior */
  double *h_rows;           /* [n_stars, max_keep * n_per, 5 + L] packed samples (n_per > 0) */
  double *h_ll;             /* [n_stars, max_keep] ll of the accepted samples (n_per > 0; may be NULL) */
} TjbMultiStarJob;
int tjb_multistar_rejection(int device, const TjbMultiStarJob *job);

/* Solver statistics accumulated by every likelihood launch of this handle since the
 * last reset: h_stats[0] = extra FP64 Householder passes (lane-epochs that needed more
 * than one; high eccentricity near pericentre), h_stats[1] = epochs that did not
 * converge within the iteration cap (their ll is still finite, counted for reporting),
 * h_stats[2..3] reserved.  Synchronises the handle's streams. */
int tjb_get_stats(TjbHandle *h, uint64_t *h_stats, int reset);

/* Where the likelihood kernel reads the star's epoch rows from (process-wide).  0 (default):
 * from the kernel's parameter block when they fit (24 KB: 768 epochs at n_linear = 2, 512
 * at 3..4) -- uniform loads, uniform-register operands --, else staged in shared memory;
 * 1: always staged in shared memory.  Both kernels give bit-identical results
 * (tests/test_gpu_parity.py::test_epoch_rows_kernels_agree); the switch exists for that
 * test and for timing comparisons.  Initial value: environment TJB_FORCE_SHARED_ROWS. */
int tjb_set_epoch_rows_mode(int mode);

/* ---- measurement helper: FP64 FMA-chain peak of the device, TFLOP/s ---- */
int tjb_fp64_peak(TjbHandle *h, int iters, double *h_tflops, double *h_ms);

#ifdef __cplusplus
}
#endif
#endif /* THEJOKER_B200_H */
