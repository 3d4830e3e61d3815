// marginal_ll.cuh -- the hot kernel: one thread per prior sample, one CTA per SM; the
// star's epoch rows come in through the kernel's parameter block (uniform loads; shared or
// global memory for long tables), the trig tables are staged in shared memory, the Gram
// sums live in registers.
//
// Replaces CJokerHelper.batch_marginal_ln_likelihood
// (thejoker/src/fast_likelihood.pyx:428-469) and everything it calls per sample
// (c_rv_from_elements pyx:453-455, get_ivar pyx:458, Lambda_K pyx:461-464,
// likelihood_worker pyx:359-425).
#pragma once

#include "linalg.cuh"
#include "prior_gen.cuh"

namespace tjb {

// Where the prior samples live.  Either SoA columns (coalesced 8-byte loads, the
// native layout) or the reference's packed AoS rows [P, e, omega, M0, s].
struct PriorView {
  const double *P, *e, *omega, *M0, *s;  // SoA (s may be null)
  const double *aos;                     // AoS rows of 5, or null
  // constant-jitter kernel on AoS rows: rows whose s differs from s_expect raise
  // *nonuniform (optimistic execution; the caller then reruns per-sample jitter)
  double s_expect;
  int *nonuniform;
};

// The third source of prior samples: none in memory at all.  Sample i of a launch is
// generated in registers from (gen.seed, index0 + i) by the counter-based sampler of
// prior_gen.cuh (kGen kernels), so a prior that is drawn -- rejection_sample(data, <int>),
// thejoker.py:215-218 -- costs no HBM traffic and no HBM capacity.
struct PriorGenView {
  PriorGenSpec gen;
  long long index0;
};

// Per-star constants.  Passed by value as a kernel parameter (constant bank).
//
// Epoch table row n (stride row_stride(L) doubles):
//   constant-jitter kernel (kJit=false): [dt_n, w_n, w_n y_n, w_n T_n1 .. w_n T_n,L-1]
//        with w_n = ivar_n / (1 + s^2 ivar_n) for the one s of the call
//   per-sample-jitter kernel (kJit=true): [dt_n, var_n, y_n, T_n1 .. T_n,L-1]
//        with var_n = 1 / ivar_n
// T = trend_M (pyx:167): T_1 is the constant column, then offsets, then dt^k.
// y is centred by the host (y - c, mu_v0 - c) when column T_1 is all ones, which
// leaves the marginal likelihood unchanged and removes most of the cancellation
// in chi2.
TJB_HD constexpr int row_stride(int L) { return (L + 2 + 1) & ~1; }

// Shape of the likelihood kernel: threads of the one CTA per SM, epochs per loop iteration
// (independent Kepler chains for ILP), and whether z takes its reciprocal from the step
// (kepler.cuh, TJB_XZ).  Timed on B200 (profiles/r02c_tune_cta_shapes.jsonl):
//   * per-sample jitter (all the Gram sums in registers): 640 threads x 4 epochs, 96 registers;
//     more threads lose to spills (1024 x 2: -7 % at L = 2, -35 % at L = 4);
//   * constant jitter, L <= 4: 1024 threads x 2 epochs, 64 registers:
//     +2.7 % / +2.9 % / +3.4 % at L = 2 / 3 / 4 over 640 x 4 -- the shorter FP32 / FP64
//     phases of more warps interleave better on the four units the kernel loads;
//   * constant jitter, L > 4: the 640 x 4 shape (not timed wider).
// A shape whose warp count is not a multiple of 4 leaves schedulers unevenly loaded.
#ifndef TJB_EPOCHS_PER_ITER
#define TJB_EPOCHS_PER_ITER 4
#endif
#ifndef TJB_LL_THREADS
#define TJB_LL_THREADS 640
#endif
#ifndef TJB_LL_MIN_CTAS
#define TJB_LL_MIN_CTAS 1
#endif
#ifndef TJB_WIDE_THREADS  // 0: one shape for every kernel
#define TJB_WIDE_THREADS 1024
#endif
#ifndef TJB_WIDE_EPOCHS
#define TJB_WIDE_EPOCHS 2
#endif
template <int L, bool kJit>
struct LLShape {
  static constexpr bool kWide = TJB_WIDE_THREADS > 0 && !kJit && L <= 4;
  static constexpr int kThreads = kWide ? TJB_WIDE_THREADS : TJB_LL_THREADS;
  static constexpr int kEpochs = kWide ? TJB_WIDE_EPOCHS : TJB_EPOCHS_PER_ITER;
  static constexpr bool kXZ = TJB_XZ != 0;
};

// header of the loop over groups of kEpochsPerIter epochs; leaves n at the first epoch of
// the remainder.  The TJB_TRIM form counts groups down (one add + compare against zero per
// iteration instead of re-loading N and comparing n + kEpochsPerIter against it).
#if TJB_TRIM
#define TJB_EPOCH_GROUPS(n, N, row, RS)                                                  \
  n = (N / kEpochsPerIter) * kEpochsPerIter;                                            \
  for (int grp_ = N / kEpochsPerIter; grp_ > 0; --grp_, row += kEpochsPerIter * RS)
#else
#define TJB_EPOCH_GROUPS(n, N, row, RS) \
  for (; n + kEpochsPerIter <= N; n += kEpochsPerIter, row += kEpochsPerIter * RS)
#endif

struct StarParams {
  int n_times;
  const double *table;              // device, [N, row_stride(L)]
  double Gc[kTri<kMaxLinear>];      // constant entries of Ainv, packed upper (kJit=false only)
  double hc[kMaxLinear];            // constant entries of h (kJit=false), mu_i/Lambda_i (kJit=true)
  double inv_Lambda[kMaxLinear];    // 1/Lambda_i, i >= 1 ([0] for a Normal K prior)
  double quad0;                     // y^T C^-1 y + sum_{i>=1} mu_i^2/Lambda_i  (kJit: only the mu part)
  double c0;                        // N log 2pi - sum log w + sum_{i>=1} log Lambda_i (kJit: no w part)
  double mu_K;                      // prior mean of K (pyx:243)
  double Lambda_K;                  // prior variance of K when K_prior_kind == 1
  int K_prior_kind;                 // 0 FixedCompanionMass, 1 Normal
  double sigma_K0_sq, inv_P0, max_K_sq;
  int apply_jitter;                 // kJit kernels: 0 -> treat s as 0 (reference behaviour)
  double zero;                      // run-time 0.0, see TrigCoef::load
  const SinCos *trig_table;         // device, kTrigTableSize nodes (null for the polynomial back-end)
  unsigned long long *stats;        // device counters of the solver's rare path (may be null)
};

// Where the likelihood kernel publishes its running max.  keys[0..n) are int64 max-keys
// on this GPU and, through NVLink peer mappings, on the other GPUs that hold shards of
// the same prior cache: every CTA max-updates all of them with system-scope atomics, so
// when all shards' kernels have finished each GPU's own key already holds the global
// max -- the max "all-reduce" of the accept step is fused into the kernel's epilogue and
// no separate collective or host round trip is needed (one process driving several
// GPUs; one rank per GPU goes through NCCL instead, sharding.py).
constexpr int kMaxPeers = 16;
struct MaxKeys {
  long long *keys[kMaxPeers];
  int n;
};

// order-preserving int64 key of a double: key(a) < key(b) <=> a < b, NaN above +inf
// (so a max-reduction propagates NaN the way numpy.max does).
TJB_HD long long ll_to_key(double x) {
  long long b;
#if defined(__CUDA_ARCH__)
  b = __double_as_longlong(x);
#else
  memcpy(&b, &x, 8);
#endif
  if (x != x) b = 0x7FF8000000000000LL;
  return b >= 0 ? b : (b ^ 0x7FFFFFFFFFFFFFFFLL);
}
TJB_HD double key_to_ll(long long k) {
  long long b = k >= 0 ? k : (k ^ 0x7FFFFFFFFFFFFFFFLL);
#if defined(__CUDA_ARCH__)
  return __longlong_as_double(b);
#else
  double x; memcpy(&x, &b, 8); return x;
#endif
}

// multiply-accumulate a running product of positive doubles without overflow:
// every 4th call the binary exponent is moved into an integer.
struct LogProduct {
  double mant;
  int expo;
  TJB_HD void init() { mant = 1.0; expo = 0; }
  TJB_HD void mul(double v) { mant *= v; }
  TJB_HD void renorm() {
    const int hi = hi32(mant);
    const int ex = ((hi >> 20) & 0x7ff) - 1023;
    expo += ex;
    mant = mk64(hi - (ex << 20), lo32(mant));
  }
  TJB_HD double log_value() const { return log(mant) + 0.69314718055994530942 * (double)expo; }
};

// ---- per-sample evaluation -------------------------------------------------
// Returns ll for one prior sample.  `tab` is the staged epoch table.
template <int L, bool kJit>
TJB_HD double sample_ll(const StarParams &sp, const double *__restrict__ tab,
                        const SinCos *__restrict__ trig, double P, double e, double omega,
                        double M0, double s) {
  constexpr int RS = row_stride(L);
  constexpr int kEpochsPerIter = LLShape<L, kJit>::kEpochs;
  constexpr bool kXZ = LLShape<L, kJit>::kXZ;
  TrigCoef tc;
#if TJB_TRIM && defined(__CUDA_ARCH__)
  // `trig` is the kernel's shared-memory staging of the tables (interleaved copies, see
  // kepler.cuh); everything off the epoch loop's main path reads the table in global memory
  tc.load(sp.zero, sp.trig_table);
  tc.use_shared_table(trig);
  constexpr bool kSh = true;
#else
  tc.load(sp.zero, trig);
  constexpr bool kSh = false;
#endif
  const OrbitConsts oc = make_orbit_consts(tc, P, e, omega, M0);
  const int N = sp.n_times;

  double G[kTri<L>];
  double h[L];
  double quad0, logdet;

  if (!kJit) {
    // sums that involve the Kepler column z: z^T C^-1 z, z^T C^-1 y, z^T C^-1 T_k
    double Szz = 0.0, Szy = 0.0;
    double SzT[L];
#pragma unroll
    for (int k = 1; k < L; k++) SzT[k] = 0.0;
    const double *row = tab;
    int n = 0;
    // kEpochsPerIter epochs per iteration: independent Kepler chains for ILP
    TJB_EPOCH_GROUPS(n, N, row, RS) {
      double dt[kEpochsPerIter], z[kEpochsPerIter];
#if TJB_TRIM && defined(__CUDA_ARCH__)
      // dt and w in one 16-byte shared-memory load (rows are 16-byte aligned)
      double wj[kEpochsPerIter];
#pragma unroll
      for (int j = 0; j < kEpochsPerIter; j++) {
        const double2 dw = *reinterpret_cast<const double2 *>(row + j * RS);
        dt[j] = dw.x;
        wj[j] = dw.y;
      }
#else
#pragma unroll
      for (int j = 0; j < kEpochsPerIter; j++) dt[j] = row[j * RS];
#endif
      rv_unit_columns<kEpochsPerIter, false, kSh, kXZ>(oc, tc, dt, z, nullptr, sp.stats);
#pragma unroll
      for (int j = 0; j < kEpochsPerIter; j++) {
        const double *rj = row + j * RS;
#if TJB_TRIM && defined(__CUDA_ARCH__)
        Szz = fma(z[j] * z[j], wj[j], Szz);
#else
        Szz = fma(z[j] * z[j], rj[1], Szz);
#endif
        Szy = fma(z[j], rj[2], Szy);
#pragma unroll
        for (int k = 1; k < L; k++) SzT[k] = fma(z[j], rj[2 + k], SzT[k]);
      }
    }
    for (; n < N; n++, row += RS) {
      const double z = rv_unit_column<false, kSh, kXZ>(oc, tc, row[0], nullptr);
      Szz = fma(z * z, row[1], Szz);
      Szy = fma(z, row[2], Szy);
#pragma unroll
      for (int k = 1; k < L; k++) SzT[k] = fma(z, row[2 + k], SzT[k]);
    }
#pragma unroll
    for (int i = 0; i < kTri<L>; i++) G[i] = sp.Gc[i];
    G[0] = Szz;
#pragma unroll
    for (int k = 1; k < L; k++) G[tri<L>(0, k)] = SzT[k];
#pragma unroll
    for (int i = 1; i < L; i++) h[i] = sp.hc[i];
    h[0] = Szy;
    quad0 = sp.quad0;
    logdet = sp.c0;
  } else {
    const double s2 = sp.apply_jitter ? s * s : 0.0;
    double Syy = 0.0;
#pragma unroll
    for (int i = 0; i < kTri<L>; i++) G[i] = 0.0;
#pragma unroll
    for (int i = 0; i < L; i++) h[i] = 0.0;
    LogProduct lp;
    lp.init();
    const double *row = tab;
    int n = 0;
    auto accumulate = [&](const double *rj, double z, double var_n) {
      const double var = var_n + s2;
      // positive and normal unless ivar == 0 (var = inf): then the epoch has zero weight
      const double w = var < 1.0e300 ? rcp_pos(var) : 0.0;
      lp.mul(var);
      const double y = rj[2];
      // column values of M for this epoch: m[0] = z, m[k] = T_k
      double m[L];
      m[0] = z;
#pragma unroll
      for (int k = 1; k < L; k++) m[k] = rj[2 + k];
      const double wy = w * y;
      Syy = fma(wy, y, Syy);
#pragma unroll
      for (int i = 0; i < L; i++) {
        const double wm = w * m[i];
        h[i] = fma(wy, m[i], h[i]);
#pragma unroll
        for (int j = i; j < L; j++) G[tri<L>(i, j)] = fma(wm, m[j], G[tri<L>(i, j)]);
      }
    };
    TJB_EPOCH_GROUPS(n, N, row, RS) {
      double dt[kEpochsPerIter], z[kEpochsPerIter];
#if TJB_TRIM && defined(__CUDA_ARCH__)
      double vn[kEpochsPerIter];
#pragma unroll
      for (int j = 0; j < kEpochsPerIter; j++) {  // dt and var in one 16-byte load
        const double2 dv = *reinterpret_cast<const double2 *>(row + j * RS);
        dt[j] = dv.x;
        vn[j] = dv.y;
      }
      rv_unit_columns<kEpochsPerIter, false, kSh, kXZ>(oc, tc, dt, z, nullptr, sp.stats);
#pragma unroll
      for (int j = 0; j < kEpochsPerIter; j++) accumulate(row + j * RS, z[j], vn[j]);
#else
#pragma unroll
      for (int j = 0; j < kEpochsPerIter; j++) dt[j] = row[j * RS];
      rv_unit_columns<kEpochsPerIter, false, kSh, kXZ>(oc, tc, dt, z, nullptr, sp.stats);
#pragma unroll
      for (int j = 0; j < kEpochsPerIter; j++) accumulate(row + j * RS, z[j], row[j * RS + 1]);
#endif
      lp.renorm();  // at most kEpochsPerIter factors between renormalisations
    }
    for (; n < N; n++, row += RS) {
      accumulate(row, rv_unit_column<false, kSh, kXZ>(oc, tc, row[0], nullptr), row[1]);
      lp.renorm();
    }
    lp.renorm();
#pragma unroll
    for (int i = 1; i < L; i++) {
      G[tri<L>(i, i)] += sp.inv_Lambda[i];
      h[i] += sp.hc[i];
    }
    quad0 = Syy + sp.quad0;
    logdet = sp.c0 + lp.log_value();  // -sum log w = +sum log var
  }

  // K column: prior variance and mean (pyx:461-464)
  const double lamK = sp.K_prior_kind == 0
                          ? lambda_K_fixed_mass(P, e, sp.sigma_K0_sq, sp.inv_P0, sp.max_K_sq, true)
                          : sp.Lambda_K;
  const double ilamK = rcp_nz(lamK);
  G[0] += ilamK;
  h[0] = fma(sp.mu_K, ilamK, h[0]);
  quad0 = fma(sp.mu_K * sp.mu_K, ilamK, quad0);

  double rD[L];
  const bool ok = ldlt<L>(G, rD);
  double quad, detG;
  ldlt_quad<L>(G, rD, h, quad, detG);
  const double ll = -0.5 * ((quad0 - quad) + (logdet + log(lamK * detG)));
  // singular Ainv: the reference returns +inf (pyx:283-284, 381-382)
  return ok ? ll : (double)INFINITY;
}

#if defined(__CUDACC__)



// what varies between the two kernels below: where a sample's parameters come from
template <bool kJit>
TJB_D void load_sample(const PriorView &pv, long long ii, double &P, double &e, double &om,
                       double &M0, double &s) {
  if (pv.aos) {
    const double *r = pv.aos + 5 * ii;
    P = r[0]; e = r[1]; om = r[2]; M0 = r[3]; s = r[4];
    if (!kJit && pv.nonuniform && s != pv.s_expect) atomicOr(pv.nonuniform, 1);
  } else {
    P = pv.P[ii]; e = pv.e[ii]; om = pv.omega[ii]; M0 = pv.M0[ii];
    s = (kJit && pv.s) ? pv.s[ii] : 0.0;
  }
}
template <bool kJit>
TJB_D void load_sample(const PriorGenView &pv, long long ii, double &P, double &e, double &om,
                       double &M0, double &s) {
  double row[5];
  prior_row(pv.gen, (unsigned long long)(pv.index0 + ii), row);
  P = row[0]; e = row[1]; om = row[2]; M0 = row[3];
  s = kJit ? row[4] : 0.0;
}

// Epoch rows as a kernel parameter (TJB_UROW, the constant bank) instead of shared memory.
// Together with a CTA-uniform loop control this lets ptxas read the rows with uniform
// loads (LDCU.64 UR, c[0x0][UR + offset]) and feed them to the FP64 pipe as uniform-register
// operands: dt, w, w y and w T_k then cost no vector-register read, and an FP64 instruction
// with two vector operands issues every 2 cycles instead of every 3.
// Tables longer than kParamRowDoubles (768 epochs at L = 2, 512 at L = 3..4) take the
// shared-memory kernel.
#ifndef TJB_UROW
#define TJB_UROW 1
#endif
constexpr int kParamRowDoubles = 3072;  // 24 KB of the 32 KB parameter space
struct EpochRowsParam {
  double v[kParamRowDoubles];
};
struct EpochRowsShared {  // rows staged in shared memory from sp.table
  int unused;
};
struct EpochRowsGlobal {  // rows read where they lie in global memory (sp.table): tables too
  int unused[2];          // long for shared memory (> ~4900 epochs at L = 2); L1 / L2 hits
};

template <int L, bool kJit, typename View, typename Rows>
__global__ void __launch_bounds__((LLShape<L, kJit>::kThreads), TJB_LL_MIN_CTAS)
marginal_ll_kernel(const __grid_constant__ StarParams sp, const View pv,
                   const long long n, double *__restrict__ ll_out, const MaxKeys mk,
                   const __grid_constant__ Rows rows) {
  constexpr int kLLThreads = LLShape<L, kJit>::kThreads;
  constexpr bool kParamRows = sizeof(Rows) == sizeof(EpochRowsParam);
  constexpr bool kGlobalRows = sizeof(Rows) == sizeof(EpochRowsGlobal);
  // dynamic shared memory: [trig table (16-byte aligned) | epoch table if EpochRowsShared]
  extern __shared__ SinCos smem_trig[];
  for (int i = threadIdx.x; i < kTrigSlots; i += blockDim.x) smem_trig[i] = sp.trig_table[i];
  const double *tab;
  if constexpr (kParamRows) {
    tab = rows.v;
  } else if constexpr (kGlobalRows) {
    tab = sp.table;
  } else {
    double *stab = reinterpret_cast<double *>(smem_trig + kTrigSlots);
    const int tab_len = sp.n_times * row_stride(L);
    for (int i = threadIdx.x; i < tab_len; i += blockDim.x) stab[i] = sp.table[i];
    tab = stab;
  }
  __syncthreads();

  long long kmax = ll_to_key(-INFINITY);
  const long long stride = (long long)gridDim.x * blockDim.x;
  // all threads of a CTA stay in the loop together (the solver votes warp-wide), and the
  // loop control depends on the block index and kernel parameters only: ptxas can then keep
  // the epoch loop's counters (and, with kParamRows, the epoch rows) in uniform registers
  for (long long base = (long long)blockIdx.x * kLLThreads; base < n; base += stride) {
    const long long i = base + threadIdx.x;
    const bool valid = i < n;
    const long long ii = valid ? i : n - 1;
    double P, e, om, M0, s;
    load_sample<kJit>(pv, ii, P, e, om, M0, s);
    const double v = sample_ll<L, kJit>(sp, tab, smem_trig, P, e, om, M0, s);
    if (valid) {
      ll_out[i] = v;
      const long long k = ll_to_key(v);
      kmax = k > kmax ? k : kmax;
    }
  }

  if (mk.n > 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const long long other = __shfl_xor_sync(0xffffffffu, kmax, o);
      kmax = other > kmax ? other : kmax;
    }
    __shared__ long long wmax[kLLThreads / 32];
    if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = kmax;
    __syncthreads();
    if (threadIdx.x == 0) {
      long long m = wmax[0];
#pragma unroll
      for (int w = 1; w < kLLThreads / 32; w++) m = wmax[w] > m ? wmax[w] : m;
      atomicMax(mk.keys[0], m);
      for (int p = 1; p < mk.n; p++) atomicMax_system(mk.keys[p], m);
    }
  }
}

#endif  // __CUDACC__

}  // namespace tjb
