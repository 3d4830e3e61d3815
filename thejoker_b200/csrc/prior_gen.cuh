// prior_gen.cuh -- counter-based sampler of the nonlinear prior [P, e, omega, M0, s].
//
// Replaces, for prior samples that are drawn rather than read from a cache,
// JokerPrior.sample (thejoker/prior.py:297-407: pm.draw of the nonlinear parameters) for
// the distribution families the reference's default prior and its documented variants
// use: UniformLog on P (thejoker/distributions.py:17-51, rng_fn lines 25-28), Beta on e
// (Kipping13Global/Long/Short, distributions.py:155-176), uniform angles omega and M0
// (prior.py:437, 469-472), and for s a constant (prior.py:476-479), a LogNormal or a
// Normal.
//
// A sample is a pure function of (seed, global sample index): every sample owns a
// Philox4x32-10 stream whose counter is (index, block number), consumed in the fixed
// parameter order P, e, omega, M0, s.  So
//   * the likelihood kernel can generate the prior in registers and the 32-40 B/sample
//     prior cache never exists in HBM (marginal_ll.cuh, PriorView::gen),
//   * the rows of the few accepted samples are re-generated from their indices
//     (prior_rows_kernel) instead of being gathered,
//   * the prior is identical for any sharding of the index range over GPUs / ranks.
// The same code compiles for the host (tools/host_emulation.cpp), which is how the
// CPU tests check the distributions.
#pragma once

#include "kepler.cuh"

namespace tjb {

enum PriorKind {
  kPriorConstant = 0,    // p0
  kPriorUniform = 1,     // U(p0, p1)
  kPriorUniformLog = 2,  // exp(U(ln p0, ln p1))                (distributions.py:25-28)
  kPriorBeta = 3,        // Beta(p0, p1) = Ga/(Ga + Gb)
  kPriorLogNormal = 4,   // exp(N(p0, p1))
  kPriorNormal = 5       // N(p0, p1)
};

struct PriorDist {
  int kind;
  double p0, p1;
  double scale;  // unit conversion into the helper's internal units, applied last
};

struct PriorGenSpec {
  PriorDist par[5];  // P, e, omega, M0, s
  unsigned long long seed;
};

// ---- Philox4x32-10 (Salmon et al. 2011, "Parallel random numbers: as easy as 1, 2, 3") ----
TJB_HD void philox_mulhilo(uint32_t a, uint32_t b, uint32_t &hi, uint32_t &lo) {
#if defined(__CUDA_ARCH__)
  lo = a * b;
  hi = __umulhi(a, b);
#else
  const uint64_t p = (uint64_t)a * b;
  lo = (uint32_t)p;
  hi = (uint32_t)(p >> 32);
#endif
}
// x[4]: counter in, random words out
TJB_HD void philox4x32_10(uint32_t &x0, uint32_t &x1, uint32_t &x2, uint32_t &x3, uint32_t k0,
                          uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t h0, l0, h1, l1;
    philox_mulhilo(0xD2511F53u, x0, h0, l0);
    philox_mulhilo(0xCD9E8D57u, x2, h1, l1);
    x0 = h1 ^ x1 ^ k0;
    x1 = l1;
    x2 = h0 ^ x3 ^ k1;
    x3 = l0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

// the stream of one sample: counter = (index low, index high, block number, domain tag),
// key = seed; a block yields two 53-bit uniforms
constexpr uint32_t kPhiloxDomain = 0x544a4232u;  // "TJB2"
struct Philox {
  uint32_t key0, key1, c0, c1;
  uint32_t blk;             // next block number
  uint32_t w0, w1, w2, w3;  // current block
  int left;                 // unread uniforms in the current block (0..2)

  TJB_HD void init(unsigned long long seed, unsigned long long index) {
    key0 = (uint32_t)seed;
    key1 = (uint32_t)(seed >> 32);
    c0 = (uint32_t)index;
    c1 = (uint32_t)(index >> 32);
    blk = 0;
    left = 0;
    w0 = w1 = w2 = w3 = 0;
  }
  // uniform in (0, 1): 53 random bits, centred in their cell so that 0 and 1 never occur
  TJB_HD double next_open() {
    if (left == 0) {
      w0 = c0; w1 = c1; w2 = blk; w3 = kPhiloxDomain;
      philox4x32_10(w0, w1, w2, w3, key0, key1);
      blk++;
      left = 2;
    }
    const uint32_t hi = left == 2 ? w0 : w2, lo = left == 2 ? w1 : w3;
    left--;
    const double v = (double)(hi >> 6) * 134217728.0 + (double)(lo >> 5);  // 26 + 27 bits
    return (v + 0.5) * (1.0 / 9007199254740992.0);
  }
};

// two independent standard normals (Box-Muller on two 53-bit uniforms)
TJB_HD void normal_pair(Philox &g, double &z0, double &z1) {
  const double u1 = g.next_open(), u2 = g.next_open();
  const double r = sqrt(-2.0 * log(u1));
  double s, c;
#if defined(__CUDA_ARCH__)
  sincospi(2.0 * u2, &s, &c);
#else
  s = sin(kTwoPi * u2);
  c = cos(kTwoPi * u2);
#endif
  z0 = r * c;
  z1 = r * s;
}

// Gamma(a, 1), a > 0 (Marsaglia & Tsang 2000; a < 1 through Gamma(a + 1) U^(1/a)).
// `z` is a standard normal for the first attempt (the caller shares one Box-Muller pair
// between the two gammas of a Beta variate); further attempts draw their own.
TJB_HD double gamma_mt(Philox &g, double a, double z) {
  const double a1 = a < 1.0 ? a + 1.0 : a;
  const double d = a1 - 1.0 / 3.0;
  const double c = 1.0 / sqrt(9.0 * d);
  double v = 1.0;
  for (int it = 0; it < 64; it++) {  // 64 rejections in a row: probability ~ 1e-90
    if (it > 0) {
      double z1;
      normal_pair(g, z, z1);
    }
    const double t = fma(c, z, 1.0);
    if (t <= 0.0) continue;
    v = t * t * t;
    const double u = g.next_open();
    const double z2 = z * z;
    if (u < fma(-0.0331 * z2, z2, 1.0)) break;
    if (log(u) < 0.5 * z2 + d * (1.0 - v + log(v))) break;
  }
  double out = d * v;
  if (a < 1.0) out *= exp(log(g.next_open()) / a);
  return out;
}

TJB_HD double draw_param(Philox &g, const PriorDist &pd) {
  double x;
  switch (pd.kind) {
    case kPriorConstant:
      x = pd.p0;
      break;
    case kPriorUniform:
      x = fma(g.next_open(), pd.p1 - pd.p0, pd.p0);
      break;
    case kPriorUniformLog: {
      const double la = log(pd.p0);
      x = exp(fma(g.next_open(), log(pd.p1) - la, la));
      break;
    }
    case kPriorBeta: {
      double z0, z1;
      normal_pair(g, z0, z1);
      const double ga = gamma_mt(g, pd.p0, z0);
      const double gb = gamma_mt(g, pd.p1, z1);
      x = ga / (ga + gb);
      // an eccentricity: keep it inside [0, 1) (ga + gb cannot underflow to 0 for the
      // shape parameters in use, but a ratio that rounds to 1 must not reach the solver)
      x = x < 1.0 - 1.0e-16 ? x : 1.0 - 1.0e-16;
      break;
    }
    case kPriorLogNormal:
    case kPriorNormal: {
      double z0, z1;
      normal_pair(g, z0, z1);
      x = fma(z0, pd.p1, pd.p0);
      if (pd.kind == kPriorLogNormal) x = exp(x);
      break;
    }
    default:
      x = 0.0;
  }
  return x * pd.scale;
}

// row[5] = [P, e, omega, M0, s] of the sample with global index `index`
TJB_HD void prior_row(const PriorGenSpec &ps, unsigned long long index, double *row) {
  Philox g;
  g.init(ps.seed, index);
#pragma unroll
  for (int k = 0; k < 5; k++) row[k] = draw_param(g, ps.par[k]);
}

#if defined(__CUDACC__)

// materialise the columns for global indices [index0, index0 + n) (any of the output
// pointers may be null)
__global__ void __launch_bounds__(256)
prior_sample_kernel(const PriorGenSpec ps, const long long index0, const long long n,
                    double *__restrict__ P, double *__restrict__ e, double *__restrict__ omega,
                    double *__restrict__ M0, double *__restrict__ s) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    double row[5];
    prior_row(ps, (unsigned long long)(index0 + i), row);
    if (P) P[i] = row[0];
    if (e) e[i] = row[1];
    if (omega) omega[i] = row[2];
    if (M0) M0[i] = row[3];
    if (s) s[i] = row[4];
  }
}

// rows[k, 5] for an explicit list of global indices (the accepted samples)
__global__ void __launch_bounds__(128)
prior_rows_kernel(const PriorGenSpec ps, const long long *__restrict__ idx, const int k,
                  double *__restrict__ rows) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= k) return;
  double row[5];
  prior_row(ps, (unsigned long long)idx[i], row);
#pragma unroll
  for (int j = 0; j < 5; j++) rows[5 * i + j] = row[j];
}

#endif  // __CUDACC__

}  // namespace tjb
