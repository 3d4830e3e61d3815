// host_emulation.cpp -- compiles the device math headers (kepler.cuh, linalg.cuh,
// marginal_ll.cuh, star_tables.hpp) for the HOST so the solver and the marginal-
// likelihood algebra can be checked on a machine without a GPU.
// Debug / test tooling only: not linked into libthejoker_b200.so, never on the
// product path (the product has no CPU path and fails loudly without CUDA).
// MUFU approximations are emulated by libm float calls, so the FP32 starter here is
// slightly *more* accurate than on the device; the FP64 stage is the same code.
#include "../thejoker_b200/csrc/star_tables.hpp"

using namespace tjb;

static const std::vector<SinCos> &trig_table() {
  static const std::vector<SinCos> t = make_trig_table();
  return t;
}

template <int L>
static void run(const StarParams &sp, const double *tab, bool jit, const double *chunk, long n,
                double *ll) {
  for (long i = 0; i < n; i++) {
    const double *r = chunk + 5 * i;
    ll[i] = jit ? sample_ll<L, true>(sp, tab, trig_table().data(), r[0], r[1], r[2], r[3], r[4])
                : sample_ll<L, false>(sp, tab, trig_table().data(), r[0], r[1], r[2], r[3], r[4]);
  }
}

extern "C" {

// z[n] for one sample; returns the solver statistics through stats[3]
void emu_design_column(double P, double e, double omega, double M0, const double *dt, int N,
                       double *z, int *stats) {
  SolveStats st = {0, 0, 0};
  TrigCoef tc;
  tc.load(0.0, trig_table().data());
  OrbitConsts oc = make_orbit_consts(tc, P, e, omega, M0);
  for (int n = 0; n < N; n++) z[n] = rv_unit_column<true>(oc, tc, dt[n], &st);
  if (stats) { stats[0] = st.extra_f32; stats[1] = st.extra_f64; stats[2] = st.not_converged; }
}

// sin / cos of an angle given in revolutions (scaled to the back-end's angle unit)
void emu_sincos_rev(double rev, double *s, double *c) {
  TrigCoef tc;
  tc.load(0.0, trig_table().data());
  sincos_units(tc, rev * kUnitsPerRev, *s, *c);
}

// ll for a chunk through the same code path selection as tjb_api.cu::run_ll:
// force_jit = 0 -> constant-jitter kernel with s = chunk[0][4]; 1 -> per-sample kernel
int emu_marginal_ll(int N, int L, double t_ref, const double *t, const double *rv,
                    const double *ivar, const double *trend, const double *mu,
                    const double *Lambda, int K_prior_kind, double sigma_K0, double P0,
                    double max_K, int jitter_mode, int force_jit, const double *chunk, long n,
                    double *ll) {
  StarHost st;
  st.N = N; st.L = L; st.t_ref = t_ref;
  st.t.assign(t, t + N); st.rv.assign(rv, rv + N); st.ivar.assign(ivar, ivar + N);
  if (L > 1) st.trend.assign(trend, trend + (size_t)N * (L - 1));
  for (int i = 0; i < L; i++) { st.mu[i] = mu[i]; st.Lambda[i] = Lambda[i]; }
  st.K_prior_kind = K_prior_kind; st.jitter_mode = jitter_mode;
  st.sigma_K0 = sigma_K0; st.P0 = P0; st.max_K = max_K;
  star_prepare(st);
  StarParams sp;
  std::vector<double> tab;
  const bool jit = force_jit && jitter_mode;
  if (jit) star_build_jit(st, sp, tab);
  else star_build_const(st, n > 0 ? chunk[4] : 0.0, sp, tab);
  sp.table = tab.data();
  switch (L) {
    case 1: run<1>(sp, tab.data(), jit, chunk, n, ll); break;
    case 2: run<2>(sp, tab.data(), jit, chunk, n, ll); break;
    case 3: run<3>(sp, tab.data(), jit, chunk, n, ll); break;
    case 4: run<4>(sp, tab.data(), jit, chunk, n, ll); break;
    case 5: run<5>(sp, tab.data(), jit, chunk, n, ll); break;
    case 6: run<6>(sp, tab.data(), jit, chunk, n, ll); break;
    case 7: run<7>(sp, tab.data(), jit, chunk, n, ll); break;
    case 8: run<8>(sp, tab.data(), jit, chunk, n, ll); break;
    default: return -1;
  }
  return 0;
}

// ---- prior_gen.cuh: the counter-based prior sampler --------------------------------------
void emu_philox4x32_10(const unsigned *ctr, const unsigned *key, unsigned *out) {
  uint32_t x0 = ctr[0], x1 = ctr[1], x2 = ctr[2], x3 = ctr[3];
  philox4x32_10(x0, x1, x2, x3, key[0], key[1]);
  out[0] = x0; out[1] = x1; out[2] = x2; out[3] = x3;
}
// rows[n, 5] for global indices index0 .. index0 + n; kind / p0 / p1 / scale: 5 entries each
void emu_prior_rows(const int *kind, const double *p0, const double *p1, const double *scale,
                    unsigned long long seed, long long index0, long n, double *rows) {
  PriorGenSpec ps;
  for (int k = 0; k < 5; k++) {
    ps.par[k].kind = kind[k]; ps.par[k].p0 = p0[k]; ps.par[k].p1 = p1[k]; ps.par[k].scale = scale[k];
  }
  ps.seed = seed;
  for (long i = 0; i < n; i++) prior_row(ps, (unsigned long long)(index0 + i), rows + 5 * i);
}
// the first n uniforms of one sample's stream
void emu_prior_uniforms(unsigned long long seed, unsigned long long index, int n, double *out) {
  Philox g;
  g.init(seed, index);
  for (int i = 0; i < n; i++) out[i] = g.next_open();
}

long long emu_ll_to_key(double x) { return ll_to_key(x); }
double emu_key_to_ll(long long k) { return key_to_ll(k); }

// the i-th double of numpy's Generator(PCG64).random() stream from (state, inc)
double emu_pcg64_double(unsigned long long s_hi, unsigned long long s_lo, unsigned long long i_hi,
                        unsigned long long i_lo, unsigned long long index);
}

#include "../thejoker_b200/csrc/accept.cuh"
extern "C" double emu_pcg64_double(unsigned long long s_hi, unsigned long long s_lo,
                                   unsigned long long i_hi, unsigned long long i_lo,
                                   unsigned long long index) {
  const u128 inc = make_u128(i_hi, i_lo);
  const Lcg128 j = lcg_power(inc, index + 1);
  return pcg_output_double(j.mult * make_u128(s_hi, s_lo) + j.plus);
}
