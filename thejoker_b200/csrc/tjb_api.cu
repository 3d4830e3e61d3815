// tjb_api.cu -- C ABI of libthejoker_b200.so (see include/thejoker_b200.h).
// Host-side glue only: spec validation, the per-star constant tables (computed in
// long double), kernel dispatch on n_linear, stream handling.  No CPU compute path.
#include <cuda_runtime.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/thejoker_b200.h"
#include "accept.cuh"
#include "comm.hpp"
#include "marginal_ll.cuh"
#include "posterior.cuh"
#include "star_tables.hpp"

using namespace tjb;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}

#define CU(call)                                                                         \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess)                                                               \
      return fail(TJB_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));        \
  } while (0)

// Device allocations of the handles go through a small per-device cache of released
// blocks: a handle lives for one star, and cudaMalloc / cudaFree (which synchronises the
// device) dominated its creation and destruction.  Blocks are handed out again only to
// requests of a similar size; tjb_destroy synchronises the device before releasing.
constexpr int kMaxDevices = 64;
struct BlockCache {
  std::mutex mu;
  std::vector<std::pair<void *, size_t>> blocks[kMaxDevices];
  size_t held[kMaxDevices] = {0};
  static constexpr size_t kMaxHeld = (size_t)1 << 30;
  void *take(int dev, size_t need, size_t *got) {
    std::lock_guard<std::mutex> lock(mu);
    auto &v = blocks[dev];
    int best = -1;
    for (int i = 0; i < (int)v.size(); i++)
      if (v[i].second >= need && v[i].second <= std::max<size_t>(2 * need, 4096) &&
          (best < 0 || v[i].second < v[best].second))
        best = i;
    if (best < 0) return nullptr;
    void *p = v[best].first;
    *got = v[best].second;
    held[dev] -= v[best].second;
    v.erase(v.begin() + best);
    return p;
  }
  bool give(int dev, void *p, size_t bytes) {
    std::lock_guard<std::mutex> lock(mu);
    if (held[dev] + bytes > kMaxHeld || blocks[dev].size() >= 256) return false;
    blocks[dev].emplace_back(p, bytes);
    held[dev] += bytes;
    return true;
  }
};
BlockCache g_blocks;

struct DevBuf {
  void *p = nullptr;
  size_t bytes = 0;
  int dev = -1;
  int ensure(size_t need) {
    if (need <= bytes) return 0;
    release();
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return -1;
    need = (need + 255) & ~(size_t)255;
    p = g_blocks.take(dev, need, &bytes);
    if (p) return 0;
    if (cudaMalloc(&p, need) != cudaSuccess) {
      p = nullptr;
      return -1;
    }
    bytes = need;
    return 0;
  }
  void release() {
    if (p && !g_blocks.give(dev, p, bytes)) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
};

// page-locked host staging (ring slots of the host-streaming paths)
struct PinBuf {
  void *p = nullptr;
  size_t bytes = 0;
  int ensure(size_t need) {
    if (need <= bytes) return 0;
    release();
    if (cudaHostAlloc(&p, need, cudaHostAllocDefault) != cudaSuccess) {
      p = nullptr;
      return -1;
    }
    bytes = need;
    return 0;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    bytes = 0;
  }
};

// Staging resources of the host-streaming entry points, one set per device for the whole
// process: handles are short-lived (one per star) while the page-locked ring is expensive
// to allocate (cudaHostAlloc ~ 0.5 ms / MB), so it outlives them.  A call holds the
// device's mutex from its first use of the pool to its final synchronisation.
struct StagePool {
  std::mutex mu;
  DevBuf host_stage[2], host_ll[2];
  PinBuf pin_in[2], pin_out[2];
  cudaEvent_t pin_in_free[2] = {nullptr, nullptr};
  cudaStream_t aux_stream[2] = {nullptr, nullptr};
};
constexpr int kMaxStageDevices = kMaxDevices;
StagePool g_stage[kMaxStageDevices];

// per-device constants shared by every handle: SM count / compute capability and the
// sin/cos node table of the Kepler solver (never freed)
struct DeviceShared {
  std::mutex mu;
  bool ready = false;
  int n_sm = 0, cc_major = 0, cc_minor = 0;
  void *trig = nullptr;
};
DeviceShared g_shared[kMaxDevices];

}  // namespace

struct TjbHandle {
  int device = 0;
  cudaStream_t stream = nullptr;
  int n_sm = 0, cc_major = 0, cc_minor = 0;
  StarHost star;  // host copy of the spec + centring (star_tables.hpp)
  int N = 0, L = 0, jitter_mode = 0;
  // device tables and their parameter blocks
  DevBuf tab_const, tab_jit;
  std::vector<double> rows_const, rows_jit;  // host copies: passed as a kernel parameter
  StarParams sp_const, sp_jit;
  bool const_valid = false;
  double const_s = 0;
  // scratch
  DevBuf acc_mask, acc_counts, acc_offsets, acc_totals, misc, stats;
  DevBuf dist_counts, dist_send, dist_gather;  // tjb_accept_dist
  void *trig = nullptr;  // shared per-device sin/cos table (DeviceShared)
  StagePool *stg = nullptr;  // this device's host-streaming resources (shared by handles)
  int ll_ctas_per_sm = 0;
  long long last_nonfinite = 0;  // NaN / inf lls seen by the last accept call
  // extra (peer) keys the likelihood kernel max-updates besides the one passed per call
  long long *peer_keys[kMaxPeers] = {nullptr};
  int n_peer_keys = 0;
};

namespace {

// ---- per-star constants (star_tables.hpp) + upload ------------------------------

int upload_table(TjbHandle *h, DevBuf &buf, const std::vector<double> &tab, StarParams &sp) {
  if (buf.ensure(tab.size() * sizeof(double))) return fail(TJB_E_NOMEM, "cudaMalloc table");
  CU(cudaMemcpyAsync(buf.p, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice,
                     h->stream));
  CU(cudaStreamSynchronize(h->stream));  // tab is a caller-owned staging buffer
  sp.table = (const double *)buf.p;
  sp.stats = (unsigned long long *)h->stats.p;
  sp.trig_table = (const SinCos *)h->trig;
  return TJB_OK;
}

// (re)build the constant-jitter table for the given s
int build_const_table(TjbHandle *h, double s) {
  std::vector<double> &tab = h->rows_const;
  star_build_const(h->star, s, h->sp_const, tab);
  int rc = upload_table(h, h->tab_const, tab, h->sp_const);
  if (rc) return rc;
  h->const_valid = true;
  h->const_s = s;
  return TJB_OK;
}

int build_jit_table(TjbHandle *h) {
  std::vector<double> &tab = h->rows_jit;
  star_build_jit(h->star, h->sp_jit, tab);
  return upload_table(h, h->tab_jit, tab, h->sp_jit);
}

// ---- kernel dispatch ---------------------------------------------------------

// TJB_FORCE_SHARED_ROWS=1 in the environment selects the shared-memory-rows kernel for every
// table size (tests cover both kernels on the same inputs; bench / tuning comparisons)
std::atomic<int> g_force_shared_rows{-1};
bool tjb_force_shared_rows() {
  int v = g_force_shared_rows.load(std::memory_order_relaxed);
  if (v < 0) {
    const char *e = std::getenv("TJB_FORCE_SHARED_ROWS");
    v = (e && e[0] == '1') ? 1 : 0;
    g_force_shared_rows.store(v, std::memory_order_relaxed);
  }
  return v != 0;
}

constexpr size_t kMaxSmemPerCta = 227 * 1024;

template <int L, bool J, typename View, typename Rows>
int launch_ll_rows(TjbHandle *h, const StarParams &sp, const View &pv, long long n, double *d_ll,
                   long long *d_key, cudaStream_t stream, const Rows &rows) {
  constexpr bool kSharedRows = sizeof(Rows) == sizeof(EpochRowsShared);
  constexpr int kLLThreads = LLShape<L, J>::kThreads;
  auto kern = marginal_ll_kernel<L, J, View, Rows>;
  const size_t smem = (size_t)kTrigSlots * sizeof(SinCos) +
                      (kSharedRows ? (size_t)h->N * row_stride(L) * sizeof(double) : 0);
  if (smem > kMaxSmemPerCta)
    return fail(TJB_E_INVALID, "epoch table does not fit in shared memory (too many epochs)");
  if (smem > 40 * 1024)  // (the static part counts towards the 48 KB default limit)
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kLLThreads, smem));
  if (per_sm < 1) per_sm = 1;
  h->ll_ctas_per_sm = per_sm;
  const long long want = (n + kLLThreads - 1) / kLLThreads;
  const int grid = (int)std::max(1LL, std::min(want, (long long)h->n_sm * per_sm));
  MaxKeys mk;
  mk.n = 0;
  if (d_key) {
    mk.keys[mk.n++] = d_key;
    for (int p = 0; p < h->n_peer_keys && mk.n < kMaxPeers; p++) mk.keys[mk.n++] = h->peer_keys[p];
  }
  kern<<<grid, kLLThreads, smem, stream>>>(sp, pv, n, d_ll, mk, rows);
  CU(cudaGetLastError());
  return TJB_OK;
}

// Epoch rows travel as a kernel parameter when they fit (kParamRowDoubles; L <= kMaxParamRowsL
// keeps the number of kernel instantiations down), else they are staged in shared memory.
constexpr int kMaxParamRowsL = 4;
template <int L, bool J, typename View>
int launch_ll(TjbHandle *h, const StarParams &sp, const View &pv, long long n, double *d_ll,
              long long *d_key, cudaStream_t stream) {
#if TJB_UROW
  if constexpr (L <= kMaxParamRowsL) {
    const std::vector<double> &host_rows = J ? h->rows_jit : h->rows_const;
    if (!tjb_force_shared_rows() && host_rows.size() <= (size_t)kParamRowDoubles &&
        host_rows.size() == (size_t)h->N * row_stride(L)) {
      static thread_local EpochRowsParam rows;  // 24 KB: not on the stack of a slot thread
      std::memcpy(rows.v, host_rows.data(), host_rows.size() * sizeof(double));
      return launch_ll_rows<L, J, View, EpochRowsParam>(h, sp, pv, n, d_ll, d_key, stream, rows);
    }
  }
#endif
  // a table beyond one CTA's shared memory is read from global memory (L1 / L2 hits)
  if ((size_t)kTrigSlots * sizeof(SinCos) + (size_t)h->N * row_stride(L) * sizeof(double) >
      kMaxSmemPerCta)
    return launch_ll_rows<L, J, View, EpochRowsGlobal>(h, sp, pv, n, d_ll, d_key, stream,
                                                       EpochRowsGlobal{{0, 0}});
  return launch_ll_rows<L, J, View, EpochRowsShared>(h, sp, pv, n, d_ll, d_key, stream,
                                                     EpochRowsShared{0});
}

template <bool J, typename View>
int dispatch_ll(TjbHandle *h, const StarParams &sp, const View &pv, long long n, double *d_ll,
                long long *d_key, cudaStream_t stream) {
  switch (h->L) {
    case 1: return launch_ll<1, J>(h, sp, pv, n, d_ll, d_key, stream);
    case 2: return launch_ll<2, J>(h, sp, pv, n, d_ll, d_key, stream);
    case 3: return launch_ll<3, J>(h, sp, pv, n, d_ll, d_key, stream);
    case 4: return launch_ll<4, J>(h, sp, pv, n, d_ll, d_key, stream);
    case 5: return launch_ll<5, J>(h, sp, pv, n, d_ll, d_key, stream);
    case 6: return launch_ll<6, J>(h, sp, pv, n, d_ll, d_key, stream);
    case 7: return launch_ll<7, J>(h, sp, pv, n, d_ll, d_key, stream);
    case 8: return launch_ll<8, J>(h, sp, pv, n, d_ll, d_key, stream);
  }
  return fail(TJB_E_INVALID, "n_linear out of range");
}

// choose the kernel: a single jitter value for the whole call is folded into the
// table (constant-jitter kernel); otherwise the per-sample-jitter kernel runs.
template <typename View>
int run_ll(TjbHandle *h, const View &pv, bool uniform_s, double s_const, long long n,
           double *d_ll, long long *d_key, cudaStream_t stream) {
  if (n <= 0) return TJB_OK;
  CU(cudaSetDevice(h->device));
  if (h->jitter_mode == 0 || uniform_s) {
    const double s = h->jitter_mode ? s_const : 0.0;
    if (!h->const_valid || h->const_s != s) {
      int rc = build_const_table(h, s);
      if (rc) return rc;
    }
    return dispatch_ll<false>(h, h->sp_const, pv, n, d_ll, d_key, stream);
  }
  return dispatch_ll<true>(h, h->sp_jit, pv, n, d_ll, d_key, stream);
}

template <int L>
int launch_posterior(TjbHandle *h, const double *d_rows, int k, int clamp, int n_per,
                     const double *d_normals, double *d_ll, double *d_a, double *d_A,
                     double *d_draws) {
  const int grid = (k + 127) / 128;
  posterior_kernel<L><<<grid, 128, 0, h->stream>>>(h->sp_jit, d_rows, k, clamp,
                                                  (double)h->star.centre, h->star.centred ? 1 : -1, n_per,
                                                  d_normals, d_ll, d_a, d_A, d_draws);
  CU(cudaGetLastError());
  return TJB_OK;
}

int dispatch_posterior(TjbHandle *h, const double *d_rows, int k, int clamp, int n_per,
                       const double *d_normals, double *d_ll, double *d_a, double *d_A,
                       double *d_draws) {
  switch (h->L) {
    case 1: return launch_posterior<1>(h, d_rows, k, clamp, n_per, d_normals, d_ll, d_a, d_A, d_draws);
    case 2: return launch_posterior<2>(h, d_rows, k, clamp, n_per, d_normals, d_ll, d_a, d_A, d_draws);
    case 3: return launch_posterior<3>(h, d_rows, k, clamp, n_per, d_normals, d_ll, d_a, d_A, d_draws);
    case 4: return launch_posterior<4>(h, d_rows, k, clamp, n_per, d_normals, d_ll, d_a, d_A, d_draws);
    case 5: return launch_posterior<5>(h, d_rows, k, clamp, n_per, d_normals, d_ll, d_a, d_A, d_draws);
    case 6: return launch_posterior<6>(h, d_rows, k, clamp, n_per, d_normals, d_ll, d_a, d_A, d_draws);
    case 7: return launch_posterior<7>(h, d_rows, k, clamp, n_per, d_normals, d_ll, d_a, d_A, d_draws);
    case 8: return launch_posterior<8>(h, d_rows, k, clamp, n_per, d_normals, d_ll, d_a, d_A, d_draws);
  }
  return fail(TJB_E_INVALID, "n_linear out of range");
}

PcgParams make_pcg(const TjbPcg64 *pcg, long long offset, unsigned long long stride) {
  PcgParams pp;
  memset(&pp, 0, sizeof(pp));
  if (!pcg) return pp;
  pp.enabled = 1;
  pp.inc = make_u128(pcg->inc_hi, pcg->inc_lo);
  u128 st = make_u128(pcg->state_hi, pcg->state_lo);
  if (offset > 0) {
    const Lcg128 j = lcg_power(pp.inc, (uint64_t)offset);
    st = j.mult * st + j.plus;
  }
  pp.state = st;
  pp.stride = lcg_power(pp.inc, stride);
  return pp;
}

}  // namespace

// =============================================================================
extern "C" {

const char *tjb_last_error(void) { return g_err.c_str(); }
int tjb_version(void) { return TJB_VERSION; }

static int validate_spec(const TjbSpec *spec) {
  if (!spec) return fail(TJB_E_INVALID, "null spec");
  if (spec->n_times < 1) return fail(TJB_E_INVALID, "n_times must be >= 1");
  if (spec->n_linear < 1 || spec->n_linear > TJB_MAX_LINEAR)
    return fail(TJB_E_INVALID, "n_linear must be in 1..8");
  if (!spec->t || !spec->rv || !spec->ivar || (spec->n_linear > 1 && !spec->trend_M))
    return fail(TJB_E_INVALID, "null data array");
  if (spec->K_prior_kind != 0 && spec->K_prior_kind != 1)
    return fail(TJB_E_INVALID, "K_prior_kind must be 0 or 1");
  return TJB_OK;
}

static int load_star(TjbHandle *h, const TjbSpec *spec) {
  const int N = h->N = spec->n_times, L = h->L = spec->n_linear;
  StarHost &st = h->star;
  st.N = N;
  st.L = L;
  st.t_ref = spec->t_ref;
  st.t.assign(spec->t, spec->t + N);
  st.rv.assign(spec->rv, spec->rv + N);
  st.ivar.assign(spec->ivar, spec->ivar + N);
  st.trend.clear();
  if (L > 1) st.trend.assign(spec->trend_M, spec->trend_M + (size_t)N * (L - 1));
  memcpy(st.mu, spec->mu, sizeof(st.mu));
  memcpy(st.Lambda, spec->Lambda, sizeof(st.Lambda));
  st.K_prior_kind = spec->K_prior_kind;
  st.jitter_mode = h->jitter_mode = spec->jitter_mode ? 1 : 0;
  st.sigma_K0 = spec->sigma_K0;
  st.P0 = spec->P0;
  st.max_K = spec->max_K;
  star_prepare(st);
  h->const_valid = false;
  int rc = build_jit_table(h);
  if (rc == TJB_OK) rc = build_const_table(h, 0.0);
  return rc;
}

int tjb_create(const TjbSpec *spec, int device, TjbHandle **out) {
  if (!out) return fail(TJB_E_INVALID, "null argument");
  *out = nullptr;
  int rc = validate_spec(spec);
  if (rc) return rc;
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev < 1)
    return fail(TJB_E_CUDA, "no CUDA device available: libthejoker_b200 has no CPU path");
  if (device < 0 || device >= n_dev) return fail(TJB_E_INVALID, "device index out of range");
  CU(cudaSetDevice(device));
  if (device >= kMaxDevices) return fail(TJB_E_INVALID, "device index out of range");
  DeviceShared &ds = g_shared[device];
  {
    std::lock_guard<std::mutex> lock(ds.mu);
    if (!ds.ready) {
      CU(cudaDeviceGetAttribute(&ds.n_sm, cudaDevAttrMultiProcessorCount, device));
      CU(cudaDeviceGetAttribute(&ds.cc_major, cudaDevAttrComputeCapabilityMajor, device));
      CU(cudaDeviceGetAttribute(&ds.cc_minor, cudaDevAttrComputeCapabilityMinor, device));
      if (ds.cc_major >= 10) {
        const std::vector<SinCos> tt = make_trig_table();
        CU(cudaMalloc(&ds.trig, tt.size() * sizeof(SinCos)));
        CU(cudaMemcpy(ds.trig, tt.data(), tt.size() * sizeof(SinCos), cudaMemcpyHostToDevice));
      }
      ds.ready = true;
    }
  }
  if (ds.cc_major < 10)
    return fail(TJB_E_CUDA, "device is not sm_100-class; this library is built for sm_100a only");

  TjbHandle *h = new TjbHandle();
  h->device = device;
  h->stg = &g_stage[device];
  h->n_sm = ds.n_sm;
  h->cc_major = ds.cc_major;
  h->cc_minor = ds.cc_minor;
  h->trig = ds.trig;
  if (h->stats.ensure(4 * sizeof(unsigned long long)) ||
      cudaMemset(h->stats.p, 0, 4 * sizeof(unsigned long long)) != cudaSuccess) {
    tjb_destroy(h);
    return fail(TJB_E_NOMEM, "cudaMalloc stats");
  }
  rc = load_star(h, spec);
  if (rc != TJB_OK) {
    tjb_destroy(h);
    return rc;
  }
  *out = h;
  return TJB_OK;
}

int tjb_update_star(TjbHandle *h, const TjbSpec *spec) {
  if (!h) return fail(TJB_E_INVALID, "null handle");
  int rc = validate_spec(spec);
  if (rc) return rc;
  CU(cudaSetDevice(h->device));
  CU(cudaStreamSynchronize(h->stream));  // kernels in flight still read the old tables
  return load_star(h, spec);
}

void tjb_destroy(TjbHandle *h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();  // released blocks are reused by the next handle
  h->tab_const.release(); h->tab_jit.release();
  h->acc_mask.release(); h->acc_counts.release(); h->acc_offsets.release();
  h->acc_totals.release(); h->misc.release(); h->stats.release();
  h->dist_counts.release(); h->dist_send.release(); h->dist_gather.release();
  delete h;
}

int tjb_set_stream(TjbHandle *h, void *cuda_stream) {
  if (!h) return fail(TJB_E_INVALID, "null handle");
  h->stream = (cudaStream_t)cuda_stream;
  return TJB_OK;
}

int tjb_set_peer_keys(TjbHandle *h, int64_t *const *d_peer_keys, const int *peer_devices, int n) {
  if (!h) return fail(TJB_E_INVALID, "null handle");
  if (n < 0 || n >= kMaxPeers) return fail(TJB_E_INVALID, "too many peer keys");
  if (n > 0 && (!d_peer_keys || !peer_devices)) return fail(TJB_E_INVALID, "null argument");
  CU(cudaSetDevice(h->device));
  for (int p = 0; p < n; p++) {
    if (peer_devices[p] == h->device) continue;
    int can = 0;
    CU(cudaDeviceCanAccessPeer(&can, h->device, peer_devices[p]));
    if (!can) return fail(TJB_E_CUDA, "no peer access between the devices");
    // the fused max exchange does atomicMax_system on the peers' keys: peer *access* alone
    // (PCIe-only P2P, some virtualised topologies) does not guarantee native peer atomics
    int native = 0;
    CU(cudaDeviceGetP2PAttribute(&native, cudaDevP2PAttrNativeAtomicSupported, h->device,
                                 peer_devices[p]));
    if (!native)
      return fail(TJB_E_CUDA, "peer access without native atomics: use the host / NCCL max combine");
    cudaError_t e = cudaDeviceEnablePeerAccess(peer_devices[p], 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
    else if (e != cudaSuccess) return fail(TJB_E_CUDA, cudaGetErrorString(e));
  }
  h->n_peer_keys = n;
  for (int p = 0; p < n; p++) h->peer_keys[p] = (long long *)d_peer_keys[p];
  return TJB_OK;
}

int tjb_get_stats(TjbHandle *h, uint64_t *h_stats, int reset) {
  if (!h || !h_stats) return fail(TJB_E_INVALID, "null argument");
  CU(cudaSetDevice(h->device));
  CU(cudaStreamSynchronize(h->stream));
  for (int i = 0; i < 2; i++)
    if (h->stg->aux_stream[i]) CU(cudaStreamSynchronize(h->stg->aux_stream[i]));
  unsigned long long v[4];
  CU(cudaMemcpy(v, h->stats.p, sizeof(v), cudaMemcpyDeviceToHost));
  for (int i = 0; i < 4; i++) h_stats[i] = v[i];
  if (reset) CU(cudaMemset(h->stats.p, 0, sizeof(v)));
  return TJB_OK;
}

int tjb_device_info(TjbHandle *h, int *n_sm, int *ctas_per_sm, int *cc_major, int *cc_minor) {
  if (!h) return fail(TJB_E_INVALID, "null handle");
  if (n_sm) *n_sm = h->n_sm;
  if (ctas_per_sm) *ctas_per_sm = h->ll_ctas_per_sm;
  if (cc_major) *cc_major = h->cc_major;
  if (cc_minor) *cc_minor = h->cc_minor;
  return TJB_OK;
}

int tjb_marginal_ll_soa(TjbHandle *h, const double *d_P, const double *d_e, const double *d_omega,
                        const double *d_M0, const double *d_s, double s_const, int64_t n,
                        double *d_ll, int64_t *d_llmax_key) {
  if (!h) return fail(TJB_E_INVALID, "null handle");
  if (n < 0) return fail(TJB_E_INVALID, "negative n");
  if (n > 0 && (!d_P || !d_e || !d_omega || !d_M0 || !d_ll))
    return fail(TJB_E_INVALID, "null device pointer");
  PriorView pv = {d_P, d_e, d_omega, d_M0, d_s, nullptr, 0.0, nullptr};
  return run_ll(h, pv, d_s == nullptr, s_const, n, d_ll, (long long *)d_llmax_key, h->stream);
}

int tjb_marginal_ll_aos(TjbHandle *h, const double *d_chunk, int uniform_s, int64_t n,
                        double *d_ll, int64_t *d_llmax_key) {
  if (!h) return fail(TJB_E_INVALID, "null handle");
  if (n < 0) return fail(TJB_E_INVALID, "negative n");
  if (n == 0) return TJB_OK;
  if (!d_chunk || !d_ll) return fail(TJB_E_INVALID, "null device pointer");
  CU(cudaSetDevice(h->device));
  double s0 = 0.0;
  if (uniform_s && h->jitter_mode) {
    CU(cudaMemcpyAsync(&s0, d_chunk + 4, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
  }
  PriorView pv = {nullptr, nullptr, nullptr, nullptr, nullptr, d_chunk, 0.0, nullptr};
  return run_ll(h, pv, uniform_s != 0, s0, n, d_ll, (long long *)d_llmax_key, h->stream);
}

namespace {

// true when the driver would have to stage a copy from / to this host pointer itself
// (ordinary pageable memory): its staged path runs at a fraction of the PCIe rate
bool is_pageable(const void *ptr) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess) {
    cudaGetLastError();
    return true;
  }
  return a.type == cudaMemoryTypeUnregistered;
}

// ---- host copies between the caller's pageable arrays and the page-locked ring --------
//
// A persistent pool of copy threads (started on first use, never joined: the object is
// leaked on purpose so that no destructor runs under threads parked on its condition
// variable at process exit).  One job at a time; the calling thread works too.  Chunks of
// 1 MB are copied with non-temporal stores: the destination (the ring, or the caller's ll
// array) is not read again by the CPU, so the read-for-ownership of a cached store would
// only add a third stream of memory traffic to a copy that is bound by memory bandwidth.
struct CopyTask {
  char *dst;
  const char *src;
  size_t bytes;
};

void nt_copy(char *dst, const char *src, size_t n) {
#if defined(__SSE2__)
  if (n >= 4096) {
    const size_t head = (16 - ((uintptr_t)dst & 15)) & 15;
    std::memcpy(dst, src, head);
    dst += head; src += head; n -= head;
    const size_t blocks = n / 64;
    for (size_t i = 0; i < blocks; i++, dst += 64, src += 64) {
      const __m128i a = _mm_loadu_si128((const __m128i *)(src));
      const __m128i b = _mm_loadu_si128((const __m128i *)(src + 16));
      const __m128i c = _mm_loadu_si128((const __m128i *)(src + 32));
      const __m128i d = _mm_loadu_si128((const __m128i *)(src + 48));
      _mm_stream_si128((__m128i *)(dst), a);
      _mm_stream_si128((__m128i *)(dst + 16), b);
      _mm_stream_si128((__m128i *)(dst + 32), c);
      _mm_stream_si128((__m128i *)(dst + 48), d);
    }
    n -= blocks * 64;
    _mm_sfence();
  }
#endif
  std::memcpy(dst, src, n);
}

class CopyPool {
 public:
  static CopyPool &get() {
    static CopyPool *pool = new CopyPool();  // leaked, see above
    return *pool;
  }
  // copies every task; returns when all are done
  void run(const std::vector<CopyTask> &tasks) {
    if (tasks.empty()) return;
    if (n_workers_ == 0 || tasks.size() == 1) {
      for (const CopyTask &t : tasks) nt_copy(t.dst, t.src, t.bytes);
      return;
    }
    std::lock_guard<std::mutex> job(job_mu_);  // one job at a time
    {
      std::lock_guard<std::mutex> lk(mu_);
      tasks_ = &tasks;
      next_.store(0, std::memory_order_relaxed);
      pending_.store((long long)tasks.size(), std::memory_order_relaxed);
      generation_++;
    }
    cv_.notify_all();
    work();
    std::unique_lock<std::mutex> lk(mu_);
    // all tasks copied and no worker still looking at the task list
    done_cv_.wait(lk, [&] { return pending_.load(std::memory_order_acquire) == 0 && active_ == 0; });
    tasks_ = nullptr;
  }

 private:
  CopyPool() {
    unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    // ranks of one node share the cores (torchrun sets LOCAL_WORLD_SIZE)
    if (const char *e = std::getenv("LOCAL_WORLD_SIZE")) {
      const int w = std::atoi(e);
      if (w > 1) hw = std::max(1u, hw / (unsigned)w);
    }
    unsigned want = 15;
    if (const char *e = std::getenv("TJB_COPY_THREADS")) want = (unsigned)std::max(0, std::atoi(e));
    n_workers_ = (int)std::min(want, hw > 1 ? hw - 1 : 0u);
    for (int i = 0; i < n_workers_; i++) std::thread([this] { loop(); }).detach();
  }
  void work() {
    const std::vector<CopyTask> &tasks = *tasks_;
    for (;;) {
      const size_t i = next_.fetch_add(1, std::memory_order_relaxed);
      if (i >= tasks.size()) return;
      nt_copy(tasks[i].dst, tasks[i].src, tasks[i].bytes);
      if (pending_.fetch_sub(1, std::memory_order_acq_rel) == 1) {
        std::lock_guard<std::mutex> lk(mu_);
        done_cv_.notify_all();
      }
    }
  }
  void loop() {
    unsigned long long seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return generation_ != seen; });
        seen = generation_;
        if (!tasks_) continue;
        active_++;
      }
      work();
      {
        std::lock_guard<std::mutex> lk(mu_);
        active_--;
        done_cv_.notify_all();
      }
    }
  }
  std::mutex job_mu_, mu_;
  std::condition_variable cv_, done_cv_;
  const std::vector<CopyTask> *tasks_ = nullptr;
  std::atomic<size_t> next_{0};
  std::atomic<long long> pending_{0};
  unsigned long long generation_ = 0;
  int active_ = 0;
  int n_workers_ = 0;
};

// dst[c][0..count) = src[c][0..count) for n_cols columns, split over the copy pool
void parallel_copy(double *const *dst, const double *const *src, int n_cols, size_t count) {
  constexpr size_t kChunk = (size_t)1 << 17;  // doubles per task: 1 MB
  if (count * (size_t)n_cols < ((size_t)1 << 16)) {
    for (int c = 0; c < n_cols; c++) std::memcpy(dst[c], src[c], count * sizeof(double));
    return;
  }
  std::vector<CopyTask> tasks;
  tasks.reserve((size_t)n_cols * (count / kChunk + 1));
  for (size_t a = 0; a < count; a += kChunk)
    for (int c = 0; c < n_cols; c++) {
      const size_t m = std::min(kChunk, count - a);
      tasks.push_back({(char *)(dst[c] + a), (const char *)(src[c] + a), m * sizeof(double)});
    }
  CopyPool::get().run(tasks);
}

}  // namespace

int tjb_marginal_ll_host(TjbHandle *h, const double *h_chunk, int64_t n, double *h_ll) {
  if (!h) return fail(TJB_E_INVALID, "null handle");
  if (n < 0) return fail(TJB_E_INVALID, "negative n");
  if (n == 0) return TJB_OK;
  if (!h_chunk || !h_ll) return fail(TJB_E_INVALID, "null host pointer");
  CU(cudaSetDevice(h->device));
  std::lock_guard<std::mutex> lock(h->stg->mu);
  StagePool &sp = *h->stg;
  // Optimistic execution: assume every row carries the jitter of row 0 (true for
  // the default prior, s = const) and run the constant-jitter kernel, which also
  // checks the assumption on the rows it reads; if any row disagrees, rerun with
  // the per-sample-jitter kernel.  No host pass over the chunk is needed.
  const double s0 = h_chunk[4];
  const int64_t slice = 1 << 20;
  const int64_t m_max = std::min(slice, n);
  const bool stage_in = is_pageable(h_chunk), stage_out = is_pageable(h_ll);
  for (int i = 0; i < 2; i++) {
    if (!sp.aux_stream[i]) CU(cudaStreamCreateWithFlags(&sp.aux_stream[i], cudaStreamNonBlocking));
    if (!sp.pin_in_free[i]) CU(cudaEventCreateWithFlags(&sp.pin_in_free[i], cudaEventDisableTiming));
    if (sp.host_stage[i].ensure((size_t)m_max * 5 * sizeof(double)) ||
        sp.host_ll[i].ensure((size_t)m_max * sizeof(double)))
      return fail(TJB_E_NOMEM, "cudaMalloc staging");
    if ((stage_in && sp.pin_in[i].ensure((size_t)m_max * 5 * sizeof(double))) ||
        (stage_out && sp.pin_out[i].ensure((size_t)m_max * sizeof(double))))
      return fail(TJB_E_NOMEM, "cudaHostAlloc staging");
  }
  if (h->acc_totals.ensure(2 * sizeof(unsigned long long))) return fail(TJB_E_NOMEM, "cudaMalloc");
  int *d_flag = (int *)h->acc_totals.p;
  CU(cudaStreamSynchronize(h->stream));  // order after earlier work on the handle's stream
  CU(cudaMemset(d_flag, 0, sizeof(int)));
  int64_t out_lo[2] = {-1, -1}, out_m[2] = {0, 0};
  auto drain_out = [&](int b) -> int {  // pin_out[b] -> h_ll once its D2H is done
    if (out_lo[b] < 0) return TJB_OK;
    CU(cudaStreamSynchronize(sp.aux_stream[b]));
    double *dst = h_ll + out_lo[b];
    const double *src = (const double *)sp.pin_out[b].p;
    parallel_copy(&dst, &src, 1, (size_t)out_m[b]);
    out_lo[b] = -1;
    return TJB_OK;
  };
  for (int pass = 0; pass < 2; pass++) {
    const bool uniform = (pass == 0);
    if (!uniform && h->jitter_mode == 0) break;
    int b = 0;
    int64_t n_slices = 0;
    for (int64_t lo = 0; lo < n; lo += slice, b ^= 1, n_slices++) {
      const int64_t m = std::min(slice, n - lo);
      cudaStream_t st = sp.aux_stream[b];
      double *d_in = (double *)sp.host_stage[b].p, *d_out = (double *)sp.host_ll[b].p;
      const double *src = h_chunk + 5 * lo;
      if (stage_out) {  // this slot's previous slice must be out of pin_out first
        int rc = drain_out(b);
        if (rc) return rc;
      }
      if (stage_in) {
        if (n_slices >= 2) CU(cudaEventSynchronize(sp.pin_in_free[b]));
        double *dst = (double *)sp.pin_in[b].p;
        parallel_copy(&dst, &src, 1, (size_t)m * 5);
        src = dst;
      }
      CU(cudaMemcpyAsync(d_in, src, (size_t)m * 5 * sizeof(double), cudaMemcpyHostToDevice, st));
      if (stage_in) CU(cudaEventRecord(sp.pin_in_free[b], st));
      PriorView pv = {nullptr, nullptr, nullptr, nullptr, nullptr, d_in, s0,
                      (uniform && h->jitter_mode) ? d_flag : nullptr};
      int rc = run_ll(h, pv, uniform, s0, m, d_out, nullptr, st);
      if (rc) return rc;
      if (stage_out) {
        CU(cudaMemcpyAsync(sp.pin_out[b].p, d_out, (size_t)m * sizeof(double),
                           cudaMemcpyDeviceToHost, st));
        out_lo[b] = lo;
        out_m[b] = m;
      } else {
        CU(cudaMemcpyAsync(h_ll + lo, d_out, (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, st));
      }
    }
    CU(cudaStreamSynchronize(sp.aux_stream[0]));
    CU(cudaStreamSynchronize(sp.aux_stream[1]));
    for (int i = 0; i < 2; i++) {
      int rc = drain_out(i);
      if (rc) return rc;
    }
    int flag = 0;
    CU(cudaMemcpy(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost));
    if (!flag) break;
  }
  return TJB_OK;
}

namespace {

// Streams host columns through the GPU in slices on two streams: host (pageable ->
// page-locked ring, threaded) | H2D | kernel | D2H all overlap.  The ll values go to
// h_ll (host) or stay in d_ll (device, with the running max in d_key), or both.
int host_soa_stream(TjbHandle *h, const double *const cols[5], int n_cols, double s_const,
                    int64_t n, double *h_ll, double *d_ll, int64_t *d_key) {
  CU(cudaSetDevice(h->device));
  std::lock_guard<std::mutex> lock(h->stg->mu);
  bool stage_in = false;
  for (int c = 0; c < n_cols; c++) stage_in = stage_in || is_pageable(cols[c]);
  const bool stage_out = h_ll && is_pageable(h_ll);
  // slice: at least 8 slices for overlap, 2 MB..32 MB per column copy
  int64_t slice = 1 << 22;
  while (slice > (1 << 18) && slice * 8 > n) slice >>= 1;
  const int64_t m_max = std::min(slice, n);
  for (int i = 0; i < 2; i++) {
    if (!h->stg->aux_stream[i]) CU(cudaStreamCreateWithFlags(&h->stg->aux_stream[i], cudaStreamNonBlocking));
    if (!h->stg->pin_in_free[i]) CU(cudaEventCreateWithFlags(&h->stg->pin_in_free[i], cudaEventDisableTiming));
    if (h->stg->host_stage[i].ensure((size_t)m_max * 5 * sizeof(double)) ||
        (!d_ll && h->stg->host_ll[i].ensure((size_t)m_max * sizeof(double))))
      return fail(TJB_E_NOMEM, "cudaMalloc staging");
    if ((stage_in && h->stg->pin_in[i].ensure((size_t)m_max * n_cols * sizeof(double))) ||
        (stage_out && h->stg->pin_out[i].ensure((size_t)m_max * sizeof(double))))
      return fail(TJB_E_NOMEM, "cudaHostAlloc staging");
  }
  CU(cudaStreamSynchronize(h->stream));
  int64_t out_lo[2] = {-1, -1}, out_m[2] = {0, 0};  // slices parked in pin_out[b]
  auto drain_out = [&](int b) -> int {              // pin_out[b] -> h_ll once its D2H is done
    if (out_lo[b] < 0) return TJB_OK;
    CU(cudaStreamSynchronize(h->stg->aux_stream[b]));
    double *dst = h_ll + out_lo[b];
    const double *src = (const double *)h->stg->pin_out[b].p;
    parallel_copy(&dst, &src, 1, (size_t)out_m[b]);
    out_lo[b] = -1;
    return TJB_OK;
  };
  int b = 0;
  int64_t n_slices = 0;
  for (int64_t lo = 0; lo < n; lo += slice, b ^= 1, n_slices++) {
    const int64_t m = std::min(slice, n - lo);
    cudaStream_t st = h->stg->aux_stream[b];
    double *d_in = (double *)h->stg->host_stage[b].p;
    double *d_out = d_ll ? d_ll + lo : (double *)h->stg->host_ll[b].p;
    const double *src[5];
    if (stage_out) {  // this slot's previous slice (two back) must be out of pin_out first
      int rc = drain_out(b);
      if (rc) return rc;
    }
    if (stage_in) {
      if (n_slices >= 2) CU(cudaEventSynchronize(h->stg->pin_in_free[b]));  // slot's last H2D done
      double *dst[5];
      for (int c = 0; c < n_cols; c++) {
        dst[c] = (double *)h->stg->pin_in[b].p + (size_t)c * m_max;
        src[c] = cols[c] + lo;
      }
      parallel_copy(dst, src, n_cols, (size_t)m);
      for (int c = 0; c < n_cols; c++) src[c] = dst[c];
    } else {
      for (int c = 0; c < n_cols; c++) src[c] = cols[c] + lo;
    }
    for (int c = 0; c < n_cols; c++)
      CU(cudaMemcpyAsync(d_in + (size_t)c * m_max, src[c], (size_t)m * sizeof(double),
                         cudaMemcpyHostToDevice, st));
    if (stage_in) CU(cudaEventRecord(h->stg->pin_in_free[b], st));
    PriorView pv = {d_in, d_in + m_max, d_in + 2 * m_max, d_in + 3 * m_max,
                    n_cols == 5 ? d_in + 4 * m_max : nullptr, nullptr, 0.0, nullptr};
    int rc = run_ll(h, pv, n_cols == 4, s_const, m, d_out, (long long *)d_key, st);
    if (rc) return rc;
    if (h_ll) {
      if (stage_out) {
        CU(cudaMemcpyAsync(h->stg->pin_out[b].p, d_out, (size_t)m * sizeof(double),
                           cudaMemcpyDeviceToHost, st));
        out_lo[b] = lo;
        out_m[b] = m;
      } else {
        CU(cudaMemcpyAsync(h_ll + lo, d_out, (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, st));
      }
    }
  }
  CU(cudaStreamSynchronize(h->stg->aux_stream[0]));
  CU(cudaStreamSynchronize(h->stg->aux_stream[1]));
  for (int i = 0; i < 2; i++) {
    int rc = drain_out(i);
    if (rc) return rc;
  }
  return TJB_OK;
}

}  // namespace

int tjb_marginal_ll_host_soa(TjbHandle *h, const double *h_P, const double *h_e,
                             const double *h_omega, const double *h_M0, const double *h_s,
                             double s_const, int64_t n, double *h_ll) {
  if (!h) return fail(TJB_E_INVALID, "null handle");
  if (n < 0) return fail(TJB_E_INVALID, "negative n");
  if (n == 0) return TJB_OK;
  if (!h_P || !h_e || !h_omega || !h_M0 || !h_ll) return fail(TJB_E_INVALID, "null host pointer");
  const double *cols[5] = {h_P, h_e, h_omega, h_M0, h_s};
  return host_soa_stream(h, cols, h_s ? 5 : 4, s_const, n, h_ll, nullptr, nullptr);
}

int tjb_marginal_ll_host_soa_resident(TjbHandle *h, const double *h_P, const double *h_e,
                                      const double *h_omega, const double *h_M0,
                                      const double *h_s, double s_const, int64_t n, double *d_ll,
                                      int64_t *d_llmax_key) {
  if (!h) return fail(TJB_E_INVALID, "null handle");
  if (n < 0) return fail(TJB_E_INVALID, "negative n");
  if (n == 0) return TJB_OK;
  if (!h_P || !h_e || !h_omega || !h_M0) return fail(TJB_E_INVALID, "null host pointer");
  if (!d_ll) return fail(TJB_E_INVALID, "null device pointer");
  const double *cols[5] = {h_P, h_e, h_omega, h_M0, h_s};
  return host_soa_stream(h, cols, h_s ? 5 : 4, s_const, n, nullptr, d_ll, d_llmax_key);
}

// ---- drawn priors (prior_gen.cuh) ---------------------------------------------------

namespace {

int make_gen(const TjbPriorGen *gen, PriorGenSpec &ps) {
  if (!gen) return fail(TJB_E_INVALID, "null prior generator");
  for (int k = 0; k < 5; k++) {
    const TjbPriorDist &d = gen->par[k];
    if (d.kind < kPriorConstant || d.kind > kPriorNormal)
      return fail(TJB_E_INVALID, "unknown prior distribution kind");
    if (d.kind == kPriorUniformLog && !(d.p0 > 0.0 && d.p0 < d.p1))
      return fail(TJB_E_INVALID, "UniformLog needs 0 < a < b");
    if (d.kind == kPriorBeta && !(d.p0 > 0.0 && d.p1 > 0.0))
      return fail(TJB_E_INVALID, "Beta needs positive shape parameters");
    if ((d.kind == kPriorNormal || d.kind == kPriorLogNormal) && !(d.p1 >= 0.0))
      return fail(TJB_E_INVALID, "negative sigma");
    ps.par[k].kind = d.kind;
    ps.par[k].p0 = d.p0;
    ps.par[k].p1 = d.p1;
    ps.par[k].scale = d.scale;
  }
  ps.seed = gen->seed;
  return TJB_OK;
}

}  // namespace

int tjb_prior_sample(int device, void *cuda_stream, const TjbPriorGen *gen, int64_t index0,
                     int64_t n, double *d_P, double *d_e, double *d_omega, double *d_M0,
                     double *d_s) {
  if (n < 0 || index0 < 0) return fail(TJB_E_INVALID, "negative n / index0");
  PriorGenSpec ps;
  int rc = make_gen(gen, ps);
  if (rc || n == 0) return rc;
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev < 1)
    return fail(TJB_E_CUDA, "no CUDA device available: libthejoker_b200 has no CPU path");
  if (device < 0 || device >= n_dev) return fail(TJB_E_INVALID, "device index out of range");
  CU(cudaSetDevice(device));
  int n_sm = 0;
  CU(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device));
  const int grid = (int)std::min<long long>((n + 255) / 256, (long long)n_sm * 8);
  prior_sample_kernel<<<grid, 256, 0, (cudaStream_t)cuda_stream>>>(ps, index0, n, d_P, d_e, d_omega,
                                                                  d_M0, d_s);
  CU(cudaGetLastError());
  return TJB_OK;
}

int tjb_prior_rows(TjbHandle *h, const TjbPriorGen *gen, const int64_t *h_idx, int64_t k,
                   double *h_rows) {
  if (!h) return fail(TJB_E_INVALID, "null handle");
  if (k < 0 || k > (1 << 30)) return fail(TJB_E_INVALID, "bad row count");
  PriorGenSpec ps;
  int rc = make_gen(gen, ps);
  if (rc || k == 0) return rc;
  if (!h_idx || !h_rows) return fail(TJB_E_INVALID, "null host pointer");
  CU(cudaSetDevice(h->device));
  if (h->misc.ensure((size_t)k * 6 * sizeof(double))) return fail(TJB_E_NOMEM, "cudaMalloc");
  long long *d_idx = (long long *)h->misc.p;
  double *d_rows = (double *)h->misc.p + k;
  CU(cudaMemcpyAsync(d_idx, h_idx, (size_t)k * sizeof(int64_t), cudaMemcpyHostToDevice, h->stream));
  prior_rows_kernel<<<(int)((k + 127) / 128), 128, 0, h->stream>>>(ps, d_idx, (int)k, d_rows);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(h_rows, d_rows, (size_t)k * 5 * sizeof(double), cudaMemcpyDeviceToHost,
                     h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return TJB_OK;
}

int tjb_marginal_ll_generated(TjbHandle *h, const TjbPriorGen *gen, int64_t index0, int64_t n,
                              double *d_ll, int64_t *d_llmax_key) {
  if (!h) return fail(TJB_E_INVALID, "null handle");
  if (n < 0 || index0 < 0) return fail(TJB_E_INVALID, "negative n / index0");
  PriorGenView pv;
  int rc = make_gen(gen, pv.gen);
  if (rc || n == 0) return rc;
  if (!d_ll) return fail(TJB_E_INVALID, "null device pointer");
  pv.index0 = index0;
  // a constant jitter prior (the default, prior.py:476-479) is folded into the epoch table
  const PriorDist &ds = pv.gen.par[4];
  const bool uniform_s = ds.kind == kPriorConstant;
  return run_ll(h, pv, uniform_s, uniform_s ? ds.p0 * ds.scale : 0.0, n, d_ll,
                (long long *)d_llmax_key, h->stream);
}

// ---- accept -----------------------------------------------------------------

int64_t tjb_double_to_key(double x) { return (int64_t)ll_to_key(x); }
double tjb_key_to_double(int64_t key) { return key_to_ll((long long)key); }

int tjb_llmax_reset(TjbHandle *h, int64_t *d_llmax_key) {
  if (!h || !d_llmax_key) return fail(TJB_E_INVALID, "null argument");
  CU(cudaSetDevice(h->device));
  const long long k = ll_to_key(-INFINITY);
  CU(cudaMemcpyAsync(d_llmax_key, &k, sizeof(k), cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return TJB_OK;
}

int tjb_llmax_update(TjbHandle *h, const double *d_ll, int64_t n, int64_t *d_llmax_key) {
  if (!h || !d_llmax_key) return fail(TJB_E_INVALID, "null argument");
  if (n <= 0) return TJB_OK;
  CU(cudaSetDevice(h->device));
  const int grid = (int)std::min<long long>((n + 255) / 256, (long long)h->n_sm * 8);
  llmax_update_kernel<<<grid, 256, 0, h->stream>>>(d_ll, n, (long long *)d_llmax_key);
  CU(cudaGetLastError());
  return TJB_OK;
}

int tjb_llmax_get(TjbHandle *h, const int64_t *d_llmax_key, double *h_max) {
  if (!h || !d_llmax_key || !h_max) return fail(TJB_E_INVALID, "null argument");
  CU(cudaSetDevice(h->device));
  long long k = 0;
  CU(cudaMemcpyAsync(&k, d_llmax_key, sizeof(k), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  *h_max = key_to_ll(k);
  return TJB_OK;
}

int tjb_pcg64_uniform(TjbHandle *h, const TjbPcg64 *pcg, int64_t offset, int64_t n, double *d_out) {
  if (!h || !pcg) return fail(TJB_E_INVALID, "null argument");
  if (n <= 0) return TJB_OK;
  if (!d_out) return fail(TJB_E_INVALID, "null device pointer");
  CU(cudaSetDevice(h->device));
  const int grid = (int)std::min<long long>((n + kAccThreads - 1) / kAccThreads, (long long)h->n_sm * 8);
  const PcgParams pp = make_pcg(pcg, offset, (unsigned long long)grid * kAccThreads);
  pcg64_uniform_kernel<<<grid, kAccThreads, 0, h->stream>>>(pp, n, d_out);
  CU(cudaGetLastError());
  return TJB_OK;
}

namespace {

// flag / scan / scatter of one contiguous ll range on the handle's stream; the totals
// [accepted, near, non-finite] are left in h->acc_totals (device).  No host synchronisation.
constexpr int kAccTotals = 3;
int accept_local_async(TjbHandle *h, const double *d_ll, int64_t n, const int64_t *d_llmax_key,
                       const double *d_uniforms, const TjbPcg64 *pcg, int64_t pcg_offset,
                       int64_t index_base, int64_t max_keep, double near_tol, int64_t *d_idx) {
  if (h->acc_totals.ensure(kAccTotals * sizeof(unsigned long long)))
    return fail(TJB_E_NOMEM, "cudaMalloc accept scratch");
  CU(cudaMemsetAsync(h->acc_totals.p, 0, kAccTotals * sizeof(unsigned long long), h->stream));
  if (n <= 0) return TJB_OK;
  const long long n_words = (n + 31) / 32;
  const int wpc = acc_words_per_cta(n_words, h->n_sm);
  const int n_cta = (int)((n_words + wpc - 1) / wpc);
  if (h->acc_mask.ensure((size_t)n_words * sizeof(unsigned)) ||
      h->acc_counts.ensure((size_t)n_cta * sizeof(unsigned)) ||
      h->acc_offsets.ensure((size_t)n_cta * sizeof(unsigned long long)))
    return fail(TJB_E_NOMEM, "cudaMalloc accept scratch");
  const PcgParams pp = make_pcg(pcg, pcg_offset, kAccThreads);
  accept_flag_kernel<<<n_cta, kAccThreads, 0, h->stream>>>(
      d_ll, n, (const long long *)d_llmax_key, d_uniforms, pp, near_tol, wpc,
      (unsigned *)h->acc_mask.p, (unsigned *)h->acc_counts.p, (unsigned long long *)h->acc_totals.p);
  CU(cudaGetLastError());
  accept_scan_kernel<<<1, 1024, 0, h->stream>>>((const unsigned *)h->acc_counts.p, n_cta,
                                               (unsigned long long *)h->acc_offsets.p);
  CU(cudaGetLastError());
  if (max_keep > 0) {
    accept_scatter_kernel<<<n_cta, kAccThreads, 0, h->stream>>>(
        (const unsigned *)h->acc_mask.p, n, (const unsigned long long *)h->acc_offsets.p,
        index_base, max_keep, wpc, (long long *)d_idx);
    CU(cudaGetLastError());
  }
  return TJB_OK;
}

}  // namespace

int tjb_accept(TjbHandle *h, const double *d_ll, int64_t n, const int64_t *d_llmax_key,
               const double *d_uniforms, const TjbPcg64 *pcg, int64_t pcg_offset,
               int64_t index_base, int64_t max_keep, double near_tol, int64_t *d_idx,
               int64_t *h_counts) {
  if (!h || !h_counts) return fail(TJB_E_INVALID, "null argument");
  h_counts[0] = h_counts[1] = h_counts[2] = 0;
  if (n <= 0) return TJB_OK;
  if (!d_ll || !d_llmax_key) return fail(TJB_E_INVALID, "null device pointer");
  if ((d_uniforms == nullptr) == (pcg == nullptr))
    return fail(TJB_E_INVALID, "exactly one of d_uniforms / pcg must be given");
  if (max_keep < 0) return fail(TJB_E_INVALID, "negative max_keep");
  if (max_keep > 0 && !d_idx) return fail(TJB_E_INVALID, "null index buffer");
  CU(cudaSetDevice(h->device));
  int rc = accept_local_async(h, d_ll, n, d_llmax_key, d_uniforms, pcg, pcg_offset, index_base,
                              max_keep, near_tol, d_idx);
  if (rc) return rc;
  unsigned long long tot[kAccTotals] = {0, 0, 0};
  CU(cudaMemcpyAsync(tot, h->acc_totals.p, sizeof(tot), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  h_counts[0] = (int64_t)tot[0];
  h_counts[1] = std::min<int64_t>((int64_t)tot[0], max_keep);
  h_counts[2] = (int64_t)tot[1];
  h->last_nonfinite = (long long)tot[2];
  return TJB_OK;
}

int tjb_accept_nonfinite(TjbHandle *h, int64_t *h_count) {
  if (!h || !h_count) return fail(TJB_E_INVALID, "null argument");
  *h_count = (int64_t)h->last_nonfinite;
  return TJB_OK;
}

// ---- multi-rank accept: NCCL inside the library (comm.hpp) -----------------------------

#define NC(call)                                                                          \
  do {                                                                                    \
    ncclResult_t r_ = (call);                                                             \
    if (r_ != ncclSuccess)                                                                \
      return fail(TJB_E_CUDA, std::string(#call) + ": " + nccl_api().GetErrorString(r_));  \
  } while (0)

int tjb_comm_unique_id(void *out_id) {
  if (!out_id) return fail(TJB_E_INVALID, "null argument");
  NcclApi &nc = nccl_api();
  if (!nc.ok) return fail(TJB_E_CUDA, nc.error);
  static_assert(sizeof(ncclUniqueId) == TJB_COMM_ID_BYTES, "NCCL unique id size");
  ncclUniqueId id;
  NC(nc.GetUniqueId(&id));
  memcpy(out_id, &id, sizeof(id));
  return TJB_OK;
}

int tjb_comm_create(const void *id_bytes, int n_ranks, int rank, int device, TjbComm **out) {
  if (!id_bytes || !out) return fail(TJB_E_INVALID, "null argument");
  *out = nullptr;
  if (n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(TJB_E_INVALID, "bad rank / n_ranks");
  NcclApi &nc = nccl_api();
  if (!nc.ok) return fail(TJB_E_CUDA, nc.error);
  CU(cudaSetDevice(device));
  ncclUniqueId id;
  memcpy(&id, id_bytes, sizeof(id));
  ncclComm_t comm = nullptr;
  NC(nc.CommInitRank(&comm, n_ranks, id, rank));
  TjbComm *c = new TjbComm();
  c->comm = comm;
  c->n_ranks = n_ranks;
  c->rank = rank;
  c->device = device;
  *out = c;
  return TJB_OK;
}

void tjb_comm_destroy(TjbComm *comm) {
  if (!comm) return;
  if (comm->comm && nccl_api().ok) {
    cudaSetDevice(comm->device);
    nccl_api().CommDestroy(comm->comm);
  }
  delete comm;
}

int tjb_comm_allreduce_max_key(TjbHandle *h, TjbComm *comm, int64_t *d_llmax_key) {
  if (!h || !comm || !d_llmax_key) return fail(TJB_E_INVALID, "null argument");
  if (comm->device != h->device) return fail(TJB_E_INVALID, "communicator is on another device");
  CU(cudaSetDevice(h->device));
  NC(nccl_api().AllReduce(d_llmax_key, d_llmax_key, 1, ncclInt64, ncclMax, comm->comm, h->stream));
  return TJB_OK;
}

int tjb_accept_dist(TjbHandle *h, TjbComm *comm, const double *d_ll, int64_t n_local,
                    int64_t *d_llmax_key, const double *d_uniforms, const TjbPcg64 *pcg,
                    int64_t global_offset, int64_t max_keep, double near_tol, int64_t *d_idx,
                    int64_t *h_counts) {
  if (!h || !comm || !h_counts) return fail(TJB_E_INVALID, "null argument");
  h_counts[0] = h_counts[1] = h_counts[2] = 0;
  if (comm->device != h->device) return fail(TJB_E_INVALID, "communicator is on another device");
  if (n_local < 0 || global_offset < 0) return fail(TJB_E_INVALID, "negative n / offset");
  if (!d_llmax_key || (n_local > 0 && !d_ll)) return fail(TJB_E_INVALID, "null device pointer");
  if (n_local > 0 && (d_uniforms == nullptr) == (pcg == nullptr))
    return fail(TJB_E_INVALID, "exactly one of d_uniforms / pcg must be given");
  if (max_keep < 0) return fail(TJB_E_INVALID, "negative max_keep");
  if (max_keep > 0 && !d_idx) return fail(TJB_E_INVALID, "null index buffer");
  NcclApi &nc = nccl_api();
  CU(cudaSetDevice(h->device));
  const int W = comm->n_ranks;
  cudaStream_t st = h->stream;
  // (1) lls.max() over all shards: integer MAX all-reduce of the order-preserving key
  NC(nc.AllReduce(d_llmax_key, d_llmax_key, 1, ncclInt64, ncclMax, comm->comm, st));
  // (2) local flag / scan / scatter; uniforms and indices are addressed globally
  const int64_t keep_local = std::min(max_keep, n_local);
  if (h->dist_send.ensure((size_t)std::max<int64_t>(keep_local, 1) * sizeof(int64_t)) ||
      h->dist_counts.ensure((size_t)W * kAccTotals * sizeof(unsigned long long)))
    return fail(TJB_E_NOMEM, "cudaMalloc accept scratch");
  int rc = accept_local_async(h, d_ll, n_local, d_llmax_key, d_uniforms, pcg, global_offset,
                              global_offset, keep_local, near_tol, (int64_t *)h->dist_send.p);
  if (rc) return rc;
  // (3) every rank learns every rank's [accepted, near, non-finite]
  NC(nc.AllGather(h->acc_totals.p, h->dist_counts.p, kAccTotals, ncclUint64, comm->comm, st));
  std::vector<unsigned long long> cnt((size_t)W * kAccTotals);
  CU(cudaMemcpyAsync(cnt.data(), h->dist_counts.p, cnt.size() * sizeof(unsigned long long),
                     cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  // good_samples_idx[:max_keep] over the rank-ordered concatenation
  // (likelihood_helpers.py:109): rank r keeps what is left of max_keep after ranks < r
  std::vector<int64_t> kept(W), pos(W);
  int64_t total = 0, near = 0, written = 0, m = 0;
  h->last_nonfinite = 0;
  for (int r = 0; r < W; r++) {
    const int64_t a = (int64_t)cnt[kAccTotals * r];
    total += a;
    near += (int64_t)cnt[kAccTotals * r + 1];
    h->last_nonfinite += (long long)cnt[kAccTotals * r + 2];
    pos[r] = written;
    kept[r] = std::max<int64_t>(0, std::min(a, max_keep - written));
    written += kept[r];
    m = std::max(m, kept[r]);
  }
  h_counts[0] = total;
  h_counts[1] = written;
  h_counts[2] = near;
  if (m == 0) return TJB_OK;
  // (4) fixed-size all-gather of the first m local indices of every rank, then the kept
  //     prefixes are laid end to end
  if (h->dist_send.bytes < (size_t)m * sizeof(int64_t)) {  // this rank keeps fewer than m
    DevBuf bigger;
    if (bigger.ensure((size_t)m * sizeof(int64_t))) return fail(TJB_E_NOMEM, "cudaMalloc");
    if (kept[comm->rank] > 0)
      CU(cudaMemcpyAsync(bigger.p, h->dist_send.p, (size_t)kept[comm->rank] * sizeof(int64_t),
                         cudaMemcpyDeviceToDevice, st));
    CU(cudaStreamSynchronize(st));
    h->dist_send.release();
    h->dist_send = bigger;
  }
  if (h->dist_gather.ensure((size_t)W * m * sizeof(int64_t))) return fail(TJB_E_NOMEM, "cudaMalloc");
  NC(nc.AllGather(h->dist_send.p, h->dist_gather.p, (size_t)m, ncclInt64, comm->comm, st));
  for (int r = 0; r < W; r++)
    if (kept[r] > 0)
      CU(cudaMemcpyAsync(d_idx + pos[r], (const int64_t *)h->dist_gather.p + (size_t)r * m,
                         (size_t)kept[r] * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
  CU(cudaStreamSynchronize(st));
  return TJB_OK;
}

// ---- posterior ------------------------------------------------------------------

static int posterior_common(TjbHandle *h, const double *h_rows, int64_t k, int clamp_K, int n_per,
                            const double *h_normals, double *h_ll, double *h_a, double *h_A,
                            double *h_draws) {
  if (!h) return fail(TJB_E_INVALID, "null handle");
  if (k < 0 || k > (1 << 30)) return fail(TJB_E_INVALID, "bad row count");
  if (k == 0) return TJB_OK;
  if (!h_rows) return fail(TJB_E_INVALID, "null rows");
  CU(cudaSetDevice(h->device));
  const int L = h->L;
  const size_t b_rows = (size_t)k * 5 * 8, b_ll = (size_t)k * 8, b_a = (size_t)k * L * 8,
               b_A = (size_t)k * L * L * 8;
  const size_t b_nrm = h_draws ? (size_t)k * n_per * L * 8 : 0;
  const size_t b_drw = h_draws ? (size_t)k * n_per * (5 + L) * 8 : 0;
  if (h->misc.ensure(b_rows + b_ll + b_a + b_A + b_nrm + b_drw))
    return fail(TJB_E_NOMEM, "cudaMalloc posterior scratch");
  char *base = (char *)h->misc.p;
  double *d_rows = (double *)base;
  double *d_ll = (double *)(base + b_rows);
  double *d_a = (double *)(base + b_rows + b_ll);
  double *d_A = (double *)(base + b_rows + b_ll + b_a);
  double *d_nrm = (double *)(base + b_rows + b_ll + b_a + b_A);
  double *d_drw = (double *)(base + b_rows + b_ll + b_a + b_A + b_nrm);
  CU(cudaMemcpyAsync(d_rows, h_rows, b_rows, cudaMemcpyHostToDevice, h->stream));
  if (h_draws) CU(cudaMemcpyAsync(d_nrm, h_normals, b_nrm, cudaMemcpyHostToDevice, h->stream));
  int rc = dispatch_posterior(h, d_rows, (int)k, clamp_K, n_per, h_draws ? d_nrm : nullptr, d_ll,
                              d_a, d_A, h_draws ? d_drw : nullptr);
  if (rc) return rc;
  if (h_ll) CU(cudaMemcpyAsync(h_ll, d_ll, b_ll, cudaMemcpyDeviceToHost, h->stream));
  if (h_a) CU(cudaMemcpyAsync(h_a, d_a, b_a, cudaMemcpyDeviceToHost, h->stream));
  if (h_A) CU(cudaMemcpyAsync(h_A, d_A, b_A, cudaMemcpyDeviceToHost, h->stream));
  if (h_draws) CU(cudaMemcpyAsync(h_draws, d_drw, b_drw, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return TJB_OK;
}

// ---- multi-star loop ---------------------------------------------------------------

namespace {

// one star in flight: handle, stream and device buffers of a slot thread
struct StarSlot {
  TjbHandle *h = nullptr;
  cudaStream_t st = nullptr;
  DevBuf ll, key, idx, rows, nrm, out, oll;
  int rc = TJB_OK;
  std::string err;
};

// everything one star needs, on the slot's stream; one host synchronisation for the
// accept counts and one at the end
int run_star(StarSlot &sl, const TjbMultiStarJob &job, int64_t i, bool load) {
  TjbHandle *h = sl.h;
  const int L = job.specs[0].n_linear, W = 5 + L;
  const int64_t keep = std::min(job.max_keep, job.n_prior);
  int rc;
  if (load && (rc = tjb_update_star(h, &job.specs[i]))) return rc;
  static const long long kNegInf = ll_to_key(-INFINITY);
  CU(cudaMemcpyAsync(sl.key.p, &kNegInf, sizeof(kNegInf), cudaMemcpyHostToDevice, sl.st));
  rc = tjb_marginal_ll_soa(h, job.d_P, job.d_e, job.d_omega, job.d_M0, job.d_s, job.s_const,
                           job.n_prior, (double *)sl.ll.p, (int64_t *)sl.key.p);
  if (rc) return rc;
  int64_t *counts = job.h_counts + 3 * i;
  rc = tjb_accept(h, (const double *)sl.ll.p, job.n_prior, (const int64_t *)sl.key.p, nullptr,
                  &job.pcg[i], 0, 0, keep, job.near_tol, (int64_t *)sl.idx.p, counts);
  if (rc) return rc;
  const int64_t k = counts[1];
  long long key = 0;
  CU(cudaMemcpyAsync(&key, sl.key.p, sizeof(key), cudaMemcpyDeviceToHost, sl.st));
  if (k > 0 && job.h_idx)
    CU(cudaMemcpyAsync(job.h_idx + i * job.max_keep, sl.idx.p, (size_t)k * sizeof(int64_t),
                       cudaMemcpyDeviceToHost, sl.st));
  if (k > 0 && job.n_per > 0) {
    const size_t n_draw = (size_t)k * job.n_per;
    gather_rows_kernel<<<(int)((k + 127) / 128), 128, 0, sl.st>>>(
        job.d_P, job.d_e, job.d_omega, job.d_M0, job.d_s, job.s_const, (const long long *)sl.idx.p,
        (int)k, (double *)sl.rows.p);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(sl.nrm.p, job.h_normals + (size_t)i * job.max_keep * job.n_per * L,
                       n_draw * L * sizeof(double), cudaMemcpyHostToDevice, sl.st));
    rc = dispatch_posterior(h, (const double *)sl.rows.p, (int)k, job.clamp_K, job.n_per,
                            (const double *)sl.nrm.p, (double *)sl.oll.p, nullptr, nullptr,
                            (double *)sl.out.p);
    if (rc) return rc;
    CU(cudaMemcpyAsync(job.h_rows + (size_t)i * job.max_keep * job.n_per * W, sl.out.p,
                       n_draw * W * sizeof(double), cudaMemcpyDeviceToHost, sl.st));
    if (job.h_ll)
      CU(cudaMemcpyAsync(job.h_ll + i * job.max_keep, sl.oll.p, (size_t)k * sizeof(double),
                         cudaMemcpyDeviceToHost, sl.st));
  }
  CU(cudaStreamSynchronize(sl.st));
  if (job.h_llmax) job.h_llmax[i] = key_to_ll(key);
  return TJB_OK;
}

// a slot thread: takes the next star off the shared counter until none is left or any
// slot has failed
void slot_main(StarSlot &sl, int device, const TjbMultiStarJob &job, std::atomic<int64_t> &next,
               std::atomic<int> &failed) {
  auto body = [&]() -> int {
    CU(cudaSetDevice(device));
    CU(cudaStreamCreateWithFlags(&sl.st, cudaStreamNonBlocking));
    const int L = job.specs[0].n_linear;
    const size_t keep = (size_t)std::max<int64_t>(1, std::min(job.max_keep, job.n_prior));
    const size_t n_draw = keep * (size_t)std::max(job.n_per, 1);
    if (sl.ll.ensure((size_t)job.n_prior * sizeof(double)) || sl.key.ensure(sizeof(long long)) ||
        sl.idx.ensure(keep * sizeof(int64_t)) || sl.rows.ensure(keep * 5 * sizeof(double)) ||
        sl.nrm.ensure(n_draw * L * sizeof(double)) ||
        sl.out.ensure(n_draw * (5 + L) * sizeof(double)) || sl.oll.ensure(keep * sizeof(double)))
      return fail(TJB_E_NOMEM, "cudaMalloc multi-star slot");
    bool first = true;
    for (;;) {
      const int64_t i = next.fetch_add(1);
      if (i >= job.n_stars || failed.load()) break;
      int rc;
      if (first) {
        if ((rc = tjb_create(&job.specs[i], device, &sl.h))) return rc;
        if ((rc = tjb_set_stream(sl.h, sl.st))) return rc;
      }
      if ((rc = run_star(sl, job, i, !first))) return rc;
      first = false;
    }
    return TJB_OK;
  };
  sl.rc = body();
  if (sl.rc) {
    sl.err = g_err;  // thread-local: hand the message to the calling thread
    failed.store(1);
  }
  if (sl.st) cudaStreamSynchronize(sl.st);
  sl.ll.release(); sl.key.release(); sl.idx.release(); sl.rows.release();
  sl.nrm.release(); sl.out.release(); sl.oll.release();
}

}  // namespace

int tjb_multistar_rejection(int device, const TjbMultiStarJob *job) {
  if (!job) return fail(TJB_E_INVALID, "null job");
  if (job->n_stars < 0 || job->n_prior < 1 || job->max_keep < 0 || job->n_per < 0)
    return fail(TJB_E_INVALID, "bad sizes");
  if (job->n_stars == 0) return TJB_OK;
  if (job->n_slots < 1 || job->n_slots > 64) return fail(TJB_E_INVALID, "n_slots must be in 1..64");
  if (!job->specs || !job->pcg || !job->h_counts) return fail(TJB_E_INVALID, "null argument");
  if (!job->d_P || !job->d_e || !job->d_omega || !job->d_M0)
    return fail(TJB_E_INVALID, "null device pointer");
  if (job->max_keep > 0 && !job->h_idx && job->n_per == 0)
    return fail(TJB_E_INVALID, "no output requested");
  if (job->n_per > 0 && job->max_keep > 0 && (!job->h_normals || !job->h_rows))
    return fail(TJB_E_INVALID, "null normals / rows");
  for (int64_t i = 0; i < job->n_stars; i++) {
    int rc = validate_spec(&job->specs[i]);
    if (rc) return rc;
    if (job->specs[i].n_linear != job->specs[0].n_linear)
      return fail(TJB_E_INVALID, "all stars of a job must have the same n_linear");
  }
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev < 1)
    return fail(TJB_E_CUDA, "no CUDA device available: libthejoker_b200 has no CPU path");
  if (device < 0 || device >= n_dev) return fail(TJB_E_INVALID, "device index out of range");

  const int n_slots = (int)std::min<int64_t>(job->n_slots, job->n_stars);
  std::vector<StarSlot> slots(n_slots);
  std::atomic<int64_t> next(0);
  std::atomic<int> failed(0);
  std::vector<std::thread> pool;
  for (int k = 1; k < n_slots; k++)
    pool.emplace_back(slot_main, std::ref(slots[k]), device, std::cref(*job), std::ref(next),
                      std::ref(failed));
  slot_main(slots[0], device, *job, next, failed);
  for (auto &t : pool) t.join();
  // handles and streams go last: tjb_destroy synchronises the whole device
  for (auto &sl : slots) {
    if (sl.h) tjb_destroy(sl.h);
    if (sl.st) cudaStreamDestroy(sl.st);
  }
  for (auto &sl : slots)
    if (sl.rc) return fail(sl.rc, sl.err);
  return TJB_OK;
}

// ---- posterior (host rows) -----------------------------------------------------------

int tjb_posterior_aA(TjbHandle *h, const double *h_rows, int64_t k, int clamp_K, double *h_ll,
                     double *h_a, double *h_A) {
  return posterior_common(h, h_rows, k, clamp_K, 0, nullptr, h_ll, h_a, h_A, nullptr);
}

int tjb_posterior_draw(TjbHandle *h, const double *h_rows, int64_t k, int n_per, int clamp_K,
                       const double *h_normals, double *h_out, double *h_ll) {
  if (n_per < 1) return fail(TJB_E_INVALID, "n_per must be >= 1");
  if (k > 0 && (!h_normals || !h_out)) return fail(TJB_E_INVALID, "null normals / output");
  return posterior_common(h, h_rows, k, clamp_K, n_per, h_normals, h_ll, nullptr, nullptr, h_out);
}

int tjb_design_column(TjbHandle *h, const double *h_row, double *h_z, int32_t *h_stats) {
  if (!h || !h_row || !h_z) return fail(TJB_E_INVALID, "null argument");
  CU(cudaSetDevice(h->device));
  const int N = h->N;
  if (h->misc.ensure((size_t)N * 16 + 64)) return fail(TJB_E_NOMEM, "cudaMalloc");
  double *d_dt = (double *)h->misc.p, *d_z = d_dt + N;
  int *d_st = (int *)(d_z + N);
  std::vector<double> dt(N);
  for (int n = 0; n < N; n++) dt[n] = h->star.t[n] - h->star.t_ref;
  CU(cudaMemcpyAsync(d_dt, dt.data(), (size_t)N * 8, cudaMemcpyHostToDevice, h->stream));
  design_column_kernel<<<1, 32, 0, h->stream>>>(d_dt, N, h_row[0], h_row[1], h_row[2], h_row[3], 0.0,
                                              (const SinCos *)h->trig, d_z,
                                              d_st);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(h_z, d_z, (size_t)N * 8, cudaMemcpyDeviceToHost, h->stream));
  int st[3] = {0, 0, 0};
  CU(cudaMemcpyAsync(st, d_st, sizeof(st), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  if (h_stats) { h_stats[0] = st[0]; h_stats[1] = st[1]; h_stats[2] = st[2]; }
  return TJB_OK;
}

int tjb_unmarginalized_ll(TjbHandle *h, const double *h_rows, int64_t n, double *h_ll) {
  if (!h) return fail(TJB_E_INVALID, "null handle");
  if (n < 0) return fail(TJB_E_INVALID, "negative n");
  if (n == 0) return TJB_OK;
  if (!h_rows || !h_ll) return fail(TJB_E_INVALID, "null host pointer");
  CU(cudaSetDevice(h->device));
  const int N = h->N, L = h->L, W = 5 + L;
  if (h->misc.ensure((size_t)n * (W + 1) * sizeof(double))) return fail(TJB_E_NOMEM, "cudaMalloc");
  double *d_rows = (double *)h->misc.p, *d_ll = d_rows + (size_t)n * W;
  CU(cudaMemcpyAsync(d_rows, h_rows, (size_t)n * W * sizeof(double), cudaMemcpyHostToDevice,
                     h->stream));
  const int threads = 128;
  const int grid = (int)std::min<long long>((n + threads - 1) / threads, (long long)h->n_sm * 16);
  unmarginalized_ll_kernel<<<grid, threads, 0, h->stream>>>(
      (const double *)h->tab_jit.p, N, L, row_stride(L),
      h->star.centred ? (double)h->star.centre : 0.0, d_rows, n, 0.0, (const SinCos *)h->trig,
      d_ll);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(h_ll, d_ll, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return TJB_OK;
}

int tjb_set_epoch_rows_mode(int mode) {
  if (mode < 0 || mode > 1) return fail(TJB_E_INVALID, "mode must be 0 or 1");
  g_force_shared_rows.store(mode, std::memory_order_relaxed);
  return TJB_OK;
}

int tjb_fp64_peak(TjbHandle *h, int iters, double *h_tflops, double *h_ms) {
  if (!h || !h_tflops) return fail(TJB_E_INVALID, "null argument");
  CU(cudaSetDevice(h->device));
  if (h->misc.ensure(64)) return fail(TJB_E_NOMEM, "cudaMalloc");
  const int grid = h->n_sm * 8;
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  fp64_peak_kernel<<<grid, 256, 0, h->stream>>>(iters / 8 + 1, (double *)h->misc.p);  // warm-up
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    CU(cudaEventRecord(e0, h->stream));
    fp64_peak_kernel<<<grid, 256, 0, h->stream>>>(iters, (double *)h->misc.p);
    CU(cudaEventRecord(e1, h->stream));
    CU(cudaEventSynchronize(e1));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    best = std::min(best, ms);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  const double flops = 2.0 * 8.0 * (double)iters * 256.0 * (double)grid;
  *h_tflops = flops / (best * 1e-3) / 1e12;
  if (h_ms) *h_ms = best;
  return TJB_OK;
}

}  // extern "C"
