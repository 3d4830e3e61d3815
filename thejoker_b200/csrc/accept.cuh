// accept.cuh -- rejection accept step on the device.
//
// Replaces, for lls resident on the GPU,
//     uu = rng.uniform(size=len(lls))
//     good = np.where(np.exp(lls - lls.max()) > uu)[0][:max_posterior_samples]
// (thejoker/likelihood_helpers.py:107-109, 183-185; multiproc_helpers.py:256-258,
// 373-375).  The max is produced by the likelihood kernel as an order-preserving
// int64 key (marginal_ll.cuh) and, across GPUs, by an integer MAX all-reduce.
//
// Three kernels:
//   accept_flag_kernel     exp(ll - max) > u per sample -> one bit per sample
//                          (warp ballot), per-CTA accepted counts, totals
//   accept_scan_kernel     exclusive scan of the per-CTA counts (one CTA)
//   accept_scatter_kernel  ordered expansion of the bit mask into ascending
//                          int64 indices, truncated at max_keep
// The uniforms are either read from a device array or generated in place by a
// leapfrogged PCG64 that reproduces numpy's Generator(PCG64).random() stream bit
// for bit (pcg64 section below), so the n uniforms never exist on the host.
#pragma once

#include "marginal_ll.cuh"

namespace tjb {

// ---------------------------------------------------------------------------
// PCG64 (XSL-RR 128/64), numpy's default BitGenerator.  next():
//     state = state * MULT + inc;  out = rotr64(hi ^ lo, hi >> 58)
// random() = (out >> 11) * 2^-53.
typedef unsigned __int128 u128;

struct Lcg128 {  // affine map x -> mult * x + plus over Z/2^128
  u128 mult, plus;
};

TJB_HD u128 make_u128(uint64_t hi, uint64_t lo) { return ((u128)hi << 64) | lo; }
TJB_HD u128 pcg_mult() { return make_u128(2549297995355413924ULL, 4865540595714422341ULL); }

// the map that advances the generator by `delta` steps
TJB_HD Lcg128 lcg_power(u128 inc, uint64_t delta) {
  Lcg128 acc = {1, 0};
  u128 cm = pcg_mult(), cp = inc;
  while (delta > 0) {
    if (delta & 1) {
      acc.mult *= cm;
      acc.plus = acc.plus * cm + cp;
    }
    cp = (cm + 1) * cp;
    cm *= cm;
    delta >>= 1;
  }
  return acc;
}

TJB_HD double pcg_output_double(u128 state) {
  const uint64_t hi = (uint64_t)(state >> 64), lo = (uint64_t)state;
  const uint64_t x = hi ^ lo;
  const unsigned rot = (unsigned)(hi >> 58);
  const uint64_t out = (x >> rot) | (x << ((64 - rot) & 63));
  return (double)(out >> 11) * (1.0 / 9007199254740992.0);
}

#if defined(__CUDACC__)

struct PcgParams {
  u128 state, inc;   // generator state before the first requested output
  Lcg128 stride;     // advance by (total threads) steps
  int enabled;
};

constexpr int kAccThreads = 256;
// Mask words (32 samples each) owned by one CTA of the flag and scatter kernels: a power
// of two in [kAccMinWordsPerCta, kAccMaxWordsPerCta] chosen per call so that small arrays
// (one star of a multi-star batch, an early round of the iterative sampler) still fill
// the GPU, while large ones amortise the per-thread PCG jump over up to 256 samples.
constexpr int kAccMaxWordsPerCta = 2048;  // 65536 samples per CTA
constexpr int kAccMinWordsPerCta = 256;   // 8192 samples per CTA: one mask word per lane in the scatter
inline int acc_words_per_cta(long long n_words, int n_sm) {
  int w = kAccMaxWordsPerCta;
  while (w > kAccMinWordsPerCta && n_words / w < 4LL * n_sm) w >>= 1;
  return w;
}

// out[i] = the (offset+i)-th uniform; grid-stride leapfrog
__global__ void __launch_bounds__(kAccThreads)
pcg64_uniform_kernel(const PcgParams pp, const long long n, double *__restrict__ out) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long T = (long long)gridDim.x * blockDim.x;
  if (gid >= n) return;
  const Lcg128 j = lcg_power(pp.inc, (uint64_t)gid + 1);
  u128 st = j.mult * pp.state + j.plus;
  for (long long i = gid; i < n; i += T) {
    out[i] = pcg_output_double(st);
    st = pp.stride.mult * st + pp.stride.plus;
  }
}

// CTA b owns mask words [b*W, (b+1)*W), W = words_per_cta; warp j of the CTA
// visits words b*W + j, b*W + j + 8, ... so that a thread's samples are 256
// apart: the PCG leapfrog stride inside a CTA is the constant blockDim.
__global__ void __launch_bounds__(kAccThreads)
accept_flag_kernel(const double *__restrict__ ll, const long long n,
                   const long long *__restrict__ llmax_key, const double *__restrict__ uniforms,
                   const PcgParams pp, const double near_tol, const int words_per_cta,
                   unsigned *__restrict__ mask, unsigned *__restrict__ cta_counts,
                   unsigned long long *__restrict__ totals) {
  const double llmax = key_to_ll(*llmax_key);
  const long long first = (long long)blockIdx.x * words_per_cta * 32 + threadIdx.x;
  const long long cta_end = min(n, ((long long)blockIdx.x + 1) * words_per_cta * 32);
  const long long n_words = (n + 31) / 32;
  // PCG64 state of this thread's first sample: a two-level jump.  One thread takes the
  // generator to the CTA's first sample (O(log n) 128-bit squarings, ~3000 instructions),
  // every thread then advances that state by its own index within the CTA (<= 8 squarings).
  // With the full jump in every thread a CTA of the smallest size (8192 samples: one star of
  // a multi-star batch) spent more instructions on jumping than on its samples.
  u128 st = 0;
  if (pp.enabled) {
    __shared__ u128 s_cta_state;
    if (threadIdx.x == 0) {
      const Lcg128 j = lcg_power(pp.inc, (uint64_t)((long long)blockIdx.x * words_per_cta * 32) + 1);
      s_cta_state = j.mult * pp.state + j.plus;
    }
    __syncthreads();
    const Lcg128 jt = lcg_power(pp.inc, (uint64_t)threadIdx.x);
    st = jt.mult * s_cta_state + jt.plus;
  }
  unsigned cnt = 0, near = 0, nonfin = 0;
  const long long cta_end_round = ((cta_end + 31) / 32) * 32;
  for (long long i = first; i < cta_end_round; i += kAccThreads) {
    bool acc = false;
    if (i < n) {
      const double u = pp.enabled ? pcg_output_double(st) : uniforms[i];
      const double v = ll[i];
      const double a = exp(v - llmax);
      acc = a > u;
      near += (fabs(a - u) <= near_tol) ? 1u : 0u;
      // NaN / +-inf lls (likelihood_helpers.py:173-176 checks np.isfinite over all of them;
      // the max key alone does not see a -inf)
      nonfin += (fabs(v) <= 1.79769313486231570815e+308) ? 0u : 1u;
    }
    if (pp.enabled) st = pp.stride.mult * st + pp.stride.plus;
    const unsigned word = __ballot_sync(0xffffffffu, acc);
    if ((threadIdx.x & 31) == 0 && (i >> 5) < n_words) {
      mask[i >> 5] = word;
      cnt += __popc(word);
    }
  }
  // CTA totals
  __shared__ unsigned s_cnt, s_near, s_nonfin;
  if (threadIdx.x == 0) { s_cnt = 0; s_near = 0; s_nonfin = 0; }
  __syncthreads();
  if (cnt) atomicAdd(&s_cnt, cnt);
  if (near) atomicAdd(&s_near, near);
  if (nonfin) atomicAdd(&s_nonfin, nonfin);
  __syncthreads();
  if (threadIdx.x == 0) {
    cta_counts[blockIdx.x] = s_cnt;
    if (s_cnt) atomicAdd(&totals[0], (unsigned long long)s_cnt);
    if (s_near) atomicAdd(&totals[1], (unsigned long long)s_near);
    if (s_nonfin) atomicAdd(&totals[2], (unsigned long long)s_nonfin);
  }
}

// exclusive scan of cta_counts[0..m) into cta_offsets (64-bit), single CTA
__global__ void __launch_bounds__(1024)
accept_scan_kernel(const unsigned *__restrict__ cta_counts, const int m,
                   unsigned long long *__restrict__ cta_offsets) {
  __shared__ unsigned long long warp_tot[32];
  __shared__ unsigned long long carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < m; base += 1024) {
    const int i = base + threadIdx.x;
    unsigned long long v = i < m ? cta_counts[i] : 0, x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long y = __shfl_up_sync(0xffffffffu, x, o);
      if ((threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      unsigned long long t = warp_tot[threadIdx.x], xs = t;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long y = __shfl_up_sync(0xffffffffu, xs, o);
        if (threadIdx.x >= o) xs += y;
      }
      warp_tot[threadIdx.x] = xs - t;  // exclusive
    }
    __syncthreads();
    const unsigned long long excl = carry + warp_tot[threadIdx.x >> 5] + (x - v);
    if (i < m) cta_offsets[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + v;
    __syncthreads();
  }
}

// CTA b expands its mask words in ascending order.  Warp j takes the contiguous
// word range [j*W/8, (j+1)*W/8) of the CTA (W = words_per_cta), lane l word (range start + 32 it + l).
__global__ void __launch_bounds__(kAccThreads)
accept_scatter_kernel(const unsigned *__restrict__ mask, const long long n,
                      const unsigned long long *__restrict__ cta_offsets,
                      const long long index_base, const long long max_keep,
                      const int words_per_cta, long long *__restrict__ idx_out) {
  constexpr int kWarps = kAccThreads / 32;
  const int kWordsPerWarp = words_per_cta / kWarps;  // a multiple of 32
  const long long n_words = (n + 31) / 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long w0 = (long long)blockIdx.x * words_per_cta + (long long)warp * kWordsPerWarp;
  const unsigned long long cta_off = cta_offsets[blockIdx.x];
  if ((long long)cta_off >= max_keep) return;

  // pass 1: accepted count of this warp's range
  unsigned wcnt = 0;
  for (int it = 0; it < kWordsPerWarp; it += 32) {
    const long long w = w0 + it + lane;
    wcnt += (w < n_words) ? __popc(mask[w]) : 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) wcnt += __shfl_xor_sync(0xffffffffu, wcnt, o);
  __shared__ unsigned s_w[kWarps];
  if (lane == 0) s_w[warp] = wcnt;
  __syncthreads();
  unsigned long long pos = cta_off;
  for (int j = 0; j < warp; j++) pos += s_w[j];

  // pass 2: ordered expansion
  for (int it = 0; it < kWordsPerWarp; it += 32) {
    const long long w = w0 + it + lane;
    unsigned word = (w < n_words) ? mask[w] : 0u;
    const unsigned c = __popc(word);
    unsigned x = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    unsigned long long p = pos + (x - c);
    while (word) {
      const int b = __ffs(word) - 1;
      word &= word - 1;
      if ((long long)p < max_keep) idx_out[p] = index_base + w * 32 + b;
      p++;
    }
    pos += __shfl_sync(0xffffffffu, x, 31);
  }
}

// max-update a key with an ll array that did not come from the likelihood kernel
__global__ void __launch_bounds__(256)
llmax_update_kernel(const double *__restrict__ ll, const long long n, long long *llmax_key) {
  long long kmax = ll_to_key(-INFINITY);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const long long k = ll_to_key(ll[i]);
    kmax = k > kmax ? k : kmax;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const long long other = __shfl_xor_sync(0xffffffffu, kmax, o);
    kmax = other > kmax ? other : kmax;
  }
  if ((threadIdx.x & 31) == 0) atomicMax(llmax_key, kmax);
}

#endif  // __CUDACC__

}  // namespace tjb
