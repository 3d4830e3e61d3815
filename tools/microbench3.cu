// microbench3.cu -- accuracy of the MUFU.RCP64H seed and of the refinements used in
// kepler.cuh::rcp_pos (decides how many FMA steps the reciprocal needs).
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>

__global__ void k(int n, double lo, double hi, double *maxerr) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x = lo + (hi - lo) * ((i + 0.37) / n);
  double r0;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(x));
  double t = fma(-x, r0, 1.0);
  // (a) third-order single step: r (1 + t + t^2)
  double ra = fma(r0, fma(t, t, t), r0);
  // (b) two Newton steps
  double rb = fma(r0, t, r0);
  double t2 = fma(-x, rb, 1.0);
  rb = fma(rb, t2, rb);
  // (c) third-order + residual correction (4 ops)
  double tc = fma(-x, ra, 1.0);
  double rc = fma(ra, tc, ra);
  double exact = 1.0 / x;
  double e0 = fabs(r0 * x - 1.0);
  double ea = fabs(ra - exact) / exact, eb = fabs(rb - exact) / exact, ec = fabs(rc - exact) / exact;
  // atomic max via integer compare on non-negative doubles
  atomicMax((unsigned long long *)&maxerr[0], __double_as_longlong(e0));
  atomicMax((unsigned long long *)&maxerr[1], __double_as_longlong(ea));
  atomicMax((unsigned long long *)&maxerr[2], __double_as_longlong(eb));
  atomicMax((unsigned long long *)&maxerr[3], __double_as_longlong(ec));
}

int main() {
  double *d, h[4];
  cudaMalloc(&d, 32);
  const double ranges[4][2] = {{1e-3, 2.0}, {0.5, 1.0}, {1.0, 2.0}, {1e-6, 1e-3}};
  for (auto &rg : ranges) {
    cudaMemset(d, 0, 32);
    int n = 1 << 24;
    k<<<(n + 255) / 256, 256>>>(n, rg[0], rg[1], d);
    cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
    printf("x in [%g, %g]: seed |x r0 - 1| max %.3e (%.1f bits); 3rd-order 1 step %.3e; 2 Newton %.3e; 3rd+1 %.3e\n",
           rg[0], rg[1], h[0], -log2(h[0]), h[1], h[2], h[3]);
  }
  return 0;
}
