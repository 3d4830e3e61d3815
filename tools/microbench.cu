// microbench.cu -- issue-slot model of the sm_100a FP64 pipe: can FP32 / INT / MUFU
// instructions be issued in the shadow of DFMAs?  Build: nvcc -arch=sm_100a -O3 -o
// build/microbench tools/microbench.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

template <int ND, int NF, int NI, int NM>
__global__ void __launch_bounds__(256) mix(int iters, double *out, float fseed, int iseed) {
  double d[8];
  float f[8];
  int q[8];
  float m[4];
#pragma unroll
  for (int i = 0; i < 8; i++) { d[i] = threadIdx.x * 1e-9 + i; f[i] = fseed + i; q[i] = iseed + i; }
#pragma unroll
  for (int i = 0; i < 4; i++) m[i] = fseed + 0.1f * i;
  const double dm = 0.999999, dc = 1e-7;
  const float fm = 0.99999f, fc = 1e-6f;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
      for (int i = 0; i < ND; i++) d[i] = fma(d[i], dm, dc);
#pragma unroll
      for (int i = 0; i < NF; i++) f[i] = fmaf(f[i], fm, fc);
#pragma unroll
      for (int i = 0; i < NI; i++) q[i] = q[i] * 3 + iseed;
#pragma unroll
      for (int i = 0; i < NM; i++) m[i] = __sinf(m[i]);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += d[i] + f[i] + q[i];
#pragma unroll
  for (int i = 0; i < 4; i++) s += m[i];
  if (s == 12345.678) out[0] = s;
}

template <int ND, int NF, int NI, int NM>
void run(const char *name, int nsm) {
  double *out;
  cudaMalloc(&out, 8);
  const int iters = 4000, grid = nsm * 8;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  mix<ND, NF, NI, NM><<<grid, 256>>>(100, out, 1.0f, 1);
  cudaEventRecord(e0);
  mix<ND, NF, NI, NM><<<grid, 256>>>(iters, out, 1.0f, 1);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  // cycles per SMSP per "round" (one r-iteration) per warp: 16 warps per SMSP resident
  int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  double cycles = ms * 1e-3 * clk_khz * 1e3;
  double rounds_per_smsp = (double)iters * 4 * (8.0 * 256 / 32 / 4);  // warps per SMSP x rounds
  printf("%-28s ND=%d NF=%d NI=%d NM=%d  %.3f ms  %.2f cycles/round/warp (model max(2ND,sum)=%d, 2ND+rest=%d)\n",
         name, ND, NF, NI, NM, ms, cycles / rounds_per_smsp, (2 * ND > ND + NF + NI + NM ? 2 * ND : ND + NF + NI + NM),
         2 * ND + NF + NI + NM);
  cudaFree(out);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  printf("%s, %d SMs, clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  int n = p.multiProcessorCount;
  run<8, 0, 0, 0>("dfma only", n);
  run<4, 0, 0, 0>("dfma only (4 chains)", n);
  run<8, 4, 0, 0>("dfma + 0.5 ffma", n);
  run<8, 8, 0, 0>("dfma + 1.0 ffma", n);
  run<4, 8, 0, 0>("dfma + 2.0 ffma", n);
  run<8, 0, 8, 0>("dfma + 1.0 imad", n);
  run<8, 4, 4, 0>("dfma + .5 ffma + .5 imad", n);
  run<8, 0, 0, 1>("dfma + 1/8 mufu", n);
  run<8, 0, 0, 2>("dfma + 2/8 mufu", n);
  run<8, 4, 2, 1>("dfma + mix", n);
  run<0, 8, 0, 0>("ffma only", n);
  run<0, 0, 0, 4>("mufu only", n);
  return 0;
}
