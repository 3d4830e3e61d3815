// microbench2.cu -- register-file pressure model: DFMA / FFMA with three distinct,
// non-reusable register operands; FP64 op classes; conversions.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(int iters, double *out, double seed, float fseed) {
  double d[12];
  float f[12];
#pragma unroll
  for (int i = 0; i < 12; i++) { d[i] = seed + threadIdx.x * 1e-9 + i * 0.01; f[i] = fseed + i * 0.01f; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
      if (MODE == 0) {  // 8 DFMA, distinct operands, no reuse
#pragma unroll
        for (int i = 0; i < 8; i++) d[i] = fma(d[i], d[(i + 3) % 12], d[(i + 7) % 12]);
      } else if (MODE == 1) {  // 8 DFMA + 8 FFMA, all distinct operands
#pragma unroll
        for (int i = 0; i < 8; i++) {
          d[i] = fma(d[i], d[(i + 3) % 12], d[(i + 7) % 12]);
          f[i] = fmaf(f[i], f[(i + 3) % 12], f[(i + 7) % 12]);
        }
      } else if (MODE == 2) {  // 8 DMUL
#pragma unroll
        for (int i = 0; i < 8; i++) d[i] = d[i] * d[(i + 3) % 12];
      } else if (MODE == 3) {  // 8 DADD
#pragma unroll
        for (int i = 0; i < 8; i++) d[i] = d[i] + d[(i + 3) % 12];
      } else if (MODE == 4) {  // 8 DFMA + 2 F2F round trips
#pragma unroll
        for (int i = 0; i < 8; i++) d[i] = fma(d[i], d[(i + 3) % 12], d[(i + 7) % 12]);
        f[0] = (float)d[8]; d[9] = (double)f[1];
        f[1] = f[0] * 1.0001f;
      } else if (MODE == 5) {  // 8 DFMA + 2 RCP64H
#pragma unroll
        for (int i = 0; i < 8; i++) d[i] = fma(d[i], d[(i + 3) % 12], d[(i + 7) % 12]);
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(d[8]) : "d"(d[9]));
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(d[9]) : "d"(d[10]));
      } else if (MODE == 6) {  // 8 FFMA distinct only
#pragma unroll
        for (int i = 0; i < 8; i++) f[i] = fmaf(f[i], f[(i + 3) % 12], f[(i + 7) % 12]);
      } else if (MODE == 7) {  // 8 DFMA + 8 FMUL-by-immediate (1 reg operand)
#pragma unroll
        for (int i = 0; i < 8; i++) {
          d[i] = fma(d[i], d[(i + 3) % 12], d[(i + 7) % 12]);
          f[i] = f[i] * 1.0001f;
        }
      } else if (MODE == 8) {  // 8 DFMA + 8 LOP3 (int)
#pragma unroll
        for (int i = 0; i < 8; i++) {
          d[i] = fma(d[i], d[(i + 3) % 12], d[(i + 7) % 12]);
          f[i] = __int_as_float(__float_as_int(f[i]) ^ __float_as_int(f[(i + 3) % 12]));
        }
      } else if (MODE == 9) {  // 8 DFMA with one operand shared by all (reuse)
#pragma unroll
        for (int i = 0; i < 8; i++) d[i] = fma(d[i], d[10], d[11]);
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) s += d[i] + f[i];
  if (s == 12345.678) out[0] = s;
}

template <int MODE>
void run(const char *name, int nsm) {
  double *out;
  cudaMalloc(&out, 8);
  const int iters = 4000, grid = nsm * 8;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<grid, 256>>>(100, out, 1e-3, 1e-3f);
  cudaEventRecord(e0);
  k<MODE><<<grid, 256>>>(iters, out, 1e-3, 1e-3f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  double cycles = ms * 1e-3 * clk_khz * 1e3;
  double rounds_per_smsp = (double)iters * 4 * 16;
  printf("%-44s %.3f ms  %.2f cycles/round/warp\n", name, ms, cycles / rounds_per_smsp);
  cudaFree(out);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int n = p.multiProcessorCount;
  run<0>("8 DFMA distinct operands", n);
  run<9>("8 DFMA shared operands (reuse)", n);
  run<1>("8 DFMA + 8 FFMA distinct", n);
  run<7>("8 DFMA + 8 FMUL imm", n);
  run<8>("8 DFMA + 8 LOP3", n);
  run<2>("8 DMUL", n);
  run<3>("8 DADD", n);
  run<4>("8 DFMA + 2 F2F", n);
  run<5>("8 DFMA + 2 RCP64H", n);
  run<6>("8 FFMA distinct", n);
  return 0;
}
