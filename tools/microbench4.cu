// microbench4.cu -- issue cost of FP64 instructions with a uniform-register operand, and of
// packed FP32 (FFMA2) next to FP64, on one SM sub-partition (4 warps per scheduler x 2 CTAs).
// Same harness as microbench2.cu: cycles per round (8 FP64 instructions + extras) per warp.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long pk(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}

template <int MODE>
__global__ void __launch_bounds__(256) k(int iters, double *out, double seed, float fseed,
                                         double u0, double u1, double u2, double u3) {
  double d[12];
  float f[12];
  unsigned long long p[6];
#pragma unroll
  for (int i = 0; i < 12; i++) { d[i] = seed + threadIdx.x * 1e-9 + i * 0.01; f[i] = fseed + i * 0.01f; }
#pragma unroll
  for (int i = 0; i < 6; i++) p[i] = pk(f[2 * i], f[2 * i + 1]);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
      if (MODE == 0) {  // 8 DFMA, three distinct vector operands
#pragma unroll
        for (int i = 0; i < 8; i++) d[i] = fma(d[i], d[(i + 3) % 12], d[(i + 7) % 12]);
      } else if (MODE == 1) {  // 8 DFMA, multiplier from a kernel parameter (uniform register)
#pragma unroll
        for (int i = 0; i < 8; i++) d[i] = fma(d[i], (i & 1) ? u0 : u1, d[(i + 7) % 12]);
      } else if (MODE == 2) {  // 8 DFMA, addend from a kernel parameter
#pragma unroll
        for (int i = 0; i < 8; i++) d[i] = fma(d[i], d[(i + 3) % 12], (i & 1) ? u2 : u3);
      } else if (MODE == 3) {  // 8 DFMA distinct + 8 FFMA distinct
#pragma unroll
        for (int i = 0; i < 8; i++) {
          d[i] = fma(d[i], d[(i + 3) % 12], d[(i + 7) % 12]);
          f[i] = fmaf(f[i], f[(i + 3) % 12], f[(i + 7) % 12]);
        }
      } else if (MODE == 4) {  // 8 DFMA distinct + 4 FFMA2 distinct (the same FP32 work)
#pragma unroll
        for (int i = 0; i < 8; i++) {
          d[i] = fma(d[i], d[(i + 3) % 12], d[(i + 7) % 12]);
          if (i < 4)
            asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(p[(i + 1) % 6]), "l"(p[(i + 3) % 6]));
        }
      } else if (MODE == 5) {  // 8 DFMA (uniform multiplier) + 8 FFMA distinct
#pragma unroll
        for (int i = 0; i < 8; i++) {
          d[i] = fma(d[i], (i & 1) ? u0 : u1, d[(i + 7) % 12]);
          f[i] = fmaf(f[i], f[(i + 3) % 12], f[(i + 7) % 12]);
        }
      } else if (MODE == 6) {  // 8 DFMA (uniform multiplier) + 4 FFMA2
#pragma unroll
        for (int i = 0; i < 8; i++) {
          d[i] = fma(d[i], (i & 1) ? u0 : u1, d[(i + 7) % 12]);
          if (i < 4)
            asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(p[(i + 1) % 6]), "l"(p[(i + 3) % 6]));
        }
      } else if (MODE == 7) {  // 8 DMUL by a uniform register
#pragma unroll
        for (int i = 0; i < 8; i++) d[i] = d[i] * ((i & 1) ? u0 : u1);
      } else if (MODE == 8) {  // 8 DFMA distinct + 8 MUFU.SIN (XU pipe next to FP64)
#pragma unroll
        for (int i = 0; i < 8; i++) {
          d[i] = fma(d[i], d[(i + 3) % 12], d[(i + 7) % 12]);
          f[i] = __sinf(f[i]);
        }
      } else if (MODE >= 10 && MODE <= 13) {  // (MODE - 10) x 8 DFMA distinct + 8 MUFU.RSQ
#pragma unroll
        for (int rep = 0; rep < MODE - 10; rep++)
#pragma unroll
          for (int i = 0; i < 8; i++) d[i] = fma(d[i], d[(i + 3) % 12], d[(i + 7) % 12]);
#pragma unroll
        for (int i = 0; i < 8; i++) asm("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
      } else if (MODE == 14) {  // 8 DFMA distinct + 8 RCP64H
#pragma unroll
        for (int i = 0; i < 8; i++) {
          d[i] = fma(d[i], d[(i + 3) % 12], d[(i + 7) % 12]);
          if (i < 4) asm("rcp.approx.ftz.f64 %0, %0;" : "+d"(d[8 + i]));
        }
#pragma unroll
        for (int i = 0; i < 4; i++) asm("rcp.approx.ftz.f64 %0, %0;" : "+d"(d[8 + i]));
      } else if (MODE == 15) {  // 8 DFMA distinct + 8 F2F.F64.F32 (results folded into FP32 adds)
#pragma unroll
        for (int i = 0; i < 8; i++) {
          d[i] = fma(d[i], d[(i + 3) % 12], d[(i + 7) % 12]);
          double t = (double)f[i];
          f[i] = __int_as_float(__double2loint(t) ^ __float_as_int(f[i]));
        }
      } else if (MODE == 16) {  // 8 DFMA distinct + 8 I2FP.F32.S32
#pragma unroll
        for (int i = 0; i < 8; i++) {
          d[i] = fma(d[i], d[(i + 3) % 12], d[(i + 7) % 12]);
          f[i] = (float)__float_as_int(f[i]);
        }
      } else if (MODE == 17) {  // 8 DFMA distinct + 8 LDS.128 with a lane-dependent (random) address
        extern __shared__ double2 sm[];
#pragma unroll
        for (int i = 0; i < 8; i++) {
          d[i] = fma(d[i], d[(i + 3) % 12], d[(i + 7) % 12]);
          const double2 v = sm[(__double2loint(d[i]) * 2654435761u >> 21) & 2047];
          d[8 + (i & 3)] += v.x * v.y;
        }
      } else if (MODE == 9) {  // 8 FFMA2 only
#pragma unroll
        for (int i = 0; i < 8; i++)
          asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i % 6]) : "l"(p[(i + 1) % 6]), "l"(p[(i + 3) % 6]));
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) s += d[i] + f[i];
#pragma unroll
  for (int i = 0; i < 6; i++) s += (double)p[i];
  if (s == 12345.678) out[0] = s;
}

template <int MODE>
void run(const char *name, int nsm) {
  double *out;
  cudaMalloc(&out, 8);
  const int iters = 4000, grid = nsm * 8;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<grid, 256, 32768>>>(100, out, 1e-3, 1e-3f, 0.999, 1.001, 1e-9, -1e-9);
  cudaEventRecord(e0);
  k<MODE><<<grid, 256, 32768>>>(iters, out, 1e-3, 1e-3f, 0.999, 1.001, 1e-9, -1e-9);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  double cycles = ms * 1e-3 * clk_khz * 1e3;
  double rounds_per_smsp = (double)iters * 4 * 16;
  printf("%-52s %.3f ms  %.2f cycles/round/warp\n", name, ms, cycles / rounds_per_smsp);
  cudaFree(out);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int n = p.multiProcessorCount;
  run<0>("8 DFMA distinct operands", n);
  run<1>("8 DFMA, multiplier in a uniform register", n);
  run<2>("8 DFMA, addend in a uniform register", n);
  run<7>("8 DMUL by a uniform register", n);
  run<3>("8 DFMA distinct + 8 FFMA distinct", n);
  run<4>("8 DFMA distinct + 4 FFMA2", n);
  run<5>("8 DFMA uniform-mult + 8 FFMA distinct", n);
  run<6>("8 DFMA uniform-mult + 4 FFMA2", n);
  run<8>("8 DFMA distinct + 8 MUFU.SIN", n);
  run<9>("8 FFMA2", n);
  run<10>("8 MUFU.RSQ", n);
  run<11>("8 DFMA distinct + 8 MUFU.RSQ", n);
  run<12>("16 DFMA distinct + 8 MUFU.RSQ", n);
  run<13>("24 DFMA distinct + 8 MUFU.RSQ", n);
  run<14>("8 DFMA distinct + 8 RCP64H", n);
  run<15>("8 DFMA distinct + 8 F2F.F64.F32 (+8 LOP3)", n);
  run<16>("8 DFMA distinct + 8 I2FP.F32.S32", n);
  run<17>("8 DFMA distinct + 8 LDS.128 random (+8 DFMA, IMAD..)", n);
  return 0;
}
