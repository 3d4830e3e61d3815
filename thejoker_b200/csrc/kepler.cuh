// kepler.cuh -- fixed-work Kepler solver and unit-amplitude RV column for one
// (prior sample, epoch), written for the sm_100a FP64 pipe.
//
// Replaces twobody.c::c_rv_from_elements as called from
// thejoker/src/fast_likelihood.pyx:453-455, 511-513, 561-563 (K = 1):
//     M   = 2 pi (t - t0) / P - M0
//     E   : E - e sin E = M
//     z   = cos(f + omega) + e cos(omega),  f = true anomaly
//
// Design.  The kernel loads four units at once -- FP64 issue (one
// warp instruction every 2 cycles per scheduler, 3 when it reads three distinct vector
// registers), the issue slots, the XU pipe (MUFU, conversions) and shared memory -- so
// what counts is the instruction count per epoch on every one of them.  Hence:
//  * angles are carried in binary fractions of a revolution (1/2048, or 1/4 for the
//    polynomial back-end) so every argument reduction is an exact magic-number rounding
//    (no Payne-Hanek, no multiples of 2 pi);
//  * the starter and one Householder refinement run on the FP32 / MUFU pipes
//    (sin.approx, cos.approx, rsqrt.approx, rcp.approx) with no branches, from a phase
//    reduced in fixed point;
//  * the FP64 stage starts from that estimate snapped to a grid whose sin / cos are the
//    product of two table nodes (TJB_TRIG2), takes one second-order (Halley) step
//    (error ~1e-6 -> below 1e-16) and carries sin E / cos E through it by an
//    angle-addition update: no double-precision polynomial on the main path;
//  * constants in the multiplier slot come from the constant bank as uniform-register
//    operands (TJB_UCONST), the others are pinned in registers for the whole epoch loop
//    (ptxas otherwise re-loads or re-materialises them every epoch);
//  * z is formed without atan2:  cos f = (cosE - e)/(1 - e cosE),
//    sin f = sqrt(1-e^2) sinE/(1 - e cosE);
//  * the FP64 step is closed by a warp-uniform convergence test (|delta|), so
//    high-eccentricity lanes near pericentre cost extra passes only in their warp, in a
//    safeguarded third-order iteration off the main path (solve_extra_passes).
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define TJB_HD __host__ __device__ __forceinline__
#define TJB_D __device__ __forceinline__
#define TJB_HD_NOINLINE __host__ __device__ __noinline__
#else
#define TJB_HD inline
#define TJB_D inline
#define TJB_HD_NOINLINE inline
#endif

namespace tjb {

// Trig back-end of the FP64 stage (both measured on B200):
//   0  angles in quarter-revolutions; sin/cos by degree-13/14 minimax polynomials and a
//      quadrant swap (15 FP64 + ~10 integer instructions per evaluation)
//   1  angles in 1/2^TJB_TRIG_TABLE_LOG2 revolution (1/2048: a 32 KB table); sin/cos of the
//      nearest table node from shared memory, rotated by the residual |r| <= pi/2048 through
//      degree-3/4 polynomials (9 FP64 instructions + one LDS.128, no quadrant logic).  With
//      TJB_TRIG2 (below) the main path needs no polynomial at all; this evaluation then
//      serves the per-sample set-up and the rare path.
#ifndef TJB_TRIG_TABLE
#define TJB_TRIG_TABLE 1
#endif

// Instruction-trimmed variant of the epoch loop (the default since round 2: timed on B200,
// profiles/r02a_tune_variants.jsonl -- together with TJB_PHASE_FIXED, TJB_HALLEY, a
// 2048-node table and 3 epochs per iteration 3.07e9 -> 3.73e9 samples/s at N = 64;
// tests/test_host_logic.py checks the numerics of every combination):
//   * e cosE / (6 f1) = (1/f1 - 1)/6: one FMA instead of two multiplies, FP64 and FP32 stage
//   * FP32 stage: t/2 computed once
//   * convergence vote on one max of the sign-stripped high words (per-epoch flags are
//     recomputed on the rare path only)
//   * 0.5 pinned in a register pair (ptxas otherwise materialises it per iteration)
// Saves ~1 FP64 and ~4.5 other instructions per (sample, epoch).
#ifndef TJB_TRIM
#define TJB_TRIM 1
#endif

// Phase reduction of the FP32 stage in fixed point (default on since round 2).
// The shipped loop rounds the unreduced phase x4 to float first, so its starter carries an
// error of ulp_float(x4)/2 -- 3.8e-4 rad at 2000 revolutions (P = 2 d over a 4000 d
// baseline), above the 2^-13 threshold of the one-pass FP64 step: such lanes send their
// warp through the extra-pass path (64 % of the warps at least once over 64 epochs on a
// 4000 d baseline, 5 % on the 155 d benchmark data; tests/test_host_logic.py::
// test_phase_reduction_variants).  With TJB_PHASE_FIXED one FP64 add of 1.5 * 2^(20+U)
// (2^U angle units per revolution) leaves x4 in 2^-(32-U) units in the low mantissa word;
// read as a signed integer that word *is* x4 modulo one revolution, good to 1e-9 rad for
// |x4| < 2^(19+U) units (2^19 revolutions).  Replaces F2F.F32.F64 + three FP32
// instructions by DADD + I2F.
#ifndef TJB_PHASE_FIXED
#define TJB_PHASE_FIXED 1
#endif

// Second-order (Halley) FP64 step on the main path instead of the third-order one (default
// on since round 2).  The FP32 stage leaves an error of a few 1e-7 / (1 - e cosE);
// a cubically convergent step takes that below 1e-16 as long as the step itself is small:
// below 2^-17 (7.6e-6) the error is <= 4.4e-16 (t^2/2 - e cosE/(6 f1)), i.e. 1e-15 at
// e = 0.9 and 1e-13 at e = 0.999 in the worst case, and sin(delta) = delta to 7e-17.  The
// shipped threshold is 2^-16 (8x those bounds; +1.1 % throughput on B200): on 20 000 rows
// with e uniform in [0.85, 0.9995] and the MUFU error emulated, ll stays within 4.4e-13 of
// the quad truth for thresholds 2^-17 .. 2^-14 and reaches 4e-11 only at 2^-12.  Saves
// 5 FP64 instructions per epoch; 2 % instead of 0.4 % of the (warp, two-epoch) groups take
// the extra-pass path on the benchmark data (host build of the device code), which keeps
// the third-order step.  Meant to be combined with TJB_PHASE_FIXED (a float-rounded phase
// alone exceeds the tighter threshold beyond ~100 revolutions).
#ifndef TJB_HALLEY
#define TJB_HALLEY 1
#endif

// Convergence vote on the high word of delta^2 (which the rotation needs anyway and which
// has no sign to strip) instead of the sign-stripped high word of delta: one integer
// instruction less per epoch.  Needs TJB_TRIM.
#ifndef TJB_VOTE_D2
#define TJB_VOTE_D2 0
#endif

// 1.5 * 2^52 (1.5 * 2^23): adding it rounds to the nearest integer and leaves
// that integer in the low mantissa bits.
constexpr double kMagic = 6755399441055744.0;
constexpr float kMagicF = 12582912.0f;
constexpr double kTwoPi = 6.28318530717958647692528676655900577;
#ifndef TJB_TRIG_TABLE_LOG2
#define TJB_TRIG_TABLE_LOG2 11
#endif
#if TJB_TRIG_TABLE
constexpr int kTrigTableSize = 1 << TJB_TRIG_TABLE_LOG2;
constexpr double kUnitsPerRev = (double)kTrigTableSize;
#else
constexpr int kTrigTableSize = 0;
constexpr double kUnitsPerRev = 4.0;
#endif
// Two-level table (TJB_TRIG2): the FP64 stage starts from the FP32 estimate E0 snapped to a
// grid of 2^-kFineLog2 angle units (2^22 points per revolution at 2048 + 2048 nodes), so
// that sin / cos of the start are the product of two table nodes -- coarse node j1 (angle
// j1 units) and fine node j2 (angle j2 2^-kFineLog2 units), 4 FP64 instructions and two
// LDS.128 -- instead of one node rotated by a residual through polynomials (9 FP64
// instructions + one LDS.128).  Snapping moves the start by at most half a grid step
// (7.5e-7 rad), well inside the range of the one-pass FP64 step (2^-16); the start's offset
// from M is then formed in FP64 (two instructions), so in total 5 FP64 instructions fewer
// per epoch.  The fine nodes follow the coarse ones in the table (kTrigNodes in all).
#ifndef TJB_TRIG2
#define TJB_TRIG2 1
#endif
// z-stage reciprocal refined from the main step's (see z_from_step_rcp; needs TJB_HALLEY).
// Off by default: it won 0.4-0.8 % at 256- and 640-thread CTAs with 4 epochs per iteration,
// and loses 1.5-2.5 % in the shipped shapes (profiles/r02c_tune_cta_shapes.jsonl: e2_1024 vs
// e2_1024_xz0, jit_xz0) -- the two registers per chain it keeps alive cost more than the MUFU.
#ifndef TJB_XZ
#define TJB_XZ 0
#endif
#if !TJB_HALLEY  // the third-order main step does not hand out its reciprocal
#undef TJB_XZ
#define TJB_XZ 0
#endif
#if !TJB_TRIG_TABLE || !TJB_TRIM  // needs the table back-end and its shared-window addressing
#undef TJB_TRIG2
#define TJB_TRIG2 0
#endif
// Fine nodes per coarse node (2^kFineLog2; the grid has 2^(TJB_TRIG_TABLE_LOG2 + kFineLog2)
// points per revolution: 2^22 at 2048 x 2048, spacing 1.5e-6 rad).
#ifndef TJB_FINE_LOG2
#define TJB_FINE_LOG2 11
#endif
constexpr int kFineLog2 = TJB_FINE_LOG2;
constexpr int kFineNodes = TJB_TRIG2 ? (1 << kFineLog2) : 0;
constexpr int kTrigNodes = kTrigTableSize + kFineNodes;  // nodes of the table in global memory
constexpr double kMagic2 = 6755399441055744.0 / (double)(1 << kFineLog2);  // ulp = 2^-kFineLog2
// (Interleaved per-lane copies of the tables in shared memory -- which make the random
// LDS.128 lookups bank-conflict free, 4 wavefronts instead of 10.4 -- were built and timed
// in round 2c with 2 x 8, 4 x 4, 1 x 8 and 1 x 4 copies: no gain,
// profiles/r02c_tune_cta_shapes.jsonl; the wavefronts are not what holds the kernel back.)
constexpr int kTrigSlots = kTrigNodes;  // 16-byte slots of the staged tables
// fixed-point phase (TJB_PHASE_FIXED): fraction bits that fit one revolution in 32 bits
constexpr double kFixOne = 4294967296.0 / kUnitsPerRev;          // 2^(32-U)
constexpr double kMagicFix = 6755399441055744.0 / kFixOne;       // 1.5 * 2^(52-(32-U))
constexpr double kRadPerUnit = kTwoPi / kUnitsPerRev;
constexpr double kUnitsPerRad = kUnitsPerRev / kTwoPi;

struct alignas(16) SinCos {
  double s, c;
};

#if defined(__CUDA_ARCH__)
#define TJB_COEF __constant__
#else
#define TJB_COEF static const
#endif
#if TJB_TRIG_TABLE
// sin(r h) = r (S0 + S1 r^2 + S2 r^4), cos(r h) = 1 + C1 r^2 + C2 r^4 for |r| <= 1/2,
// h = 2 pi / table size (Taylor; truncation 5e-22 and 1.2e-18 at 1024 nodes).  From 4096
// nodes on the r^5 term of the sine is below 3e-18 and is dropped.
// (TJB_TRIM drops it from 2048 nodes on: 7e-17 at most, below one ulp of the result.)
constexpr bool kSinQuintic = kTrigTableSize < (TJB_TRIM ? 2048 : 4096);
constexpr int kNSin = 3, kNCos = 3;
TJB_COEF double kSinC[3] = {kRadPerUnit, -(kRadPerUnit * kRadPerUnit * kRadPerUnit) / 6.0,
                            (kRadPerUnit * kRadPerUnit * kRadPerUnit * kRadPerUnit * kRadPerUnit) /
                                120.0};
TJB_COEF double kCosC[3] = {1.0, -(kRadPerUnit * kRadPerUnit) / 2.0,
                            (kRadPerUnit * kRadPerUnit * kRadPerUnit * kRadPerUnit) / 24.0};
#else
// sin(w pi/2) = w * sum_k S[k] w^2k,  cos(w pi/2) = sum_k C[k] w^2k on |w| <= 1/2:
// the fdlibm __kernel_sin/__kernel_cos minimax coefficients (|x| <= pi/4) with the
// powers of pi/2 folded in (tests/test_host_logic.py::test_sincos_units checks it).
constexpr int kNSin = 7, kNCos = 8;
TJB_COEF double kSinC[7] = {1.570796326794896619231322,     -0.6459640975062449269023451,
                            0.07969262624606334392301012,   -0.004681754132625933413797963,
                            0.0001604411526673809583408432, -0.000003598649570240691901625632,
                            5.634704113884750951753422e-8};
TJB_COEF double kCosC[8] = {1.0,
                            -1.233700550136169827354311,
                            0.2536695079010476193552061,
                            -0.02086348076333075982186824,
                            0.000919260274390553378473447,
                            -0.00002520203791691774237050555,
                            0.0000004710641505803501879438872,
                            -6.324746678866069891109223e-9};
#endif
// 1/6, angle units per radian, radians per angle unit, 1/2 (TJB_TRIM only)
constexpr int kNMisc = TJB_TRIM ? 4 : 3;
TJB_COEF double kMisc[4] = {1.0 / 6.0, kUnitsPerRad, kRadPerUnit, 0.5};

// ---- bit helpers / pipe-specific primitives -------------------------------
#if defined(__CUDA_ARCH__)
TJB_D int lo32(double x) { return __double2loint(x); }
TJB_D int hi32(double x) { return __double2hiint(x); }
TJB_D double mk64(int hi, int lo) { return __hiloint2double(hi, lo); }
TJB_D float fsin_approx(float x) { return __sinf(x); }
TJB_D float fcos_approx(float x) { return __cosf(x); }
TJB_D float frsqrt_approx(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
TJB_D float frcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// reciprocal of a non-zero, normal double: MUFU.RCP64H seed (20 bits, measured:
// tools/microbench3.cu) + one third-order step r (1 + t + t^2), t = 1 - x r: relative
// error <= 2.2e-16 with 3 FMAs.  The argument in the epoch loop is always 1 - e cosE in
// (1-e, 1+e], so none of the denormal / overflow handling of a general division is needed.
TJB_D double rcp_pos(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  const double t = fma(-x, r, 1.0);
  return fma(r, fma(t, t, t), r);
}
// the same with one Newton step: relative error <= 1e-12, for factors that multiply a
// step of at most 2^-17
TJB_D double rcp_pos_newton(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  return fma(r, fma(-x, r, 1.0), r);
}
#else
inline double rcp_pos_newton(double x) { return 1.0 / x; }
inline int lo32(double x) { int64_t b; memcpy(&b, &x, 8); return (int)(uint32_t)(b & 0xffffffff); }
inline int hi32(double x) { int64_t b; memcpy(&b, &x, 8); return (int)(uint32_t)((uint64_t)b >> 32); }
inline double mk64(int hi, int lo) {
  uint64_t b = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double x; memcpy(&x, &b, 8); return x;
}
#if defined(TJB_EMU_MUFU_NOISE)
// host emulation only: the documented absolute error of sin.approx / cos.approx on
// [-pi, pi] (2^-21.41) as a deterministic pseudo-random perturbation, to estimate how often
// the device's FP32 stage misses the extra-pass threshold
inline float emu_mufu_noise(float x) {
  uint32_t b; memcpy(&b, &x, 4);
  b = b * 2654435761u; b ^= b >> 15; b *= 2246822519u; b ^= b >> 13;
  return ((float)(b >> 8) * (1.0f / 8388608.0f) - 1.0f) * 3.6e-7f;
}
inline float fsin_approx(float x) { return sinf(x) + emu_mufu_noise(x); }
inline float fcos_approx(float x) { return cosf(x) + emu_mufu_noise(-x); }
#else
inline float fsin_approx(float x) { return sinf(x); }
inline float fcos_approx(float x) { return cosf(x); }
#endif
inline float frsqrt_approx(float x) { return 1.0f / sqrtf(x); }
inline float frcp_approx(float x) { return 1.0f / x; }
inline double rcp_pos(double x) { return 1.0 / x; }
#endif

#define TJB_SC(i) tc.s[i]
#define TJB_CC(i) tc.c[i]
#define TJB_MC(i) tc.m[i]
// Constants that sit in the multiplier slot of an FP64 instruction on the main path.  With
// TJB_UCONST they are read from the constant bank where they are used: ptxas hoists them
// into uniform registers (LDCU.64 before the loop, `DFMA R, R, UR, R`), which saves a
// vector-register operand read per use (an FP64 instruction with three distinct vector
// operands issues every 3 cycles, with two every 2: DESIGN.md section 4.1).
#ifndef TJB_UCONST
#define TJB_UCONST 1
#endif
#if TJB_UCONST && defined(__CUDA_ARCH__)
#define TJB_SCM(i) kSinC[i]
#define TJB_CCM(i) kCosC[i]
#define TJB_MCM(i) kMisc[i]
#else
#define TJB_SCM(i) tc.s[i]
#define TJB_CCM(i) tc.c[i]
#define TJB_MCM(i) tc.m[i]
#endif

// The FP64 constants of the epoch loop, pinned in registers.  Left to itself ptxas
// re-loads (LDC) or re-materialises (MOV) each of them on every epoch, which costs issue
// slots; measured: pinned 2.75e9 samples/s, LDC per use 2.61e9, literals 2.57e9.
struct TrigCoef {
  double s[kNSin], c[kNCos], m[kNMisc];
  // kTrigNodes nodes in global (host emulation: host) memory: coarse sin/cos(2 pi j / size),
  // then the fine nodes of TJB_TRIG2; null without a table
  const SinCos *table;
#if TJB_TRIM && defined(__CUDA_ARCH__)
  // 32-bit shared-window address of the tables where they are staged in shared memory (the
  // likelihood kernel; sincos_units<true>): ptxas otherwise re-derives the window base of
  // a generic pointer inside the epoch loop (S2UR / UIADD3 / ULEA / moves).  The fine nodes
  // follow the coarse ones, at a constant byte offset that goes into the load instruction.
  unsigned table_s;
  TJB_D void use_shared_table(const SinCos *staged) {
    table_s = (unsigned)__cvta_generic_to_shared(staged);
    asm volatile("" : "+r"(table_s));  // opaque: keep it in a register
  }
#endif
  // `zero` must be a run-time 0.0 (a kernel parameter): coefficient + zero is an
  // FP64 result ptxas will not rematerialise, so the values stay in registers.
  TJB_HD void load(double zero, const SinCos *tab) {
    table = tab;
#pragma unroll
    for (int i = 0; i < kNSin; i++) s[i] = kSinC[i] + zero;
#pragma unroll
    for (int i = 0; i < kNCos; i++) c[i] = kCosC[i] + zero;
#pragma unroll
    for (int i = 0; i < kNMisc; i++) m[i] = kMisc[i] + zero;
  }
};

// sin and cos of an angle v given in angle units (1/kUnitsPerRev revolution), any size
// up to 2^51: exact reduction by magic-number rounding.
template <bool kSharedTable = false>
TJB_HD void sincos_units(const TrigCoef &tc, double v, double &s, double &c) {
  const double tv = v + kMagic;
  const double r = v - (tv - kMagic);  // in [-0.5, 0.5]
  const int k = lo32(tv);              // nearest node / quadrant (mod 2^32)
  const double r2 = r * r;
#if TJB_TRIG_TABLE
#if TJB_TRIM && defined(__CUDA_ARCH__)
  SinCos node;
  if (kSharedTable) {
    unsigned addr;  // mask, then one multiply-add (the compiler's shift / mask / add is three)
    asm("mad.lo.u32 %0, %1, 16, %2;" : "=r"(addr) : "r"(k & (kTrigTableSize - 1)), "r"(tc.table_s));
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(node.s), "=d"(node.c) : "r"(addr));
  } else {
    node = tc.table[k & (kTrigTableSize - 1)];
  }
#else
  const SinCos node = tc.table[k & (kTrigTableSize - 1)];
#endif
  const double sr = kSinQuintic ? r * fma(r2, fma(r2, TJB_SC(2), TJB_SC(1)), TJB_SC(0))
                                : r * fma(r2, TJB_SCM(1), TJB_SC(0));
  const double cr = fma(r2, fma(r2, TJB_CCM(2), TJB_CC(1)), 1.0);
  s = fma(node.c, sr, node.s * cr);
  c = fma(-node.s, sr, node.c * cr);
#else
  double ps = fma(TJB_SC(6), r2, TJB_SC(5));
  double pc = fma(TJB_CC(7), r2, TJB_CC(6));
  ps = fma(ps, r2, TJB_SC(4));
  pc = fma(pc, r2, TJB_CC(5));
  ps = fma(ps, r2, TJB_SC(3));
  pc = fma(pc, r2, TJB_CC(4));
  ps = fma(ps, r2, TJB_SC(2));
  pc = fma(pc, r2, TJB_CC(3));
  ps = fma(ps, r2, TJB_SC(1));
  pc = fma(pc, r2, TJB_CC(2));
  ps = fma(ps, r2, TJB_SC(0));
  pc = fma(pc, r2, TJB_CC(1));
  const double sx = ps * r;
  const double cx = fma(pc, r2, 1.0);
  // quadrant rotation: k=0 (s,c) k=1 (c,-s) k=2 (-s,-c) k=3 (-c,s)
  const bool swap = k & 1;
  const double s0 = swap ? cx : sx;
  const double c0 = swap ? sx : cx;
  // sign flips on the high words (integer pipe): bit1 of k for sin, bit1 of k+1 for cos
  s = mk64(hi32(s0) ^ ((k << 30) & 0x80000000), lo32(s0));
  c = mk64(hi32(c0) ^ (((k + 1) << 30) & 0x80000000), lo32(c0));
#endif
}

// per-sample constants of the Kepler / RV evaluation
struct OrbitConsts {
  double nu;    // kUnitsPerRev / P            [angle units per day]
  double ph;    // M0 * kUnitsPerRad           [angle units]
  double e;     // eccentricity
  double e6;    // e / 6
  double a;     // cos(omega)
  double b;     // -sqrt(1-e^2) sin(omega)
  double ea;    // e cos(omega)
  float ef;     // (float) e
  float g0f;    // (float)(1 + e^2)
};

TJB_HD OrbitConsts make_orbit_consts(const TrigCoef &tc, double P, double e, double omega,
                                     double M0) {
  OrbitConsts oc;
  oc.nu = kUnitsPerRev / P;
  oc.ph = M0 * kUnitsPerRad;
  oc.e = e;
  oc.e6 = e * (1.0 / 6.0);
  // sin / cos of omega through the same reduction as the epochs (|omega| is a few
  // radians; the product with kUnitsPerRad costs < 1e-15 rad)
  double so, co;
  sincos_units(tc, omega * kUnitsPerRad, so, co);
  oc.a = co;
  oc.b = -sqrt(fma(-e, e, 1.0)) * so;
  oc.ea = e * co;
  oc.ef = fminf((float)e, 0.99999994f);  // keep the FP32 starter finite as e -> 1
  oc.g0f = (float)fma(e, e, 1.0);
  return oc;
}

// warp-uniform "any lane" vote; on the host emulation a lane is its own warp
#if defined(__CUDA_ARCH__)
TJB_D bool any_lane(bool p) { return __any_sync(0xffffffffu, p); }
#else
inline bool any_lane(bool p) { return p; }
#endif

struct SolveStats {
  int extra_f32;  // unused (the FP32 stage is fixed-work); kept for the ABI
  int extra_f64;  // FP64 Householder passes beyond the first
  int not_converged;
};

// A third-order Householder step maps an error eps to ~C eps^4 with C = O(1) for
// e <~ 0.95 (tools/kepler_solver_study.py): a lane repeats the pass while it moved by
// 2^-13 (1.2e-4) or more, at most kF64MaxIter times (bisection steps of the safeguard
// included: e = 1 - 1e-5 at M = 1e-5 takes ~25).
constexpr int kF64MaxIter = 64;
// high word of the threshold 2^-TJB_NEED_LOG2 (exponent field only)
#ifndef TJB_NEED_LOG2
#define TJB_NEED_LOG2 (TJB_HALLEY ? 16 : 13)
#endif
constexpr unsigned kNeedHi = (unsigned)(1023 - TJB_NEED_LOG2) << 20;
constexpr unsigned kNeedHiSq = (unsigned)(1023 - 2 * TJB_NEED_LOG2) << 20;  // of the threshold squared

// run-wide solver statistics (device counters, touched only on the rare path):
// [0] extra FP64 passes (lane-epochs), [1] epochs that hit kF64MaxIter
TJB_HD void count_event(unsigned long long *gstats, int which) {
#if defined(__CUDA_ARCH__)
  if (gstats) atomicAdd(gstats + which, 1ULL);
#else
  if (gstats) gstats[which]++;
#endif
}

// One third-order Householder step from (D, sE, cE): returns delta.
//   f = D - e sinE, f1 = 1 - e cosE, f2 = e sinE, f3 = e cosE
//   u = -f/f1, t = f2/f1, b6 = f3/(6 f1), delta = u (1 + u (-t/2 + u (t^2/2 - b6)))
TJB_HD double householder3(const OrbitConsts &oc, const TrigCoef &tc, double D, double sE,
                           double cE) {
  const double es = oc.e * sE;
  const double r = rcp_pos(fma(-oc.e, cE, 1.0));
  const double t = es * r;
  const double u = fma(-D, r, t);
#if TJB_TRIM
  // e cosE / f1 = (1 - f1)/f1 = r - 1; the 2e-16 r absolute error enters delta times u^3
  const double b6 = fma(r, TJB_MC(0), -TJB_MC(0));
#else
  (void)tc;
  const double b6 = (oc.e6 * cE) * r;
#endif
  const double th = 0.5 * t;
  const double q = fma(th, t, -b6);
  return u * fma(u, fma(u, q, -th), 1.0);
}

// rotate (sinE, cosE) by delta, |delta| <= 1e-4: sin d = d (1 - d^2/6) + O(1e-22),
// cos d = 1 - d^2/2 + O(4e-18)
TJB_HD void rotate_small(const TrigCoef &tc, double del, double &sE, double &cE) {
  const double d2 = del * del;
  const double sd = del * fma(d2, -TJB_MC(0), 1.0);
#if TJB_TRIM
  const double cd = fma(d2, -TJB_MC(3), 1.0);
#else
  const double cd = fma(d2, -0.5, 1.0);
#endif
  const double sN = fma(cE, sd, sE * cd);
  cE = fma(-sE, sd, cE * cd);
  sE = sN;
}

// Main-path step and rotation of the TJB_HALLEY variant: delta = u (1 - u t / 2) with
// u = -f/f1, t = f2/f1; rotation by |delta| < 2^-17 with sin d = d, cos d = 1 - d^2/2.
TJB_HD double halley2(const OrbitConsts &oc, double D, double sE, double cE, double &r) {
  const double es = oc.e * sE;
  // 1/f1 to 1e-12 is enough: its error enters delta times |u| <= 2^-17
  r = rcp_pos_newton(fma(-oc.e, cE, 1.0));
  const double t = es * r;
  const double u = fma(-D, r, t);
  return u * fma(u * -0.5, t, 1.0);
}
TJB_HD void rotate_tiny(const TrigCoef &tc, double del, double &sE, double &cE) {
  (void)tc;
#if TJB_TRIM
  const double cd = fma(del * del, -TJB_MCM(3), 1.0);
#else
  const double cd = fma(del * del, -0.5, 1.0);
#endif
  const double sN = fma(cE, del, sE * cd);
  cE = fma(-sE, del, cE * cd);
  sE = sN;
}
#if TJB_HALLEY
#define TJB_MAIN_STEP(oc, tc, D, s, c, r) halley2(oc, D, s, c, r)
#define TJB_MAIN_ROTATE rotate_tiny
#else
#define TJB_MAIN_STEP(oc, tc, D, s, c, r) householder3(oc, tc, D, s, c)
#define TJB_MAIN_ROTATE rotate_small
#endif

// The rare path of the solver: lanes whose first FP64 step moved by 2^-13 or more
// (e >~ 0.8 near pericentre, a phase beyond the FP32 stage's range) re-evaluate sincos in
// full at the updated E and repeat the step until it is small; lanes that had converged
// (`nd` false) return `frozen`, exactly the values of the normal path.  All lanes of the
// warp call together (the loop votes).  A function of its own, not inlined: its registers
// and code stay out of the epoch loop's allocation and instruction stream, and it takes
// scalars only so that the caller's constants stay in registers.
//
// Safeguard: the root obeys |E - M| <= e and g(D) = D - e sin(M + D) increases with D, so
// [-e, e] brackets it.  Every pass narrows the bracket by the sign of g; a start or a
// Householder step that leaves it (or is NaN: wild FP32 estimate, e -> 1 at M -> 0) is
// replaced by the clamp / the midpoint, so the iteration cannot diverge and the result is
// finite for every e < 1 (tests/test_host_logic.py::test_kepler_solver_extreme_cases).
template <bool kCountStats>
TJB_HD_NOINLINE SinCos solve_extra_passes(double e, const SinCos *table, double x4, double Dstart,
                                          SinCos frozen, bool nd, SolveStats *st,
                                          unsigned long long *gstats) {
  TrigCoef tc;
  tc.load(0.0, table);
  OrbitConsts oc;
  oc.e = e;
  oc.e6 = e * (1.0 / 6.0);
  double Dlo = -e, Dhi = e;
  double Dk = fmin(fmax(Dstart, Dlo), Dhi);
  SinCos out = frozen;
#pragma unroll 1
  for (int it = 1; it < kF64MaxIter && any_lane(nd); ++it) {
    double s2, c2;
    sincos_units(tc, fma(Dk, TJB_MC(1), x4), s2, c2);
    const double d2 = householder3(oc, tc, Dk, s2, c2);
    if (nd) {
      if (kCountStats) st->extra_f64++;
      count_event(gstats, 0);
      if ((unsigned)(hi32(d2) & 0x7fffffff) < kNeedHi) {
        rotate_small(tc, d2, s2, c2);
        out.s = s2;
        out.c = c2;
        nd = false;
      } else {
        // |g| >> rounding here, so its sign is reliable
        if (fma(-e, s2, Dk) > 0.0) Dhi = Dk; else Dlo = Dk;
        const double Dn = Dk + d2;
        Dk = (Dn > Dlo && Dn < Dhi) ? Dn : 0.5 * (Dlo + Dhi);  // a NaN step bisects
      }
    }
  }
  if (nd) {  // did not converge within kF64MaxIter passes: best estimate, counted
    if (kCountStats) st->not_converged++;
    count_event(gstats, 1);
    sincos_units(tc, fma(Dk, TJB_MC(1), x4), out.s, out.c);
  }
  return out;
}

// z from (sinE, cosE) after the main step, with 1 / (1 - e cosE) refined from the step's own
// reciprocal r0 = 1 / (1 - e cosE0) (good to 1e-12) instead of a fresh MUFU.RCP64H seed
// (TJB_XZ): eps = 1 - f1 r0 = (e sinE0 / f1) delta + O(1e-12) with |delta| < 2^-16 on the
// main path, and r = r0 (1 + eps + eps^2 + eps^3) is good to eps^4 (< 1e-14 for e <= 0.9987
// in the worst case).  One FP64 instruction more, one MUFU (8 cycles of the XU pipe per
// warp instruction, profiles/r02c_microbench4.txt) and one move less per epoch.
TJB_HD double z_from_step_rcp(const OrbitConsts &oc, double sE, double cE, double r0) {
  const double eps = fma(-fma(-oc.e, cE, 1.0), r0, 1.0);
  const double p = fma(fma(eps, eps, eps), eps, eps);
  const double r = fma(r0, p, r0);
  const double num = fma(oc.b, sE, fma(oc.a, cE, -oc.ea));
  return fma(num, r, oc.ea);
}

// K epochs of one sample at once (K independent dependency chains interleaved by
// the compiler).  dt[k] = t_n - t_ref [day]; z[k] receives z_n.  All lanes of a warp
// must call together.  The result of a lane depends only on that lane's inputs (not
// on its warp-mates): lanes that need extra passes iterate under a per-lane flag.
template <int K, bool kCountStats, bool kSharedTable = false, bool kXZ = (TJB_XZ != 0)>
TJB_HD void rv_unit_columns(const OrbitConsts &oc, const TrigCoef &tc, const double *dt, double *z,
                            SolveStats *st, unsigned long long *gstats = nullptr) {
  double x4[K], D[K], sE[K], cE[K];  // x4: mean anomaly in angle units (unreduced)
  float Df[K];

  // ---- FP32: reduce, starter D0 = e sinM / sqrt(1 - 2 e cosM + e^2), one
  //      third-order Householder step.  Errors of this stage (including the
  //      ~1e-5 from rounding x4 to float) only move the FP64 starting point.
  const float ef = oc.ef;
#pragma unroll
  for (int k = 0; k < K; k++) {
    x4[k] = fma(dt[k], oc.nu, -oc.ph);
#if TJB_PHASE_FIXED
    const float Mf = (float)lo32(x4[k] + kMagicFix) * (float)(kRadPerUnit / kFixOne);  // [-pi, pi)
#elif TJB_TRIM
    const float x4f = (float)x4[k];
    // adding 1.5 * 2^23 * (units per revolution) rounds to whole revolutions, in angle units
    const float rev = (x4f + kMagicF * (float)kUnitsPerRev) - kMagicF * (float)kUnitsPerRev;
    const float Mf = (x4f - rev) * (float)kRadPerUnit;  // M in [-pi, pi]
#else
    const float x4f = (float)x4[k];
    const float r4 = (x4f * (float)(1.0 / kUnitsPerRev) + kMagicF) - kMagicF;  // whole revolutions
    const float Mf = fmaf(r4, -(float)kUnitsPerRev, x4f) * (float)kRadPerUnit;  // M in [-pi, pi]
#endif
    const float sM = fsin_approx(Mf), cM = fcos_approx(Mf);
    const float D0 = ef * sM * frsqrt_approx(fmaf(-2.0f * ef, cM, oc.g0f));
    const float Ef = Mf + D0;
    const float es = ef * fsin_approx(Ef), ec = ef * fcos_approx(Ef);
    const float r = frcp_approx(1.0f - ec);
    const float t = es * r;           // f2/f1
    const float u = fmaf(-D0, r, t);  // -f/f1
#if TJB_TRIM
    const float th = 0.5f * t;
    const float q = fmaf(th, t, fmaf(r, -1.0f / 6.0f, 1.0f / 6.0f));  // -ec r/6 = (1 - r)/6
    const float del = u * fmaf(u, fmaf(u, q, -th), 1.0f);
    Df[k] = D0 + del;  // not clamped here: a wild or NaN estimate is caught by the FP64 test
#else
    const float q = fmaf(0.5f * t, t, ec * r * (-1.0f / 6.0f));
    const float del = u * fmaf(u, fmaf(u, q, -0.5f * t), 1.0f);
    Df[k] = fminf(fmaxf(D0 + del, -ef), ef);  // |E - M| <= e holds for the root
#endif
  }

  // ---- FP64: exact reduction of E0 = M + D0, one full sincos, one Householder step
  double del[K];
  double r0[K];  // 1 / (1 - e cosE0) to 1e-12 from the main step (TJB_XZ)
  bool need[K];
  bool any_need = false;
#if TJB_TRIM
  unsigned top = 0;  // max over the epochs of the high word without its sign bit
#endif
#pragma unroll
  for (int k = 0; k < K; k++) {
#if TJB_TRIG2
    {
      const double v0 = fma((double)Df[k], TJB_MCM(1), x4[k]);  // E0 in angle units
      const double tv = v0 + kMagic2;                            // ... on the fine grid
      const int idx = lo32(tv);  // grid index modulo 2^32: fine node | coarse node << kFineLog2
      D[k] = ((tv - kMagic2) - x4[k]) * TJB_MCM(2);  // snapped E0 - M [rad] (difference exact)
      SinCos n1, n2;
#if defined(__CUDA_ARCH__)
      if (kSharedTable) {
        unsigned a1, a2;
        asm("mad.lo.u32 %0, %1, 16, %2;" : "=r"(a1)
            : "r"(((unsigned)idx >> kFineLog2) & (kTrigTableSize - 1)), "r"(tc.table_s));
        asm("mad.lo.u32 %0, %1, 16, %2;" : "=r"(a2) : "r"(idx & ((1 << kFineLog2) - 1)),
            "r"(tc.table_s));
        asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(n1.s), "=d"(n1.c) : "r"(a1));
        asm("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(n2.s), "=d"(n2.c)
            : "r"(a2), "n"(kTrigTableSize * sizeof(SinCos)));
      } else
#endif
      {
        n1 = tc.table[((unsigned)idx >> kFineLog2) & (kTrigTableSize - 1)];
        n2 = tc.table[kTrigTableSize + (idx & ((1 << kFineLog2) - 1))];
      }
      sE[k] = fma(n1.c, n2.s, n1.s * n2.c);
      cE[k] = fma(-n1.s, n2.s, n1.c * n2.c);
    }
#elif TJB_TRIM
    D[k] = (double)Df[k];                // E0 - M [rad]
    const double d4 = D[k] * TJB_MCM(1);  // in angle units (1-ulp rounding: < 2e-16 rad)
    sincos_units<kSharedTable>(tc, x4[k] + d4, sE[k], cE[k]);
#else
    const double d4 = (double)(Df[k] * (float)kUnitsPerRad);  // D0 in angle units
    sincos_units<kSharedTable>(tc, x4[k] + d4, sE[k], cE[k]);
    D[k] = d4 * TJB_MC(2);  // E0 - M [rad]
#endif
    del[k] = TJB_MAIN_STEP(oc, tc, D[k], sE[k], cE[k], r0[k]);
    // error map of the step: eps -> ~C eps^4 (tools/kepler_solver_study.py); a lane
    // that moved by 2^-13 (1.2e-4) or more, or produced a NaN, takes further passes.
    // The test reads the exponent field on the integer pipe instead of a DSETP.
#if TJB_TRIM && TJB_VOTE_D2
    const unsigned hk = (unsigned)hi32(del[k] * del[k]);  // NaN: 0x7ff8.. / 0xfff8.. -> above
    top = hk > top ? hk : top;
#elif TJB_TRIM
    const unsigned hk = (unsigned)hi32(del[k]) << 1;
    top = hk > top ? hk : top;
#else
    need[k] = (unsigned)(hi32(del[k]) & 0x7fffffff) >= kNeedHi;
    any_need = any_need || need[k];
#endif
  }
#if TJB_TRIM && TJB_VOTE_D2
  any_need = top >= kNeedHiSq;
#elif TJB_TRIM
  any_need = top >= (kNeedHi << 1);
#endif
  if (!any_lane(any_need)) {
    // the normal case: every lane of the warp converged in one pass
#pragma unroll
    for (int k = 0; k < K; k++) TJB_MAIN_ROTATE(tc, del[k], sE[k], cE[k]);
    if (kXZ) {
#pragma unroll
      for (int k = 0; k < K; k++) z[k] = z_from_step_rcp(oc, sE[k], cE[k], r0[k]);
      return;
    }
  } else {
    // rare: see solve_extra_passes
    bool redo[K];
#pragma unroll
    for (int k = 0; k < K; k++) {
      SinCos fr = {sE[k], cE[k]};
      TJB_MAIN_ROTATE(tc, del[k], fr.s, fr.c);
#if TJB_TRIM && TJB_VOTE_D2
      need[k] = (unsigned)hi32(del[k] * del[k]) >= kNeedHiSq;
#elif TJB_TRIM
      need[k] = ((unsigned)hi32(del[k]) << 1) >= (kNeedHi << 1);
#endif
      // (kXZ) a converged lane keeps exactly the z of the main path: its result must not
      // depend on what its warp-mates needed
      redo[k] = need[k];
      if (kXZ) z[k] = z_from_step_rcp(oc, fr.s, fr.c, r0[k]);
      const SinCos r = solve_extra_passes<kCountStats>(oc.e, tc.table, x4[k], D[k] + del[k], fr,
                                                       need[k], st, gstats);
      sE[k] = r.s;
      cE[k] = r.c;
    }
    if (kXZ) {
#pragma unroll
      for (int k = 0; k < K; k++) {
        const double r = rcp_pos(fma(-oc.e, cE[k], 1.0));
        const double num = fma(oc.b, sE[k], fma(oc.a, cE[k], -oc.ea));
        if (redo[k]) z[k] = fma(num, r, oc.ea);
      }
      return;
    }
  }

  // ---- z = [a (cosE - e) + b sinE] / (1 - e cosE) + e a ---------------------
#pragma unroll
  for (int k = 0; k < K; k++) {
    const double r = rcp_pos(fma(-oc.e, cE[k], 1.0));
    const double num = fma(oc.b, sE[k], fma(oc.a, cE[k], -oc.ea));
    z[k] = fma(num, r, oc.ea);
  }
}

// One epoch.
template <bool kCountStats, bool kSharedTable = false, bool kXZ = (TJB_XZ != 0)>
TJB_HD double rv_unit_column(const OrbitConsts &oc, const TrigCoef &tc, double dt, SolveStats *st) {
  double z;
  rv_unit_columns<1, kCountStats, kSharedTable, kXZ>(oc, tc, &dt, &z, st);
  return z;
}

}  // namespace tjb
