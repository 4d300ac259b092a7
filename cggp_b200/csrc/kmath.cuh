// Stationary-kernel arithmetic (GPflow K_r2 / K_r; see include/cggp_b200.h for the reference call sites).
//   SE        : v * exp(-0.5 r2)                          (no clamp: GPflow SquaredExponential.K_r2)
//   Matern-1/2: r = sqrt(max(r2, 1e-36)); v * exp(-r)
//   Matern-3/2: v * (1 + sqrt3 r) * exp(-sqrt3 r)
//   Matern-5/2: v * (1 + sqrt5 r + 5/3 r^2) * exp(-sqrt5 r)
#pragma once
#include <cuda_runtime.h>

#include "../../include/cggp_b200.h"

template <typename T>
struct KConst;
template <>
struct KConst<double> {
  static __device__ __forceinline__ double sqrt3() { return 1.7320508075688772; }
  static __device__ __forceinline__ double sqrt5() { return 2.23606797749979; }
  static __device__ __forceinline__ double c53() { return 5.0 / 3.0; }
  static __device__ __forceinline__ double clamp() { return 1e-36; }
};
template <>
struct KConst<float> {
  // float32 casts of the float64 constants, as GPflow builds them in the default float
  static __device__ __forceinline__ float sqrt3() { return 1.7320508075688772f; }
  static __device__ __forceinline__ float sqrt5() { return 2.23606797749979f; }
  static __device__ __forceinline__ float c53() { return (float)(5.0 / 3.0); }
  static __device__ __forceinline__ float clamp() { return 1e-36f; }
};

__device__ __forceinline__ double xexp(double x) { return exp(x); }
__device__ __forceinline__ float xexp(float x) { return expf(x); }
__device__ __forceinline__ double xsqrt(double x) { return sqrt(x); }
__device__ __forceinline__ float xsqrt(float x) { return sqrtf(x); }
__device__ __forceinline__ double xmax(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ float xmax(float a, float b) { return fmaxf(a, b); }

// Library-math kernel value (1-ulp exp / correctly rounded sqrt): the accuracy baseline of the library.
template <typename T, int KIND>
__device__ __forceinline__ T kernel_value(T r2, T variance) {
  if (KIND == CGGP_SE) {
    return variance * xexp(T(-0.5) * r2);
  }
  T r = xsqrt(xmax(r2, KConst<T>::clamp()));
  if (KIND == CGGP_MATERN12) {
    return variance * xexp(-r);
  } else if (KIND == CGGP_MATERN32) {
    T s = KConst<T>::sqrt3() * r;
    return variance * (T(1) + s) * xexp(-s);
  } else {
    T s = KConst<T>::sqrt5() * r;
    return variance * (T(1) + s + KConst<T>::c53() * (r * r)) * xexp(-s);
  }
}

// ---------------------------------------------------------------------------------------------------------
// FP64 fast path for the fused matvec kernel.  On B200 the FP64 pipe (64 DFMA/clk/SM) is the binding unit, so the
// transcendental parts are written to minimise FP64-pipe instructions and push everything else to the ALU / SFU /
// shuffle pipes, which issue in the gaps.
// ---------------------------------------------------------------------------------------------------------

// sqrt(u) for normal u > 0: MUFU.RSQ64H seed (2^-22.9) + two residual-corrected Newton steps: 1 DMUL + 4 DFMA.
__device__ __forceinline__ double fast_sqrt_pos(double u) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(u));
  // h = y / 2 by an exponent decrement on the ALU pipe (y is a normal number far from the subnormal range here)
  double h = __hiloint2double(__double2hiint(y) - 0x00100000, __double2loint(y));
  double g = u * y;
  double e = fma(-g, g, u);
  g = fma(e, h, g);
  e = fma(-g, g, u);
  g = fma(e, h, g);
  return g;
}

// Same, for the issue-bound pipelined matvec: the low word of the seed is left UNDEFINED instead of zeroed (one
// integer move less per use).  Any low word perturbs the 2^-22.9 seed by < 2^-20 relative; the two corrected Newton
// steps square the error twice (2^-20 -> 2^-39 -> 2^-77), so the result is unchanged to the last bit or two.
__device__ __forceinline__ double fast_sqrt_pos_lean(double u) {
  double y, h;
  asm("{\n"
      ".reg .b32 ulo, uhi, yhi, hhi;\n"
      ".reg .f64 yy;\n"
      "rsqrt.approx.ftz.f64 yy, %2;\n"
      "mov.b64 {ulo, yhi}, yy;\n"
      "mov.b64 {ulo, uhi}, %2;\n"
      "add.s32 hhi, yhi, -1048576;\n"
      "mov.b64 %0, {ulo, yhi};\n"
      "mov.b64 %1, {ulo, hhi};\n"
      "}\n"
      : "=d"(y), "=d"(h)
      : "d"(u));
  double g = u * y;
  double e = fma(-g, g, u);
  g = fma(e, h, g);
  e = fma(-g, g, u);
  g = fma(e, h, g);
  return g;
}

// sqrt(u) by ONE third-order step from the MUFU.RSQ64H seed y (relative error 2^-22.9):
//   g = u y,  r = 1 - g y,  sqrt(u) = g (1 - r)^(-1/2) = g (1 + r/2 + 3 r^2/8 + O(r^3)),  |r|^3 < 2^-65.
// 2 DMUL + 3 DFMA like the two Newton steps above, but no y/2 is needed (one integer add and one register move
// less per value on the issue-bound matvec) and the dependency chain is one FP64 instruction shorter.  r is exact
// for the rounded g (FMA), so the result carries the rounding of g (2^-54 relative after the square root) plus the
// final rounding: < 1 ulp.  The low word of the seed is left undefined as in fast_sqrt_pos_lean.
__device__ __forceinline__ double fast_sqrt_pos_cubic(double u) {
  double y;
  asm("{\n"
      ".reg .b32 ulo, uhi, yhi;\n"
      ".reg .f64 yy;\n"
      "rsqrt.approx.ftz.f64 yy, %1;\n"
      "mov.b64 {ulo, yhi}, yy;\n"
      "mov.b64 {ulo, uhi}, %1;\n"
      "mov.b64 %0, {ulo, yhi};\n"
      "}\n"
      : "=d"(y)
      : "d"(u));
  const double g = u * y;
  const double r = fma(-g, y, 1.0);
  const double p = fma(r, 0.375, 0.5);
  const double gr = g * r;
  return fma(gr, p, g);
}

// exp(x) for x <= ~0 (any x in [-745, 700] works).  x*32/ln2 = n + f, n = 32 k + j:
//   exp(x) = 2^k * 2^(j/32) * exp(d),  d = x - n ln2/32,  |d| <= ln2/64
// 2^(j/32) comes from a 32-entry table held one entry per lane (two SHFLs, no shared memory), exp(d) from a
// degree-5 near-minimax polynomial (|rel err| < 1e-16 on the interval).  FP64-pipe cost: 3 (reduction) + 5 (poly)
// + 1 (table multiply) = 9 instructions; the exponent add and table indexing run on the ALU pipe.
struct FastExpTable {
  int hi, lo;  // this lane's entry 2^(lane/32)
};
__device__ __forceinline__ FastExpTable fast_exp_table() {
  double t = exp2((double)(threadIdx.x & 31) * (1.0 / 32.0));
  FastExpTable r;
  r.hi = __double2hiint(t);
  r.lo = __double2loint(t);
  return r;
}
__device__ __forceinline__ double fast_exp(double x, const FastExpTable& tab) {
  const double L2E32 = 46.16624130844682903551758979206054839765;  // 32 / ln 2
  const double MAGIC = 6755399441055744.0;                           // 1.5 * 2^52
  const double LN2_32 = 0.02166084939249829091928849858592451515688;
  // clamp on the ALU pipe: for x < 0 the high word grows with |x| as an unsigned integer
  {
    unsigned hx = (unsigned)__double2hiint(x);
    if (hx > 0xC0874000u) x = -744.0;  // x < -744 (0xC0874000 = high word of -744.0)
  }
  double t = fma(x, L2E32, MAGIC);
  int n = __double2loint(t);
  double nf = t - MAGIC;
  double d = fma(nf, -LN2_32, x);
  // minimax (Remez) coefficients of exp(d) on |d| <= ln2/64, max rel err 1.41e-16 (tools/fit_exp_poly.py)
  double q = fma(d, 8.33337406147829918e-03, 4.16668703096581480e-02);
  q = fma(q, d, 1.66666666664448626e-01);
  q = fma(q, d, 4.99999999994028277e-01);
  q = fma(q, d, 1.0);
  q = fma(q, d, 1.0);
  int j = n & 31;
  int hi = __shfl_sync(0xffffffffu, tab.hi, j);
  int lo = __shfl_sync(0xffffffffu, tab.lo, j);
  hi += (n >> 5) << 20;  // * 2^k; stays normal for x >= -708, below that the result is < 1e-307 and flushed
  // below x = -708 the scaled table value would leave the normal range; the true result is < 1e-307: flush to 0
  // (integer compare on the high word: x < -708.0 <=> hi(x) > hi(-708.0) as unsigned, ALU pipe)
  if ((unsigned)__double2hiint(x) > 0xC0862000u) { hi = 0; lo = 0; }
  return __hiloint2double(hi, lo) * q;
}

// exp(-a) for a >= 0 (the Matern argument): same scheme as fast_exp with the negation folded into the FMA operand
// modifiers, so no FP64-pipe instruction is spent on forming -a.
__device__ __forceinline__ double fast_exp_neg(double a, const FastExpTable& tab) {
  const double L2E32 = 46.16624130844682903551758979206054839765;
  const double MAGIC = 6755399441055744.0;
  const double LN2_32 = 0.02166084939249829091928849858592451515688;
  if ((unsigned)__double2hiint(a) > 0x40874000u) a = 744.0;  // a > 744 (also catches NaN): result flushed below
  double t = fma(a, -L2E32, MAGIC);
  int n = __double2loint(t);
  double nf = t - MAGIC;
  double d = fma(nf, -LN2_32, -a);
  double q = fma(d, 8.33337406147829918e-03, 4.16668703096581480e-02);
  q = fma(q, d, 1.66666666664448626e-01);
  q = fma(q, d, 4.99999999994028277e-01);
  q = fma(q, d, 1.0);
  q = fma(q, d, 1.0);
  int j = n & 31;
  int hi = __shfl_sync(0xffffffffu, tab.hi, j);
  int lo = __shfl_sync(0xffffffffu, tab.lo, j);
  hi += (n >> 5) << 20;
  if ((unsigned)__double2hiint(a) > 0x40862000u) { hi = 0; lo = 0; }  // a > 708: true result < 1e-307, flush to 0
  return __hiloint2double(hi, lo) * q;
}

// Table for the clamp-free cores below: the high word of entry j is pre-decremented by j << 15, so that the scaling
// by 2^k (n = 32 k + j) is ONE integer multiply-add, hi = n * 2^15 + table_hi[j], instead of shift + mask + add.
// (The kernel is bound by instruction issue - every FP64 instruction holds the dispatch port for 2 cycles, every other
// instruction for 1 - so integer instructions are not free, see DESIGN.md.)
__device__ __forceinline__ FastExpTable fast_exp_table_biased() {
  const int lane = threadIdx.x & 31;
  double t = exp2((double)lane * (1.0 / 32.0));
  FastExpTable r;
  r.hi = __double2hiint(t) - (lane << 15);
  r.lo = __double2loint(t);
  return r;
}

// Clamp-free cores used by the pipelined matvec (they take the BIASED table), whose caller bounds the argument with two integer min / max on the
// high word (ALU pipe) instead: x in [-708.4, 709] resp. a in [0, 708.4], results are normal numbers.
__device__ __forceinline__ double fast_exp_core(double x, const FastExpTable& tab) {
  const double L2E32 = 46.16624130844682903551758979206054839765;
  const double MAGIC = 6755399441055744.0;
  const double LN2_32 = 0.02166084939249829091928849858592451515688;
  double t = fma(x, L2E32, MAGIC);
  int n = __double2loint(t);
  double nf = t - MAGIC;
  double d = fma(nf, -LN2_32, x);
  double q = fma(d, 8.33337406147829918e-03, 4.16668703096581480e-02);
  q = fma(q, d, 1.66666666664448626e-01);
  q = fma(q, d, 4.99999999994028277e-01);
  q = fma(q, d, 1.0);
  q = fma(q, d, 1.0);
  int hi = __shfl_sync(0xffffffffu, tab.hi, n);  // SHFL uses the lane index modulo 32
  int lo = __shfl_sync(0xffffffffu, tab.lo, n);
  hi += n << 15;  // = table_hi[j] + (k << 20)
  return __hiloint2double(hi, lo) * q;
}
__device__ __forceinline__ double fast_exp_neg_core(double a, const FastExpTable& tab) {
  const double L2E32 = 46.16624130844682903551758979206054839765;
  const double MAGIC = 6755399441055744.0;
  const double LN2_32 = 0.02166084939249829091928849858592451515688;
  double t = fma(a, -L2E32, MAGIC);
  int n = __double2loint(t);
  double nf = t - MAGIC;
  double d = fma(nf, -LN2_32, -a);
  double q = fma(d, 8.33337406147829918e-03, 4.16668703096581480e-02);
  q = fma(q, d, 1.66666666664448626e-01);
  q = fma(q, d, 4.99999999994028277e-01);
  q = fma(q, d, 1.0);
  q = fma(q, d, 1.0);
  int hi = __shfl_sync(0xffffffffu, tab.hi, n);
  int lo = __shfl_sync(0xffffffffu, tab.lo, n);
  hi += n << 15;
  return __hiloint2double(hi, lo) * q;
}

// ---------------------------------------------------------------------------------------------------------
// Shared-memory table variant of the clamp-free cores: 2^(j / 2^TBITS) from a table of 2^TBITS entries in shared
// memory (one LDS.64 instead of two SHFLs), which shortens the polynomial: TBITS = 10 -> degree 3 (2 DFMA less per
// value than the 32-entry / degree-5 scheme).  Entries are (lo, hi - (j << (20 - TBITS))), so that the 2^k scaling is
// one integer multiply-add on n = 2^TBITS k + j; the table is built on the host (correctly rounded from long
// double, matvec_pipe.cu).  Max relative error of the polynomial in exact arithmetic with the double coefficients:
// 1.37e-16 (the degree-5 / 32-entry scheme: 1.41e-16); tools/fit_exp_poly.py.
// ---------------------------------------------------------------------------------------------------------
template <int TBITS>
struct ExpSmemConst;
template <>
struct ExpSmemConst<10> {
  static constexpr double SCALE = 1477.319721870298529136562873346117549;   // 1024 / ln 2
  static constexpr double STEP = 6.769015435155715912277655808101410987e-4;  // ln 2 / 1024
  static __device__ __forceinline__ double poly(double d) {
    double q = fma(d, 1.66666667859884626e-01, 5.00000003579653907e-01);
    q = fma(q, d, 1.0);
    return fma(q, d, 1.0);
  }
};
template <int TBITS>
__device__ __forceinline__ double fast_exp_core_smem(double x, const int2* tab) {
  const double MAGIC = 6755399441055744.0;
  double t = fma(x, ExpSmemConst<TBITS>::SCALE, MAGIC);
  int n = __double2loint(t);
  double nf = t - MAGIC;
  double d = fma(nf, -ExpSmemConst<TBITS>::STEP, x);
  double q = ExpSmemConst<TBITS>::poly(d);
  const int2 e = tab[n & ((1 << TBITS) - 1)];
  return __hiloint2double(e.y + (n << (20 - TBITS)), e.x) * q;
}
template <int TBITS>
__device__ __forceinline__ double fast_exp_neg_core_smem(double a, const int2* tab) {
  const double MAGIC = 6755399441055744.0;
  double t = fma(a, -ExpSmemConst<TBITS>::SCALE, MAGIC);
  int n = __double2loint(t);
  double nf = t - MAGIC;
  double d = fma(nf, -ExpSmemConst<TBITS>::STEP, -a);
  double q = ExpSmemConst<TBITS>::poly(d);
  const int2 e = tab[n & ((1 << TBITS) - 1)];
  return __hiloint2double(e.y + (n << (20 - TBITS)), e.x) * q;
}

// sqrt + exp(-sqrt) of the Matern family in one routine with the range reduction started EARLY: the table index n and
// nf = round(a 2^TBITS / ln 2) are taken from the seed-accurate g = u y (relative error 2^-22.9, i.e. < 0.01 table
// steps for a < 50) while the third-order sqrt step still runs, so the table load and two FP64 instructions leave the
// dependent chain sqrt -> reduction -> polynomial.  d = -a - nf STEP is formed from the FINAL a: the only effect of
// the approximate n is |d| <= 0.51 STEP instead of 0.5 (polynomial error x 1.1).  Same instruction count.
// MPOLY: 0 = exp(-a) alone (Matern-1/2), 1 = (1 + a) exp(-a) (3/2), 2 = (1 + a + u / 3) exp(-a) (5/2, u = a^2).  The
// Matern polynomial is multiplied into the TABLE value while the exp polynomial is still being evaluated, so the
// result is one multiply after the last Horner step.
template <int TBITS, int MPOLY>
__device__ __forceinline__ double fast_matern_early(double u, const int2* tab) {
  double y;
  asm("{\n"
      ".reg .b32 ulo, uhi, yhi;\n"
      ".reg .f64 yy;\n"
      "rsqrt.approx.ftz.f64 yy, %1;\n"
      "mov.b64 {ulo, yhi}, yy;\n"
      "mov.b64 {ulo, uhi}, %1;\n"
      "mov.b64 %0, {ulo, yhi};\n"
      "}\n"
      : "=d"(y)
      : "d"(u));
  const double MAGIC = 6755399441055744.0;
  const double g = u * y;
  const double t = fma(g, -ExpSmemConst<TBITS>::SCALE, MAGIC);
  const int n = __double2loint(t);
  const double nf = t - MAGIC;
  const int2 e = tab[n & ((1 << TBITS) - 1)];
  const double r = fma(-g, y, 1.0);
  const double p = fma(r, 0.375, 0.5);
  const double gr = g * r;
  const double a = fma(gr, p, g);
  const double d = fma(nf, -ExpSmemConst<TBITS>::STEP, -a);
  double tv = __hiloint2double(e.y + (n << (20 - TBITS)), e.x);
  if constexpr (MPOLY == 1) tv *= 1.0 + a;
  if constexpr (MPOLY == 2) tv *= fma(u, 1.0 / 3.0, 1.0 + a);
  return tv * ExpSmemConst<TBITS>::poly(d);
}

template <int TBITS>
__device__ __forceinline__ double fast_sqrt_exp_neg_early(double u, const int2* tab, double& a_out) {
  double y;
  asm("{\n"
      ".reg .b32 ulo, uhi, yhi;\n"
      ".reg .f64 yy;\n"
      "rsqrt.approx.ftz.f64 yy, %1;\n"
      "mov.b64 {ulo, yhi}, yy;\n"
      "mov.b64 {ulo, uhi}, %1;\n"
      "mov.b64 %0, {ulo, yhi};\n"
      "}\n"
      : "=d"(y)
      : "d"(u));
  const double MAGIC = 6755399441055744.0;
  const double g = u * y;
  const double t = fma(g, -ExpSmemConst<TBITS>::SCALE, MAGIC);
  const int n = __double2loint(t);
  const double nf = t - MAGIC;
  const int2 e = tab[n & ((1 << TBITS) - 1)];
  const double r = fma(-g, y, 1.0);
  const double p = fma(r, 0.375, 0.5);
  const double gr = g * r;
  const double a = fma(gr, p, g);
  a_out = a;
  const double d = fma(nf, -ExpSmemConst<TBITS>::STEP, -a);
  const double q = ExpSmemConst<TBITS>::poly(d);
  return __hiloint2double(e.y + (n << (20 - TBITS)), e.x) * q;
}

// Fast kernel value on r2 (variance is applied by the caller once per row, not per entry).
template <int KIND>
__device__ __forceinline__ double kernel_value_fast_unit(double r2, const FastExpTable& tab) {
  if (KIND == CGGP_SE) {
    return fast_exp(-0.5 * r2, tab);
  }
  // max(r2, 1e-36) as a signed compare of the high words (ALU pipe; negative r2 has a negative high word)
  double rc = (__double2hiint(r2) < 0x38754484) ? 1e-36 : r2;
  double r = fast_sqrt_pos(rc);
  if (KIND == CGGP_MATERN12) {
    return fast_exp(-r, tab);
  } else if (KIND == CGGP_MATERN32) {
    double s = 1.7320508075688772 * r;
    return (1.0 + s) * fast_exp(-s, tab);
  } else {
    double s = 2.23606797749979 * r;
    double poly = fma(5.0 / 3.0, rc, 1.0 + s);  // r*r == rc up to one rounding
    return poly * fast_exp(-s, tab);
  }
}
