// Device-side FP64 complex arithmetic and special functions for the Laplace-Hankel
// kernels (sm_100a).  Two families live here:
//   * "literal" routines that reproduce, operation for operation, what the reference
//     build executes: GCC's Fortran-rules complex * and / (4-multiply product, Smith
//     division with no NaN/Inf rescue, real operands promoted to complex first) and
//     glibc's csqrt/ccosh/csinh/cexp component formulas including their overflow
//     staging (laplace_hankel_solutions.f90 uses the intrinsics sqrt/cosh/sinh/exp on
//     complex(8), which gfortran lowers to those glibc routines).  Where the reference
//     overflows to Inf/NaN these do too -- that flow is part of the contract
//     (integration.f90:140-160 truncates the Wynn series at the first non-finite term).
//   * j0_dev(): J0 from the generated table (tools/gen_j0_table.py), replacing the
//     bessel_j0 intrinsic at laplace_hankel_solutions.f90:118.
#pragma once
#include <cfloat>
#include <math.h>
#define UNC_J0_QUAL __device__
#include "j0_table.h"

namespace unc {

// 16-byte alignment: shared/local/global accesses of a cplx are single 128-bit transactions
struct __align__(16) cplx {
  double re, im;
};

__host__ __device__ __forceinline__ cplx mk(double r, double i) { cplx z; z.re = r; z.im = i; return z; }
__host__ __device__ __forceinline__ cplx operator+(cplx a, cplx b) { return mk(a.re + b.re, a.im + b.im); }
__device__ __forceinline__ cplx operator-(cplx a, cplx b) { return mk(a.re - b.re, a.im - b.im); }
__device__ __forceinline__ cplx operator-(cplx a) { return mk(-a.re, -a.im); }
__device__ __forceinline__ cplx operator*(cplx a, cplx b) {
  return mk(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
}
// GCC expand_complex_div_wide
__device__ __forceinline__ cplx operator/(cplx a, cplx b) {
  if (fabs(b.re) < fabs(b.im)) {
    double ratio = b.re / b.im;
    double div = (b.re * ratio) + b.im;
    double tr = (a.re * ratio) + a.im;
    double ti = (a.im * ratio) - a.re;
    return mk(tr / div, ti / div);
  } else {
    double ratio = b.im / b.re;
    double div = (b.im * ratio) + b.re;
    double tr = (a.im * ratio) + a.re;
    double ti = a.im - (a.re * ratio);
    return mk(tr / div, ti / div);
  }
}
// mixed-mode: the real operand is converted to (x,0) first (Fortran semantics)
__device__ __forceinline__ cplx operator*(cplx a, double x) { return a * mk(x, 0.0); }
__device__ __forceinline__ cplx operator*(double x, cplx a) { return mk(x, 0.0) * a; }
__device__ __forceinline__ cplx operator/(cplx a, double x) { return a / mk(x, 0.0); }
__device__ __forceinline__ cplx operator/(double x, cplx a) { return mk(x, 0.0) / a; }
__device__ __forceinline__ cplx operator+(cplx a, double x) { return mk(a.re + x, a.im + 0.0); }
__device__ __forceinline__ cplx operator+(double x, cplx a) { return mk(x + a.re, 0.0 + a.im); }
__device__ __forceinline__ cplx operator-(double x, cplx a) { return mk(x - a.re, 0.0 - a.im); }
__device__ __forceinline__ cplx operator-(cplx a, double x) { return mk(a.re - x, a.im - 0.0); }
__device__ __forceinline__ cplx conj(cplx a) { return mk(a.re, -a.im); }
// plain componentwise scaling (used where all operands are known finite)
__device__ __forceinline__ cplx scale(cplx a, double x) { return mk(a.re * x, a.im * x); }
__device__ __forceinline__ cplx fma_acc(double w, cplx f, cplx acc) {
  return mk(fma(w, f.re, acc.re), fma(w, f.im, acc.im));
}

__host__ __device__ __forceinline__ double cabs_d(cplx z) { return hypot(z.re, z.im); }
// utility.f90:59-64
__host__ __device__ __forceinline__ bool is_finite_c(cplx z) {
  double a = hypot(z.re, z.im);
  return !(isnan(a) || a > DBL_MAX);
}

// same predicate without the hypot call when both components are far from overflow; the rare
// rest goes through a real call (hypot is ~100 instructions, and this sits in many unrolled loops)
__host__ __device__ __noinline__ bool is_finite_slowc(cplx z) { return is_finite_c(z); }
__host__ __device__ __forceinline__ bool is_finite_fastc(cplx z) {
  if (fabs(z.re) < 1e150 && fabs(z.im) < 1e150) return true;   // false for NaN
  return is_finite_slowc(z);
}

// ---- glibc-shaped complex elementary functions (finite arguments) ----------
#define UNC_EXP_T 709  /* (int)((DBL_MAX_EXP-1)*M_LN2) */

__device__ __forceinline__ void sincos_g(double y, double *s, double *c) {
  if (fabs(y) > DBL_MIN) sincos(y, s, c);
  else { *s = y; *c = 1.0; }
}

// glibc s_ccosh_template.c (finite/finite branch)
__device__ __noinline__ cplx ccosh_g(cplx x) {
  const double t = (double)UNC_EXP_T;
  double s, c;
  sincos_g(x.im, &s, &c);
  if (fabs(x.re) > t) {
    const double exp_t = exp(t);
    double rx = fabs(x.re);
    if (signbit(x.re)) s = -s;
    rx -= t;
    s *= exp_t / 2;
    c *= exp_t / 2;
    if (rx > t) { rx -= t; s *= exp_t; c *= exp_t; }
    if (rx > t) return mk(DBL_MAX * c, DBL_MAX * s);
    double ev = exp(rx);
    return mk(ev * c, ev * s);
  }
  return mk(cosh(x.re) * c, sinh(x.re) * s);
}

// glibc s_csinh_template.c
__device__ __noinline__ cplx csinh_g(cplx x) {
  const double t = (double)UNC_EXP_T;
  const bool negate = signbit(x.re);
  double rx = fabs(x.re);
  double s, c;
  sincos_g(x.im, &s, &c);
  if (negate) c = -c;
  if (rx > t) {
    const double exp_t = exp(t);
    rx -= t;
    s *= exp_t / 2;
    c *= exp_t / 2;
    if (rx > t) { rx -= t; s *= exp_t; c *= exp_t; }
    if (rx > t) return mk(DBL_MAX * c, DBL_MAX * s);
    double ev = exp(rx);
    return mk(ev * c, ev * s);
  }
  return mk(sinh(rx) * c, cosh(rx) * s);
}

// glibc s_cexp_template.c
__device__ __noinline__ cplx cexp_g(cplx x) {
  const double t = (double)UNC_EXP_T;
  double s, c;
  sincos_g(x.im, &s, &c);
  double rx = x.re;
  if (rx > t) {
    const double exp_t = exp(t);
    rx -= t;
    s *= exp_t;
    c *= exp_t;
    if (rx > t) { rx -= t; s *= exp_t; c *= exp_t; }
  }
  if (rx > t) return mk(DBL_MAX * c, DBL_MAX * s);
  double ev = exp(rx);
  return mk(ev * c, ev * s);
}

// glibc s_csqrt_template.c (finite arguments; scaling branches for huge/tiny inputs kept)
__device__ __noinline__ cplx csqrt_g(cplx x) {
  double re = x.re, im = x.im;
  if (im == 0.0) {
    if (re < 0.0) return mk(0.0, copysign(sqrt(-re), im));
    return mk(fabs(sqrt(re)), copysign(0.0, im));
  }
  if (re == 0.0) {
    double r = (fabs(im) >= 2.0 * DBL_MIN) ? sqrt(0.5 * fabs(im)) : 0.5 * sqrt(2.0 * fabs(im));
    return mk(r, copysign(r, im));
  }
  int sc = 0;
  if (fabs(re) > DBL_MAX / 4) { sc = 1; re = scalbn(re, -2); im = scalbn(im, -2); }
  else if (fabs(im) > DBL_MAX / 4) { sc = 1; im = scalbn(im, -2); re = (fabs(re) >= 4 * DBL_MIN) ? scalbn(re, -2) : 0.0; }
  else if (fabs(re) < 2 * DBL_MIN && fabs(im) < 2 * DBL_MIN) { sc = -((DBL_MANT_DIG + 1) / 2); re = scalbn(re, -2 * sc); im = scalbn(im, -2 * sc); }
  double d = hypot(re, im);
  double r, s;
  if (re > 0) {
    r = sqrt(0.5 * (d + re));
    if (sc == 1 && fabs(im) < 1) { s = im / r; r = scalbn(r, sc); sc = 0; }
    else s = 0.5 * (im / r);
  } else {
    s = sqrt(0.5 * (d - re));
    if (sc == 1 && fabs(im) < 1) { r = fabs(im / s); s = scalbn(s, sc); sc = 0; }
    else r = fabs(0.5 * (im / s));
  }
  if (sc) { r = scalbn(r, sc); s = scalbn(s, sc); }
  return mk(r, copysign(s, im));
}

// principal log; glibc clog uses log1p-style care near |z|=1, where the absolute
// error of log(hypot) is already ~1 ulp of 1 -- enough for cbknu's smu=log(2/z)
__device__ __forceinline__ cplx clog_g(cplx z) { return mk(log(hypot(z.re, z.im)), atan2(z.im, z.re)); }

// ---- J0 ---------------------------------------------------------------------
// x < zeros[NINT] (~99.75): degree-21 polynomial on the interval between consecutive
// zeros (abs err ~1e-16); beyond: Hankel asymptotic expansion (A&S 9.2.1, 9.2.9-10).
__device__ __noinline__ double j0_dev(double x) {
  x = fabs(x);
  if (x < UNC_J0_ZEROS[UNC_J0_NINT]) {
    int k = (int)fma(x, 0.318309886183790671538, 0.25);  // zeros ~ (k - 1/4) pi
    k = min(k, UNC_J0_NINT - 1);
    if (x < UNC_J0_ZEROS[k]) k -= 1;
    else if (x >= UNC_J0_ZEROS[k + 1]) k += 1;
    k = max(0, min(k, UNC_J0_NINT - 1));
    const double u = (x - UNC_J0_MID[k]) * UNC_J0_IHALF[k];
    const double *c = &UNC_J0_COEF[k * (UNC_J0_DEG + 1)];
    double acc = c[UNC_J0_DEG];
#pragma unroll
    for (int i = UNC_J0_DEG - 1; i >= 0; --i) acc = fma(acc, u, c[i]);
    return acc;
  }
  const double y = 1.0 / x, y2 = y * y;
  double P = UNC_J0_PC[UNC_J0_NASY - 1], Q = UNC_J0_QC[UNC_J0_NASY - 1];
#pragma unroll
  for (int i = UNC_J0_NASY - 2; i >= 0; --i) { P = fma(P, y2, UNC_J0_PC[i]); Q = fma(Q, y2, UNC_J0_QC[i]); }
  Q *= y;
  double s, c;
  sincos(x, &s, &c);
  // sqrt(2/(pi x)) (P cos(x-pi/4) - Q sin(x-pi/4)),  cos(x-pi/4)=(c+s)/sqrt2, sin(x-pi/4)=(s-c)/sqrt2
  return 0.564189583547756286948 * sqrt(y) * (P * (c + s) - Q * (s - c));
}

}  // namespace unc
