// TEST INFRASTRUCTURE ONLY -- CPU oracle for the unconfined_b200 parity tests.
//
// PARITY UNPINNED: the reference (klkuhlm/unconfined) ships no golden outputs and
// no Fortran compiler exists in this environment, so this restatement cannot be
// checked against the reference binary.  It is pinned only at component level
// (J0 zeros vs mishra-neuman/malama-sp/besJ0zeros.dat, K0/K1 vs scipy's Amos,
// quadrature/inversion identities) -- see tests/ and DESIGN.md.
//
// Nothing under unconfined_b200/ may include, link or call this file.  Only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs use it, and only as the checker / CPU baseline.
//
// This header: complex arithmetic under GCC's Fortran rules (-fcx-fortran-rules:
// 4-multiply product, Smith division without NaN/Inf rescue, real operands promoted
// to complex before * and /), and thin wrappers over the glibc routines gfortran
// lowers its intrinsics to (csqrt/ccosh/csinh/cexp/clog/cabs, j0/j1, lgamma).
#pragma once
#include <cmath>
#include <limits>

// glibc's C99 complex routines (in C++ <complex.h> maps to <complex>, so declare them)
extern "C" {
double _Complex csqrt(double _Complex) noexcept;
double _Complex ccosh(double _Complex) noexcept;
double _Complex csinh(double _Complex) noexcept;
double _Complex cexp(double _Complex) noexcept;
double _Complex clog(double _Complex) noexcept;
double cabs(double _Complex) noexcept;
long double _Complex csqrtl(long double _Complex) noexcept;
long double _Complex ccoshl(long double _Complex) noexcept;
long double _Complex csinhl(long double _Complex) noexcept;
long double _Complex cexpl(long double _Complex) noexcept;
long double _Complex clogl(long double _Complex) noexcept;
long double cabsl(long double _Complex) noexcept;
}

namespace orc {

// ---- optional "other libm" jitter ------------------------------------------
// When enabled (orc_set_jitter), every result of the glibc routines below is
// multiplied by (1 + u*ulps*2^-53), u in [-1,1] a hash of the argument bits and a
// seed.  This emulates an equally valid libm whose last bits differ (the CUDA math
// library is 1-2 ulp vs glibc's <1 ulp) and is used by the tests to measure, per
// output point, how far rounding noise alone moves the reference's own result.
struct jitter_state { double ulps; unsigned long long seed; };
inline jitter_state &jitter() { static jitter_state j = {0.0, 0ULL}; return j; }
inline double jitter_u(double a, double b, unsigned salt) {
  unsigned long long x, y;
  __builtin_memcpy(&x, &a, 8); __builtin_memcpy(&y, &b, 8);
  unsigned long long h = x * 0x9E3779B97F4A7C15ULL ^ (y + 0xD1B54A32D192ED03ULL) * 0xBF58476D1CE4E5B9ULL
                         ^ (jitter().seed + salt) * 0x94D049BB133111EBULL;
  h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ULL; h ^= h >> 32;
  return ((double)(h >> 11) * (1.0 / 9007199254740992.0)) * 2.0 - 1.0;
}
template <class T> inline T jit(T v, double a, double b, unsigned salt) {
  if (jitter().ulps == 0.0) return v;
  return v * (T)(1.0 + jitter_u(a, b, salt) * jitter().ulps * 1.1102230246251565e-16);
}

template <class T> struct cx {
  T re, im;
  cx() : re(0), im(0) {}
  cx(T r, T i) : re(r), im(i) {}
  explicit cx(T r) : re(r), im(0) {}
};

template <class T> inline cx<T> operator+(cx<T> a, cx<T> b) { return {a.re + b.re, a.im + b.im}; }
template <class T> inline cx<T> operator-(cx<T> a, cx<T> b) { return {a.re - b.re, a.im - b.im}; }
template <class T> inline cx<T> operator-(cx<T> a) { return {-a.re, -a.im}; }
// Under jitter, complex products/quotients are also perturbed by ~half an ulp of their
// MODULUS per component: an FMA-contracting build (gfortran -O3 -march=native, nvcc)
// rounds a.re*b.re - a.im*b.im differently from this unfused code, and when the two
// products nearly cancel that difference is large relative to the component itself.
template <class T> inline cx<T> jit_mod(cx<T> r, double a, double b, unsigned salt) {
  if (jitter().ulps == 0.0) return r;
  T m = std::fabs(r.re) > std::fabs(r.im) ? std::fabs(r.re) : std::fabs(r.im);
  if (!(m <= std::numeric_limits<T>::max())) return r;
  const double e = 0.5 * jitter().ulps * 1.1102230246251565e-16;
  return {r.re + (T)(jitter_u(a, b, salt) * e) * m, r.im + (T)(jitter_u(b, a, salt + 7) * e) * m};
}
// complex*complex, plain 4-multiply form (GCC tree-complex, Fortran rules)
template <class T> inline cx<T> operator*(cx<T> a, cx<T> b) {
  cx<T> r = {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
  return jit_mod(r, (double)r.re, (double)r.im, 60);
}
// complex/complex, GCC expand_complex_div_wide (Smith, no rescue)
template <class T> inline cx<T> operator/(cx<T> a, cx<T> b) {
  cx<T> r;
  if (std::fabs(b.re) < std::fabs(b.im)) {
    T ratio = b.re / b.im;
    T div = (b.re * ratio) + b.im;
    T tr = (a.re * ratio) + a.im;
    T ti = (a.im * ratio) - a.re;
    r = {tr / div, ti / div};
  } else {
    T ratio = b.im / b.re;
    T div = (b.im * ratio) + b.re;
    T tr = (a.im * ratio) + a.re;
    T ti = a.im - (a.re * ratio);
    r = {tr / div, ti / div};
  }
  return jit_mod(r, (double)r.re, (double)r.im, 70);
}
// mixed real/complex: Fortran converts the real operand to complex first
template <class T> inline cx<T> operator*(cx<T> a, T x) { return a * cx<T>(x, T(0)); }
template <class T> inline cx<T> operator*(T x, cx<T> a) { return cx<T>(x, T(0)) * a; }
template <class T> inline cx<T> operator/(cx<T> a, T x) { return a / cx<T>(x, T(0)); }
template <class T> inline cx<T> operator/(T x, cx<T> a) { return cx<T>(x, T(0)) / a; }
template <class T> inline cx<T> operator+(cx<T> a, T x) { return {a.re + x, a.im + T(0)}; }
template <class T> inline cx<T> operator+(T x, cx<T> a) { return {x + a.re, T(0) + a.im}; }
template <class T> inline cx<T> operator-(T x, cx<T> a) { return {x - a.re, T(0) - a.im}; }
template <class T> inline cx<T> operator-(cx<T> a, T x) { return {a.re - x, a.im - T(0)}; }
template <class T> inline cx<T> conj(cx<T> a) { return {a.re, -a.im}; }

// ---- libm bindings (double -> glibc double routines, long double -> *l) ----
template <class T> struct lm;
template <> struct lm<double> {
  typedef double _Complex C;
  static C mk(cx<double> z) { C c; __real__ c = z.re; __imag__ c = z.im; return c; }
  static cx<double> un(C c) { return {__real__ c, __imag__ c}; }
  static cx<double> jc(cx<double> r, cx<double> z, unsigned salt) {
    return {jit(r.re, z.re, z.im, salt), jit(r.im, z.re, z.im, salt + 1)};
  }
  static cx<double> sqrt(cx<double> z) { return jc(un(::csqrt(mk(z))), z, 10); }
  static cx<double> cosh(cx<double> z) { return jc(un(::ccosh(mk(z))), z, 20); }
  static cx<double> sinh(cx<double> z) { return jc(un(::csinh(mk(z))), z, 30); }
  static cx<double> exp(cx<double> z) { return jc(un(::cexp(mk(z))), z, 40); }
  static cx<double> log(cx<double> z) { return un(::clog(mk(z))); }
  static double abs(cx<double> z) { return ::cabs(mk(z)); }
  // J0: relative jitter plus an absolute one of 0.25 ulp(1)*amplitude (near its zeros
  // any implementation is only absolutely accurate)
  static double j0(double x) {
    double v = ::j0(x);
    if (jitter().ulps == 0.0) return v;
    double amp = 1.0 / std::sqrt(1.0 + std::fabs(x));
    return jit(v, x, 0.0, 50) + 0.25 * jitter_u(x, 1.0, 51) * jitter().ulps * 1.1102230246251565e-16 * amp;
  }
  static double j1(double x) { return ::j1(x); }
  static double lgamma(double x) { return ::lgamma(x); }
};
template <> struct lm<long double> {
  typedef long double _Complex C;
  static C mk(cx<long double> z) { C c; __real__ c = z.re; __imag__ c = z.im; return c; }
  static cx<long double> un(C c) { return {__real__ c, __imag__ c}; }
  static cx<long double> sqrt(cx<long double> z) { return un(::csqrtl(mk(z))); }
  static cx<long double> cosh(cx<long double> z) { return un(::ccoshl(mk(z))); }
  static cx<long double> sinh(cx<long double> z) { return un(::csinhl(mk(z))); }
  static cx<long double> exp(cx<long double> z) { return un(::cexpl(mk(z))); }
  static cx<long double> log(cx<long double> z) { return un(::clogl(mk(z))); }
  static long double abs(cx<long double> z) { return ::cabsl(mk(z)); }
  static long double j0(long double x) { return ::j0l(x); }
  static long double j1(long double x) { return ::j1l(x); }
  static long double lgamma(long double x) { return ::lgammal(x); }
};

template <class T> inline cx<T> csqrt(cx<T> z) { return lm<T>::sqrt(z); }
template <class T> inline cx<T> ccosh(cx<T> z) { return lm<T>::cosh(z); }
template <class T> inline cx<T> csinh(cx<T> z) { return lm<T>::sinh(z); }
template <class T> inline cx<T> cexp(cx<T> z) { return lm<T>::exp(z); }
template <class T> inline cx<T> clog(cx<T> z) { return lm<T>::log(z); }
template <class T> inline T cabs(cx<T> z) { return lm<T>::abs(z); }

// utility.f90:59-64  is_finite: .not.(isnan(abs(x)) .or. abs(x) > huge(abs(x)))
template <class T> inline bool is_finite(cx<T> z) {
  T a = cabs(z);
  return !(std::isnan(a) || a > std::numeric_limits<T>::max());
}

}  // namespace orc
