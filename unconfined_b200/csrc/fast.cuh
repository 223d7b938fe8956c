// Fast evaluation of the Laplace-Hankel kernels (laplace_hankel_solutions.f90:30-116)
// for the regime where nothing in the reference's formulas can overflow.
//
// Every model 0-5 kernel, for fixed (a,p), is  f(z) = k0 + cp*exp(eta z) + cm*exp(-eta z)
// with (k0,cp,cm) depending on the layer of z only.  With h = bD/2, m = (dD1+lD1)/2
// (dD1 = 1-dD, lD1 = 1-lD) the three layer functions of hantush (:175-198),
//   below  g3*cosh(eta z),  g3 = exp(-eta lD1) - (ff1 + exp(-eta) ff2)/sinh(eta)
//   beside 1 - (ff1 cosh(eta z) + ff2 cosh(eta(1-z)))/sinh(eta)
//   above  cosh(eta(dD1-z)) - (ff1 cosh(eta z) + ff2 cosh(eta(1-z)))/sinh(eta)
// (ff1 = sinh(eta dD), ff2 = sinh(eta lD1)) reduce algebraically to
//   below  2 sinh(eta h) cosh(eta(1-m)) / sinh(eta) * cosh(eta z)
//   above  2 sinh(eta h) cosh(eta m)     / sinh(eta) * cosh(eta(1-z))
// (the classical Hantush forms), and the water-table term of models 3-5 (:84-92),
//   - udp(1) cosh(eta z)/D  or  - udp(1) exp(eta(z-1))/D',
// only adds to cp/cm.  So the ~10 complex sinh/cosh/exp per (a,p,z) of the reference
// collapse to 3-5 complex exponentials per (a,p) plus ONE per (a,p,z).  The algebra is
// exact; only rounding differs, and the product forms avoid the cancellation the
// reference's own expressions suffer (DESIGN.md "parity" discusses the noise floor).
// Valid while Re(eta)*m <= UNC_FAST_EXP_MAX (fast_eta_max), where neither formulation can overflow;
// beyond that the caller uses the literal path, which overflows exactly where the
// reference does.
#pragma once
#include "params.cuh"
#include "cmath.cuh"
#include <cstring>

namespace unc {

// Largest Re(eta)*m for which neither formulation overflows: the caller passes
// eta_max = UNC_FAST_EXP_MAX / m, m = the largest exponent multiplier the REFERENCE's
// literal formulas form for the z-values served (fast_eta_max() below).
#define UNC_FAST_EXP_MAX 700.0

// m: sinh(eta) -> 1; udp(zD=1) of models 3/5 and the layer-3 branch form
// sinh(eta dD)*cosh(eta z) -> up to 1+dD (laplace_hankel_solutions.f90:179,81); layer-2
// products stay below 1; |z| itself for callers outside 0<=zD<=1.
// Error-budget switches (tools/error_budget.py builds one library per switch and tabulates
// |gpu - oracle| on the BASELINE decks; never defined in the product build):
//   UNC_BUDGET_LITERAL   no closed forms: every abscissa takes the literal path
//   UNC_BUDGET_LIBM      CUDA libm exp/sincos instead of exp_pm / sincos_q
//   UNC_BUDGET_IEEE_DIV  IEEE division instead of rcp.approx + 2 Newton steps
//   UNC_BUDGET_SEQSUM    point kernel: abscissae summed sequentially in the reference's order
//   UNC_BUDGET_NEVILLE   point kernel: R level sums + extraptozero on the device (implies SEQSUM)
__host__ __device__ __forceinline__ double fast_eta_max(const DevParams &P, int lay_mask, double zabs_max) {
#ifdef UNC_BUDGET_LITERAL
  return -1.0;
#endif
  double m = 1.0;
  if (P.model == 3 || P.model == 5 || (lay_mask & 4)) m = 1.0 + P.dD;
  if (P.model == 6) m = 1.03;   // Delta0 = eta sinh(eta) - u cosh(eta): headroom for the factors eta, u
  if (zabs_max > m) m = zabs_max;
  return UNC_FAST_EXP_MAX / m;
}

// Polynomial coefficients live in __constant__ memory so that DFMA takes them as
// c[bank][offset] operands (ncu showed 35% of all issued instructions were UMOV/IMAD
// constant materialisation when they were immediates).
#define UNC_KEXP_INIT { \
    1.4426950408889634074,             \
    6.93147180369123816490e-01,        \
    1.90821492927058770002e-10,        \
    6755399441055744.0,                \
    2.08767569878680989792e-09,        \
    2.75573192239858906526e-07,        \
    2.48015873015873015873e-05,        \
    1.38888888888888888889e-03,        \
    4.16666666666666666667e-02,        \
    1.60590438368216145994e-10,        \
    2.50521083854417187751e-08,        \
    2.75573192239858906526e-06,        \
    1.98412698412698412698e-04,        \
    8.33333333333333333333e-03,        \
    1.66666666666666666667e-01,        \
    0.0}
__constant__ double KEXP_D[16] = UNC_KEXP_INIT;
static const double KEXP_H[16] = UNC_KEXP_INIT;

// sin r = r + r^3 g(r^2), cos r = 1 - r^2/2 + r^4 h(r^2) on |r| <= pi/4
// (tools/gen_sincos_poly.py: abs err 7e-17 / 1.2e-16)
#define UNC_KTRIG_INIT { \
    6.36619772367581382433e-01,        \
    1.57079632679489655800e+00,        \
    6.12323399573676603587e-17,        \
    6755399441055744.0,                \
    -1.66666666666666657415e-01, 8.33333333333090113537e-03, -1.98412698366860079397e-04,    \
    2.75573160649867732950e-06, -2.50511240313820812782e-08, 1.59175674789067991500e-10, \
    4.16666666666666643537e-02, -1.38888888888873671991e-03, 2.48015872987202915510e-05,     \
    -2.75573172482293668695e-07, 2.08761413802949554280e-09, -1.13822809577026335532e-11}
__constant__ double KTRIG_D[16] = UNC_KTRIG_INIT;
static const double KTRIG_H[16] = UNC_KTRIG_INIT;


#ifdef __CUDA_ARCH__
#define KEXP KEXP_D
#define KTRIG KTRIG_D
#define UNC_LOINT(x) __double2loint(x)
#define UNC_HIINT(x) __double2hiint(x)
#define UNC_HILO2D(h, l) __hiloint2double(h, l)
#else
// host mirrors: the same functions compile for the CPU so that tests/ can check the
// fast-path algebra against the oracle without a GPU (tests/hostcheck); never used by
// the product, whose entry points only launch kernels.
#define KEXP KEXP_H
#define KTRIG KTRIG_H
static inline int unc_loint_h(double x) { long long b; std::memcpy(&b, &x, 8); return (int)(b & 0xffffffffLL); }
static inline int unc_hiint_h(double x) { long long b; std::memcpy(&b, &x, 8); return (int)(b >> 32); }
static inline double unc_hilo2d_h(int h, int l) {
  long long b = ((long long)h << 32) | (unsigned int)l; double x; std::memcpy(&x, &b, 8); return x;
}
#define UNC_LOINT(x) unc_loint_h(x)
#define UNC_HIINT(x) unc_hiint_h(x)
#define UNC_HILO2D(h, l) unc_hilo2d_h(h, l)
#endif

// exp(+-x) from one range reduction: x = k ln2 + r, |r| <= ln2/2, exp(+-r) = cosh r +- sinh r
// (even/odd Taylor polynomials), scaled by 2^(+-k).  |x| <= ~700 on this path, so 2^k and
// 2^-k are normal doubles.  *c_out/*s_out = cosh r, sinh r and *k_out = k for callers that
// need cosh x / sinh x without cancellation.
__host__ __device__ __forceinline__ void exp_pm_core(double x, double *ep, double *em, double *c_out,
                                            double *s_out, int *k_out) {
#ifdef UNC_BUDGET_LIBM
  *ep = exp(x); *em = exp(-x); *c_out = cosh(x); *s_out = sinh(x); *k_out = 0;
  return;
#endif
  const double km = fma(x, KEXP[0], KEXP[3]);
  const int k = UNC_LOINT(km);
  const double kf = km - KEXP[3];
  double r = fma(-kf, KEXP[1], x);
  r = fma(-kf, KEXP[2], r);
  const double r2 = r * r;
  double c = fma(KEXP[4], r2, KEXP[5]);
  c = fma(c, r2, KEXP[6]);
  c = fma(c, r2, KEXP[7]);
  c = fma(c, r2, KEXP[8]);
  c = fma(c, r2, 0.5);
  c = fma(c, r2, 1.0);
  double s = fma(KEXP[9], r2, KEXP[10]);
  s = fma(s, r2, KEXP[11]);
  s = fma(s, r2, KEXP[12]);
  s = fma(s, r2, KEXP[13]);
  s = fma(s, r2, KEXP[14]);
  s = fma(s * r2, r, r);
  const double sp = UNC_HILO2D((1023 + k) << 20, 0);
  const double sm = UNC_HILO2D((1023 - k) << 20, 0);
  *ep = (c + s) * sp;
  *em = (c - s) * sm;
  *c_out = c;
  *s_out = s;
  *k_out = k;
}

struct rexp {
  double ep, em, ch, sh;
};
__host__ __device__ __forceinline__ rexp exp_pm(double x) {
  rexp o;
  double c, s;
  int k;
  exp_pm_core(x, &o.ep, &o.em, &c, &s, &k);
  // k == 0: the polynomials ARE cosh/sinh (no cancellation for small x)
  o.ch = (k == 0) ? c : 0.5 * (o.ep + o.em);
  o.sh = (k == 0) ? s : 0.5 * (o.ep - o.em);
  return o;
}

// sin and cos of y, |y| < ~2^20 (here |Im(eta) z| is at most a few thousand):
// two-term Cody-Waite reduction by pi/2 (the FMA keeps k*pi/2_hi exact), kernel
// polynomials, quadrant fix-up on the sign/high words.  <= ~1 ulp of 1 absolute.
__host__ __device__ __forceinline__ void sincos_q(double y, double *sn, double *cs) {
#ifdef UNC_BUDGET_LIBM
  sincos(y, sn, cs);
  return;
#endif
  const double km = fma(y, KTRIG[0], KTRIG[3]);
  const int n = UNC_LOINT(km);
  const double kf = km - KTRIG[3];
  double r = fma(-kf, KTRIG[1], y);
  r = fma(-kf, KTRIG[2], r);
  const double t = r * r;
  double g = fma(KTRIG[9], t, KTRIG[8]);
  g = fma(g, t, KTRIG[7]);
  g = fma(g, t, KTRIG[6]);
  g = fma(g, t, KTRIG[5]);
  g = fma(g, t, KTRIG[4]);
  const double s = fma(g * t, r, r);
  double h = fma(KTRIG[15], t, KTRIG[14]);
  h = fma(h, t, KTRIG[13]);
  h = fma(h, t, KTRIG[12]);
  h = fma(h, t, KTRIG[11]);
  h = fma(h, t, KTRIG[10]);
  const double c = fma(h * t, t, fma(t, -0.5, 1.0));
  // quadrant n&3: sin = [s, c, -s, -c], cos = [c, -s, -c, s]
  const bool sw = n & 1;
  const double a = sw ? c : s, b = sw ? s : c;
  const int sa = (n & 2) << 30, sb = ((n + 1) & 2) << 30;
  *sn = UNC_HILO2D(UNC_HIINT(a) ^ sa, UNC_LOINT(a));
  *cs = UNC_HILO2D(UNC_HIINT(b) ^ sb, UNC_LOINT(b));
}

#ifndef UNC_RCP_CUBIC
#define UNC_RCP_CUBIC 1
#endif
// 1/x for x in the normal range: hardware approximation (MUFU.RCP64H, ~20 bits) + two
// Newton steps (4 DFMA) instead of the ~25-instruction IEEE division sequence; <= 1 ulp.
__host__ __device__ __forceinline__ double rcp_fast(double x) {
#if defined(__CUDA_ARCH__) && !defined(UNC_BUDGET_IEEE_DIV)
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#if UNC_RCP_CUBIC
  // one third-order step: y (1 + e + e^2), e = 1 - x y <= 2^-20  ->  residual e^3 < 2^-60
  const double e = fma(-x, y, 1.0);
  const double t = fma(e, e, e);
  return fma(y, t, y);
#else
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  return y;
#endif
#else
  return 1.0 / x;
#endif
}

// plain complex helpers for finite operands (no real->complex promotion)
__host__ __device__ __forceinline__ cplx cmulf(cplx a, cplx b) {
  return mk(fma(a.re, b.re, -(a.im * b.im)), fma(a.re, b.im, a.im * b.re));
}
__host__ __device__ __forceinline__ cplx crecipf(cplx b) {
  const double d = rcp_fast(fma(b.re, b.re, b.im * b.im));
  return mk(b.re * d, -(b.im * d));
}
__host__ __device__ __forceinline__ cplx cdivf(cplx a, cplx b) { return cmulf(a, crecipf(b)); }
__host__ __device__ __forceinline__ cplx caddf(cplx a, cplx b) { return mk(a.re + b.re, a.im + b.im); }
__host__ __device__ __forceinline__ cplx csubf(cplx a, cplx b) { return mk(a.re - b.re, a.im - b.im); }
__host__ __device__ __forceinline__ cplx cscalef(cplx a, double x) { return mk(a.re * x, a.im * x); }

// complex exp(+-w), cosh w, sinh w of w = eta*c
struct cbundle {
  cplx ep, em, ch, sh;
};
__host__ __device__ __forceinline__ cbundle cexp_bundle(double wr, double wi) {
  const rexp e = exp_pm(wr);
  double s, c;
  sincos_q(wi, &s, &c);
  cbundle b;
  b.ep = mk(e.ep * c, e.ep * s);
  b.em = mk(e.em * c, -(e.em * s));
  b.ch = mk(e.ch * c, e.sh * s);
  b.sh = mk(e.sh * c, e.ch * s);
  return b;
}

// sqrt of z with Re z > 0 (eta = sqrt((p+a^2)/kappa)), glibc's formula for that branch
__host__ __device__ __forceinline__ cplx csqrt_pos(cplx z) {
  if (z.im == 0.0) return mk(sqrt(z.re), 0.0);
  const double d = sqrt(fma(z.re, z.re, z.im * z.im));
  const double r = sqrt(0.5 * (d + z.re));
  return mk(r, 0.5 * (z.im * rcp_fast(r)));
}

struct Coef {  // f(z) = k0 + cp*exp(eta z) + cm*exp(-eta z)
  cplx k0, cp, cm;
};

// Per-(a,p) terms shared by every z.  lay_mask: bit (L-1) set if layer L occurs among
// the z-values served.  Returns false if Re(eta) exceeds the fast-path bound (the
// caller must then use the literal path for this abscissa).  `w` (quadrature weight
// times a*J0(a rD)) is folded into the coefficients.
//   aux  : model 3: sum_m 1/(1+p/gamma_m);  model 2: A0(p) = 2/(p CDw K0 + xi K1)
//   aux2 : model 2: p*tDb + 1
// MODEL >= 0: the model is a compile-time constant (the other models' branches vanish from the
// instantiation: less code, fewer live values); MODEL = -1: taken from P at run time.
// The coefficients of a layer go to sink.set(L, k0, cp, cm) as soon as they are final (the
// water-table term is formed before the layers are assembled), so that a sink that stores them
// -- the grid kernel's shared-memory stage -- does not carry 18 doubles to the end of the function.
struct CoefRegSink {   // the plain Coef[3] of the point kernel, the lanes<->z kernel and the host checks
  Coef *co;
  __host__ __device__ __forceinline__ void set(int L, cplx k0, cplx cp, cplx cm) const {
    co[L].k0 = k0; co[L].cp = cp; co[L].cm = cm;
  }
};
template <int MODEL, class Sink>
__host__ __device__ __forceinline__ bool ap_terms_fast_s(const DevParams &P, cplx p, cplx aux, cplx aux2,
                                                double a2, double w, int lay_mask, double eta_max,
                                                cplx *eta_out, const Sink &sink) {
  const int model = (MODEL >= 0) ? MODEL : P.model;
  const cplx pa = mk(p.re + a2, p.im);
  const cplx zero = mk(0.0, 0.0);
  if (model == 0) {
    const cplx th = cscalef(crecipf(pa), 2.0 * w);  // theis :122-131
    *eta_out = zero;
#pragma unroll
    for (int L = 0; L < 3; ++L) sink.set(L, th, zero, zero);
    return true;
  }
  const cplx eta = csqrt_pos(cscalef(pa, 1.0 / P.kappa));
  *eta_out = eta;
  if (!(eta.re <= eta_max && eta.im <= 2.0e5)) return false;  // overflow bound; sincos_q range
  const cbundle E1 = cexp_bundle(eta.re, eta.im);
  if (model == 6) {
    // mishraNeumanMalama (laplace_hankel_solutions.f90:404-442):
    //   2/(kappa eta^2) (1 + u/Delta0 cosh(eta z)),  Delta0 = eta sinh(eta) - u cosh(eta),
    //   u = u0 (1 - sqrt(1 + (eta1/u0)^2)),  eta1 = sqrt((p vartheta + a^2)/kappa).
    // u keeps the reference's form u0 (1 - v): for |eta1| << u0 its value is dominated by the
    // rounding of 1 + (eta1/u0)^2 and of the square root, which are the same IEEE operations
    // here, so the (large) cancellation noise of the reference is tracked, not "fixed".
    const cplx e1sq = cscalef(mk(fma(p.re, P.mn_vartheta, a2), p.im * P.mn_vartheta), 1.0 / P.kappa);
    const double iu02 = 1.0 / (P.mn_u0 * P.mn_u0);
    const cplx v = csqrt_pos(mk(1.0 + e1sq.re * iu02, e1sq.im * iu02));
    const cplx u = cscalef(mk(1.0 - v.re, -v.im), P.mn_u0);
    const cplx th = cscalef(crecipf(pa), 2.0 * w);       // 2/(kappa eta^2) = 2/(p + a^2)
    // u/Delta0.  |Delta0|^2 overflows at Re(eta) ~ 350 although Delta0 itself (and the
    // reference's Smith division) stays finite up to ~709: for Re(eta) > 20, sinh = cosh =
    // e^eta/2 to 2^-57, so Delta0 = e^eta (eta - u)/2 and u/Delta0 = 2 u e^-eta/(eta - u).
    cplx ud;
    if (eta.re > 20.0) ud = cscalef(cmulf(cdivf(u, csubf(eta, u)), E1.em), 2.0);
    else ud = cdivf(u, csubf(cmulf(eta, E1.sh), cmulf(u, E1.ch)));
    const cplx g = cscalef(cmulf(th, ud), 0.5);
#pragma unroll
    for (int L = 0; L < 3; ++L) sink.set(L, th, g, g);
    return true;
  }
  cplx K0;  // common prefactor of the layer functions, weight folded in
  if (model == 2) K0 = cscalef(cdivf(aux, cmulf(pa, aux2)), w / P.bD);            // uDf/bD :265-266,299
  else K0 = cscalef(crecipf(pa), 2.0 * w / ((model == 4) ? 1.0 : P.bD));          // theis/bD
  if (model == 4) {
    // water-table term (:69-92) on top = K0
    const cplx xi = cdivf(cscalef(eta, P.alphaD), p);
    const cplx bex = cscalef(cmulf(eta, xi), P.beta);
    const double MAXEXP = 12.014551129705717;          // constants.f90:66
    cplx dcp, dcm;
    if (eta.re < MAXEXP) {
      const cplx D = caddf(cmulf(mk(1.0 + bex.re, bex.im), E1.ch), cmulf(xi, E1.sh));
      dcp = cscalef(cdivf(K0, D), 0.5);
      dcm = dcp;
    } else {
      const cplx D = mk(1.0 + bex.re + xi.re, bex.im + xi.im);
      dcp = cmulf(cdivf(K0, D), E1.em);
      dcm = zero;
    }
#pragma unroll
    for (int L = 0; L < 3; ++L) sink.set(L, K0, csubf(zero, dcp), csubf(zero, dcm));
    return true;
  }
  // 1/sinh(eta); for Re(eta) > 20, sinh = e^eta (1 - e^-2eta)/2 with e^-2eta < 2^-57
  const cplx ish = (eta.re > 20.0) ? cscalef(E1.em, 2.0) : crecipf(E1.sh);
  const double h = 0.5 * P.bD, m = 0.5 * (P.dD1 + P.lD1);
  const bool need13 = (lay_mask & 5) || model == 3 || model == 5;
  cplx G3 = zero, sK = zero;
  cbundle Em;
  Em.ep = Em.em = Em.ch = Em.sh = zero;
  if (need13) {
    const cbundle Eh = cexp_bundle(eta.re * h, eta.im * h);
    Em = cexp_bundle(eta.re * m, eta.im * m);
    sK = cmulf(K0, cmulf(Eh.sh, ish));   // K0 sinh(eta h)/sinh(eta)
    // above the screen: 2 sK cosh(eta m) cosh(eta(1-z))
    G3 = cmulf(sK, Em.ch);               // (half of it: the 2 cancels the 1/2 of cosh)
  }
  cplx dcp = zero, dcm = zero;
  if (model >= 3) {
    // water-table term, :69-92, on top = udp at zD=1 (layer 3) = 2 G3 (cosh(0) = 1)
    const cplx top = cscalef(G3, 2.0);
    cplx xi = cdivf(cscalef(eta, P.alphaD), p);
    if (model == 3) xi = cdivf(cscalef(xi, (double)P.moench_M), aux);
    const cplx bex = cscalef(cmulf(eta, xi), P.beta);  // beta*eta*xi
    const double MAXEXP = 12.014551129705717;          // constants.f90:66
    if (eta.re < MAXEXP) {
      const cplx D = caddf(cmulf(mk(1.0 + bex.re, bex.im), E1.ch), cmulf(xi, E1.sh));
      dcp = cscalef(cdivf(top, D), 0.5);
      dcm = dcp;
    } else {
      const cplx D = mk(1.0 + bex.re + xi.re, bex.im + xi.im);
      dcp = cmulf(cdivf(top, D), E1.em);
      dcm = zero;
    }
  }
  const bool wt = model >= 3;
  if (need13) {
    sink.set(2, zero, wt ? csubf(cmulf(G3, E1.em), dcp) : cmulf(G3, E1.em),
             wt ? csubf(cmulf(G3, E1.ep), dcm) : cmulf(G3, E1.ep));
    // below the screen: 2 sK cosh(eta(1-m)) cosh(eta z)
    const cplx c1m = cscalef(caddf(cmulf(E1.ep, Em.em), cmulf(E1.em, Em.ep)), 0.5);
    const cplx G1 = cmulf(sK, c1m);
    sink.set(0, zero, wt ? csubf(G1, dcp) : G1, wt ? csubf(G1, dcm) : G1);
  }
  if (lay_mask & 2) {
    // beside the screen: K0 (1 - A cosh(eta z) - B cosh(eta(1-z)))
    const cbundle Ed = cexp_bundle(eta.re * P.dD, eta.im * P.dD);
    const cbundle El = cexp_bundle(eta.re * P.lD1, eta.im * P.lD1);
    const cplx A = cmulf(Ed.sh, ish), B = cmulf(El.sh, ish);
    const cplx cp1 = cmulf(K0, cscalef(caddf(A, cmulf(B, E1.em)), -0.5));
    const cplx cm1 = cmulf(K0, cscalef(caddf(A, cmulf(B, E1.ep)), -0.5));
    sink.set(1, K0, wt ? csubf(cp1, dcp) : cp1, wt ? csubf(cm1, dcm) : cm1);
  }
  return true;
}

template <int MODEL>
__host__ __device__ __forceinline__ bool ap_terms_fast_t(const DevParams &P, cplx p, cplx aux, cplx aux2,
                                                double a2, double w, int lay_mask, double eta_max,
                                                cplx *eta_out, Coef *co /* [3], indexed by layer-1 */) {
  const CoefRegSink sink{co};
  return ap_terms_fast_s<MODEL>(P, p, aux, aux2, a2, w, lay_mask, eta_max, eta_out, sink);
}

__host__ __device__ __forceinline__ bool ap_terms_fast(const DevParams &P, cplx p, cplx aux, cplx aux2,
                                              double a2, double w, int lay_mask, double eta_max,
                                              cplx *eta_out, Coef *co) {
  return ap_terms_fast_t<-1>(P, p, aux, aux2, a2, w, lay_mask, eta_max, eta_out, co);
}

// f(z) for one z given the per-(a,p) terms: one exp_pm + one sincos + 12 FMA
__host__ __device__ __forceinline__ cplx eval_z_fast(cplx eta, const Coef &c, double z) {
  double ep, em, cc, ss, s, cs;
  int kk;
  exp_pm_core(eta.re * z, &ep, &em, &cc, &ss, &kk);
  sincos_q(eta.im * z, &s, &cs);
  const double pc = ep * cs, ps = ep * s, mc = em * cs, ms = em * s;
  // k0 + cp*(pc + i ps) + cm*(mc - i ms)
  double fr = fma(c.cp.re, pc, c.k0.re);
  fr = fma(-c.cp.im, ps, fr);
  fr = fma(c.cm.re, mc, fr);
  fr = fma(c.cm.im, ms, fr);
  double fi = fma(c.cp.re, ps, c.k0.im);
  fi = fma(c.cp.im, pc, fi);
  fi = fma(c.cm.im, mc, fi);
  fi = fma(-c.cm.re, ms, fi);
  return mk(fr, fi);
}

}  // namespace unc
