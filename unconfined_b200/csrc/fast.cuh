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
// Valid while Re(eta) <= UNC_FAST_ETA_MAX, where neither formulation can overflow;
// beyond that the caller uses the literal path, which overflows exactly where the
// reference does.
#pragma once
#include "cmath.cuh"

namespace unc {

#define UNC_FAST_ETA_MAX 345.0  /* every intermediate of both formulations <= e^(2*345) < DBL_MAX */

// exp(+-x), cosh x, sinh x from one range reduction: x = k ln2 + r, |r| <= ln2/2,
// cosh r / sinh r by even/odd Taylor polynomials, exp(+-r) = cosh r +- sinh r.
struct rexp {
  double ep, em, ch, sh;
};
__device__ __forceinline__ rexp exp_pm(double x) {
  const double L2E = 1.4426950408889634074, LN2H = 6.93147180369123816490e-01,
               LN2L = 1.90821492927058770002e-10;
  const double kf = rint(x * L2E);
  double r = fma(-kf, LN2H, x);
  r = fma(-kf, LN2L, r);
  const double r2 = r * r;
  double c = 2.08767569878680989792e-09;            // 1/12!
  c = fma(c, r2, 2.75573192239858906526e-07);       // 1/10!
  c = fma(c, r2, 2.48015873015873015873e-05);       // 1/8!
  c = fma(c, r2, 1.38888888888888888889e-03);       // 1/6!
  c = fma(c, r2, 4.16666666666666666667e-02);       // 1/4!
  c = fma(c, r2, 0.5);
  c = fma(c, r2, 1.0);
  double s = 1.60590438368216145994e-10;            // 1/13!
  s = fma(s, r2, 2.50521083854417187751e-08);       // 1/11!
  s = fma(s, r2, 2.75573192239858906526e-06);       // 1/9!
  s = fma(s, r2, 1.98412698412698412698e-04);       // 1/7!
  s = fma(s, r2, 8.33333333333333333333e-03);       // 1/5!
  s = fma(s, r2, 1.66666666666666666667e-01);       // 1/3!
  s = fma(s * r2, r, r);
  const int k = (int)kf;
  // |k| <= 1020 on this path (|x| <= ~700): 2^k and 2^-k are normal doubles
  const double sp = __longlong_as_double((long long)(1023 + k) << 52);
  const double sm = __longlong_as_double((long long)(1023 - k) << 52);
  rexp o;
  o.ep = (c + s) * sp;
  o.em = (c - s) * sm;
  // k == 0: the polynomials ARE cosh/sinh (no cancellation for small x)
  o.ch = (k == 0) ? c : 0.5 * (o.ep + o.em);
  o.sh = (k == 0) ? s : 0.5 * (o.ep - o.em);
  return o;
}

// plain complex helpers for finite operands (no real->complex promotion)
__device__ __forceinline__ cplx cmulf(cplx a, cplx b) {
  return mk(fma(a.re, b.re, -(a.im * b.im)), fma(a.re, b.im, a.im * b.re));
}
__device__ __forceinline__ cplx crecipf(cplx b) {
  const double d = 1.0 / fma(b.re, b.re, b.im * b.im);
  return mk(b.re * d, -(b.im * d));
}
__device__ __forceinline__ cplx cdivf(cplx a, cplx b) { return cmulf(a, crecipf(b)); }
__device__ __forceinline__ cplx caddf(cplx a, cplx b) { return mk(a.re + b.re, a.im + b.im); }
__device__ __forceinline__ cplx csubf(cplx a, cplx b) { return mk(a.re - b.re, a.im - b.im); }
__device__ __forceinline__ cplx cscalef(cplx a, double x) { return mk(a.re * x, a.im * x); }

// complex exp(+-w), cosh w, sinh w of w = eta*c
struct cbundle {
  cplx ep, em, ch, sh;
};
__device__ __forceinline__ cbundle cexp_bundle(double wr, double wi) {
  const rexp e = exp_pm(wr);
  double s, c;
  sincos(wi, &s, &c);
  cbundle b;
  b.ep = mk(e.ep * c, e.ep * s);
  b.em = mk(e.em * c, -(e.em * s));
  b.ch = mk(e.ch * c, e.sh * s);
  b.sh = mk(e.sh * c, e.ch * s);
  return b;
}

// sqrt of z with Re z > 0 (eta = sqrt((p+a^2)/kappa)), glibc's formula for that branch
__device__ __forceinline__ cplx csqrt_pos(cplx z) {
  if (z.im == 0.0) return mk(sqrt(z.re), 0.0);
  const double d = sqrt(fma(z.re, z.re, z.im * z.im));
  const double r = sqrt(0.5 * (d + z.re));
  return mk(r, 0.5 * (z.im / r));
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
__device__ __forceinline__ bool ap_terms_fast(const DevParams &P, cplx p, cplx aux, cplx aux2,
                                              double a2, double w, int lay_mask, cplx *eta_out,
                                              Coef *co /* [3], indexed by layer-1 */) {
  const int model = P.model;
  const cplx pa = mk(p.re + a2, p.im);
  const cplx zero = mk(0.0, 0.0);
  if (model == 0) {
    const cplx th = cscalef(crecipf(pa), 2.0 * w);  // theis :122-131
    *eta_out = zero;
#pragma unroll
    for (int L = 0; L < 3; ++L) { co[L].k0 = th; co[L].cp = zero; co[L].cm = zero; }
    return true;
  }
  const cplx eta = csqrt_pos(cscalef(pa, 1.0 / P.kappa));
  *eta_out = eta;
  if (!(eta.re <= UNC_FAST_ETA_MAX)) return false;
  const cbundle E1 = cexp_bundle(eta.re, eta.im);
  cplx K0;  // common prefactor of the layer functions, weight folded in
  if (model == 2) K0 = cscalef(cdivf(aux, cmulf(pa, aux2)), w / P.bD);            // uDf/bD :265-266,299
  else K0 = cscalef(crecipf(pa), 2.0 * w / ((model == 4) ? 1.0 : P.bD));          // theis/bD
  cplx top = zero;  // udp at zD=1 (layer 3) for models 3 and 5
  if (model == 4) {
#pragma unroll
    for (int L = 0; L < 3; ++L) { co[L].k0 = K0; co[L].cp = zero; co[L].cm = zero; }
    top = K0;
  } else {
    const cplx ish = crecipf(E1.sh);  // 1/sinh(eta)
    const double h = 0.5 * P.bD, m = 0.5 * (P.dD1 + P.lD1);
    const bool need13 = (lay_mask & 5) || model == 3 || model == 5;
    if (need13) {
      const cbundle Eh = cexp_bundle(eta.re * h, eta.im * h);
      const cbundle Em = cexp_bundle(eta.re * m, eta.im * m);
      const cplx sK = cmulf(K0, cmulf(Eh.sh, ish));  // K0 sinh(eta h)/sinh(eta)
      // above the screen: 2 sK cosh(eta m) cosh(eta(1-z))
      const cplx G3 = cmulf(sK, Em.ch);              // (half of it: the 2 cancels the 1/2 of cosh)
      co[2].k0 = zero;
      co[2].cp = cmulf(G3, E1.em);
      co[2].cm = cmulf(G3, E1.ep);
      top = cscalef(G3, 2.0);                        // cosh(0) = 1
      // below the screen: 2 sK cosh(eta(1-m)) cosh(eta z)
      const cplx c1m = cscalef(caddf(cmulf(E1.ep, Em.em), cmulf(E1.em, Em.ep)), 0.5);
      const cplx G1 = cmulf(sK, c1m);
      co[0].k0 = zero;
      co[0].cp = G1;
      co[0].cm = G1;
    }
    if (lay_mask & 2) {
      // beside the screen: K0 (1 - A cosh(eta z) - B cosh(eta(1-z)))
      const cbundle Ed = cexp_bundle(eta.re * P.dD, eta.im * P.dD);
      const cbundle El = cexp_bundle(eta.re * P.lD1, eta.im * P.lD1);
      const cplx A = cmulf(Ed.sh, ish), B = cmulf(El.sh, ish);
      co[1].k0 = K0;
      co[1].cp = cmulf(K0, cscalef(caddf(A, cmulf(B, E1.em)), -0.5));
      co[1].cm = cmulf(K0, cscalef(caddf(A, cmulf(B, E1.ep)), -0.5));
    }
  }
  if (model >= 3) {
    // water-table term, :69-92
    cplx xi = cdivf(cscalef(eta, P.alphaD), p);
    if (model == 3) xi = cdivf(cscalef(xi, (double)P.moench_M), aux);
    const cplx bex = cscalef(cmulf(eta, xi), P.beta);  // beta*eta*xi
    const double MAXEXP = 12.014551129705717;          // constants.f90:66
    cplx dcp, dcm;
    if (eta.re < MAXEXP) {
      const cplx D = caddf(cmulf(mk(1.0 + bex.re, bex.im), E1.ch), cmulf(xi, E1.sh));
      dcp = cscalef(cdivf(top, D), 0.5);
      dcm = dcp;
    } else {
      const cplx D = mk(1.0 + bex.re + xi.re, bex.im + xi.im);
      dcp = cmulf(cdivf(top, D), E1.em);
      dcm = zero;
    }
#pragma unroll
    for (int L = 0; L < 3; ++L) {
      co[L].cp = csubf(co[L].cp, dcp);
      co[L].cm = csubf(co[L].cm, dcm);
    }
  }
  return true;
}

// f(z) for one z given the per-(a,p) terms: one exp_pm + one sincos + 12 FMA
__device__ __forceinline__ cplx eval_z_fast(cplx eta, const Coef &c, double z) {
  const rexp e = exp_pm(eta.re * z);
  double s, cs;
  sincos(eta.im * z, &s, &cs);
  const double pc = e.ep * cs, ps = e.ep * s, mc = e.em * cs, ms = e.em * s;
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
