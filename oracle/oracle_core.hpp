// TEST INFRASTRUCTURE ONLY -- CPU oracle (see oracle_math.hpp header note; PARITY UNPINNED).
//
// Restatement of the reference's Laplace-Hankel hot path, templated on the real
// type (double = the reference's EP=DP build, constants.f90:35; long double is used
// only to measure the double build's own rounding floor).  Every function cites the
// reference lines it follows (paths relative to /root/reference).
#pragma once
#include <cstdint>
#include <vector>
#include <cmath>
#include <limits>
#include <algorithm>
#include "oracle_math.hpp"

namespace orc {

// Same field order/types as include/unconfined_b200.h:unc_params (declared again
// here so that the oracle shares no source with the product).
struct params {
  int32_t model;
  int32_t M;
  double alpha;
  double tol;
  double tee_mult;
  int32_t time_type;
  int32_t n_time_par;
  const double *time_par;
  int32_t ts_k;
  int32_t ts_R;
  int32_t gl_nacc;
  int32_t gl_ord;
  int32_t n_j0z;
  int32_t moench_M;
  const double *j0z;
  const double *moench_gamma;
  double kappa, alphaD, beta;
  double lD, dD, bD, rDw;
  double l, d, Ss, rDwobs, sF;
  int32_t mn_type, mn_reserved;
  double mn_ak, mn_psia, mn_psik, mn_b, mn_Sy;
};

template <class T> struct consts {
  static T pi() { return T(4) * std::atan(T(1)); }        // constants.f90:51,54
  static T piov2() { return T(2) * std::atan(T(1)); }     // constants.f90:57
  static T eps() { return std::numeric_limits<T>::epsilon(); }
  static T maxexp() { return -std::log(eps()) / T(3); }   // constants.f90:66
};

// ---------------------------------------------------------------------------
// integration.f90:31-67  tanh_sinh_setup (weights for level k; abscissae if asked)
template <class T>
void tanh_sinh_setup(int k, T s, std::vector<T> &w, std::vector<T> *a) {
  const int N = (1 << k) - 1;
  const int r = (N - 1) / 2;
  const T h = T(4) / T(1 << k);
  std::vector<T> u1(N), u2(N);
  for (int i = -r; i <= r; ++i) {
    u1[i + r] = consts<T>::piov2() * std::cosh(h * T(i));
    u2[i + r] = consts<T>::piov2() * std::sinh(h * T(i));
  }
  w.resize(N);
  T sum = 0;
  for (int i = 0; i < N; ++i) {
    T c = std::cosh(u2[i]);
    w[i] = u1[i] / (c * c);
  }
  for (int i = 0; i < N; ++i) sum += w[i];
  for (int i = 0; i < N; ++i) w[i] = T(2) * w[i] / sum;
  if (a) {
    a->resize(N);
    for (int i = 0; i < N; ++i) (*a)[i] = (std::tanh(u2[i]) + T(1)) * s / T(2);
  }
}

// integration.f90:70-120  gauss_lobatto_setup (interior nodes/weights only)
template <class T>
void gauss_lobatto_setup(int ord, std::vector<T> &gx, std::vector<T> &gw) {
  const int N = ord - 1, N1 = N + 1;
  std::vector<T> x(ord), xold(ord, T(2)), w(ord);
  std::vector<std::vector<T>> P(ord + 1, std::vector<T>(ord, T(0)));  // P[k][i], k=1..ord
  for (int i = 0; i <= N; ++i) x[i] = std::cos(consts<T>::pi() * T(i) / T(N));
  for (;;) {
    T mx = 0;
    for (int i = 0; i < ord; ++i) mx = std::max(mx, std::fabs(x[i] - xold[i]));
    if (!(mx > consts<T>::eps())) break;  // spacing(1.0_EP)
    xold = x;
    for (int i = 0; i < ord; ++i) { P[1][i] = T(1); P[2][i] = x[i]; }
    for (int k = 2; k <= N; ++k)
      for (int i = 0; i < ord; ++i)
        P[k + 1][i] = (T(2 * k - 1) * x[i] * P[k][i] - T(k - 1) * P[k - 1][i]) / T(k);
    for (int i = 0; i < ord; ++i)
      x[i] = xold[i] - (x[i] * P[N1][i] - P[N][i]) / (T(N1) * P[N1][i]);
  }
  for (int i = 0; i < ord; ++i) w[i] = T(2) / (T(N * N1) * (P[N1][i] * P[N1][i]));
  gx.assign(x.begin() + 1, x.begin() + ord - 1);
  gw.assign(w.begin() + 1, w.begin() + ord - 1);
}

// integration.f90:125-189  wynn_epsilon.  Returns the accelerated sum; *info (optional):
// 0 full table, 1 truncated, 2 sentinel (<4 finite terms), 3 "epsilon cancel" early exit.
template <class T>
cx<T> wynn_epsilon(const cx<T> *series, int n, int *info = nullptr) {
  const int MINTERMS = 4;
  int ns = n;
  if (info) *info = 0;
  // eps(i,j): i=1..n, j=-1..n-1  -> e[(j+1)*(n+1)+i]
  std::vector<cx<T>> e((size_t)(n + 1) * (n + 2));
  auto E = [&](int i, int j) -> cx<T> & { return e[(size_t)(j + 1) * (n + 1) + i]; };
  for (int i = 1; i <= ns; ++i) {
    if (!is_finite(series[i - 1])) {
      ns = i - 1;
      if (ns < MINTERMS) {
        if (info) *info = 2;
        return cx<T>(T(-999999.9f), T(0));  // default-real literal -999999.9 (integration.f90:147)
      }
      if (info) *info = 1;
      break;
    }
    cx<T> s(0, 0);
    for (int k = 0; k < i; ++k) s = s + series[k];  // sum(series(1:i))
    E(i, 0) = s;
  }
  for (int i = 1; i <= n; ++i) E(i, -1) = cx<T>(0, 0);
  for (int j = 0; j <= ns - 2; ++j) {
    for (int m = 1; m <= ns - (j + 1); ++m) {
      cx<T> denom = E(m + 1, j) - E(m, j);
      if (cabs(denom) > consts<T>::eps()) {  // epsilon(abs(denom)): absolute
        E(m, j + 1) = E(m + 1, j - 1) + T(1) / denom;
      } else {
        if (info) *info = 3;
        return E(m + 1, j);
      }
    }
  }
  return (ns % 2 == 0) ? E(2, ns - 2) : E(2, ns - 3);
}

// integration.f90:192-237  extraptozero (polynomial extrapolation to x=0)
template <class T>
cx<T> extraptozero(const T *xin, const cx<T> *yin, int n) {
  int ns = 1;
  for (int i = 2; i <= n; ++i) if (xin[i - 1] < xin[ns - 1]) ns = i;  // minloc (first minimum)
  std::vector<cx<T>> c(yin, yin + n), d(yin, yin + n), den(n);
  cx<T> y = yin[ns - 1];
  ns -= 1;
  for (int m = 1; m <= n - 1; ++m) {
    for (int i = 1; i <= n - m; ++i) den[i - 1] = cx<T>(xin[i - 1] - xin[i + m - 1], T(0));
    for (int i = 1; i <= n - m; ++i) den[i - 1] = (c[i] - d[i - 1]) / den[i - 1];
    for (int i = 1; i <= n - m; ++i) {
      d[i - 1] = xin[i + m - 1] * den[i - 1];
      c[i - 1] = xin[i - 1] * den[i - 1];
    }
    cx<T> dy;
    if (2 * ns < n - m) dy = c[ns];
    else { dy = d[ns - 1]; ns -= 1; }
    y = y + dy;
  }
  return y;
}

// invlap.f90:154-172  deHoog_pvalues
template <class T>
void dehoog_pvalues(T tee, const params &prm, std::vector<cx<T>> &p) {
  const int np = 2 * prm.M + 1;
  p.resize(np);
  T sigma = T(prm.alpha) - std::log(T(prm.tol)) / (T(2) * tee);
  for (int i = 0; i < np; ++i) p[i] = cx<T>(sigma, consts<T>::pi() * T(i) / tee);
}

// invlap.f90:46-152  deHoog_invLap (scalar t)
template <class T>
T dehoog_invlap(T t, T tee, const cx<T> *fp, const params &prm) {
  const int M = prm.M, n2 = 2 * M;
  // maxval(abs(fp)) > tiny  (gfortran MAXVAL skips NaNs unless all are NaN)
  T mx = -std::numeric_limits<T>::infinity();
  bool anynum = false;
  for (int i = 0; i <= n2; ++i) {
    T a = cabs(fp[i]);
    if (!std::isnan(a)) { anynum = true; if (a > mx) mx = a; }
  }
  if (!anynum || !(mx > std::numeric_limits<T>::min())) return T(0);
  std::vector<cx<T>> f(fp, fp + n2 + 1);
  for (int i = 0; i <= n2; ++i)
    if (std::isnan(f[i].re) || std::isnan(f[i].im)) f[i] = cx<T>(0, 0);
  // gamma uses the DP module constants (invlap.f90:77): all DP when EP=DP
  T gamma = T(prm.alpha) - std::log(T(prm.tol)) / (T(2) * tee);
  // only column r-1 of e and columns r, r+1 of q are live
  std::vector<cx<T>> eprev(n2 + 1, cx<T>(0, 0)), ecur(n2 + 1), q(n2 + 1), qn(n2 + 1), d(n2 + 1);
  q[0] = f[1] / (f[0] / T(2));
  for (int i = 1; i <= n2 - 1; ++i) q[i] = f[i + 1] / f[i];
  d[0] = f[0] / T(2);
  for (int r = 1; r <= M; ++r) {
    int mx_ = 2 * (M - r);
    for (int i = 0; i <= mx_; ++i) ecur[i] = q[i + 1] - q[i] + eprev[i + 1];
    d[2 * r - 1] = -q[0];
    d[2 * r] = -ecur[0];
    if (r != M) {
      int rq = r + 1;
      mx_ = 2 * (M - rq) + 1;
      for (int i = 0; i <= mx_; ++i) qn[i] = q[i + 1] * ecur[i + 1] / ecur[i];
      q.swap(qn);
    }
    eprev.swap(ecur);
  }
  // z = exp(i*pi*t/tee): (cmplx(0,1)*PI)*t/tee  (invlap.f90:110)
  const T PI = consts<T>::pi();
  cx<T> z = cexp(((cx<T>(0, 1) * PI) * t) / tee);
  cx<T> Am2(0, 0), Am1 = d[0], Bm2(1, 0), Bm1(1, 0);
  for (int n = 1; n <= n2 - 1; ++n) {
    cx<T> An = Am1 + d[n] * Am2 * z;
    cx<T> Bn = Bm1 + d[n] * Bm2 * z;
    Am2 = Am1; Am1 = An; Bm2 = Bm1; Bm1 = Bn;
  }
  cx<T> brem = (T(1) + (d[n2 - 1] - d[n2]) * z) / T(2);
  cx<T> rem = (-brem) * (T(1) - csqrt(T(1) + d[n2] * z / (brem * brem)));
  cx<T> A2M = Am1 + rem * Am2;
  cx<T> B2M = Bm1 + rem * Bm2;
  return std::exp(gamma * t) / tee * (A2M / B2M).re;
}

// time.f90:34-124  lapTime
template <class T>
void lap_time(const params &prm, const std::vector<cx<T>> &p, std::vector<cx<T>> &mult) {
  const int np = (int)p.size();
  mult.resize(np);
  auto par = [&](int i) { return T(prm.time_par[i - 1]); };  // 1-based
  const int tt = prm.time_type;
  for (int k = 0; k < np; ++k) {
    cx<T> pk = p[k];
    cx<T> m;
    if (tt == 1) {
      m = cexp((-par(1)) * pk) / pk;
    } else if (tt == 2) {
      m = cexp((-par(1)) * pk) / pk - cexp((-par(2)) * pk) / pk;
    } else if (tt == 3) {
      m = cexp((-par(1)) * pk);
    } else if (tt == 4) {
      m = T(1) / (pk - pk * cexp((-par(1)) * pk)) * (T(1) - cexp((-par(2)) * pk)) / pk;
    } else if (tt == 5) {
      m = cexp((-par(2)) * pk) / (pk + pk * cexp((-par(1)) * pk));
    } else if (tt == 6) {
      m = cexp((-par(2)) * pk) * pk / (pk * pk + par(1) * par(1));
    } else if (tt == 7) {
      // as written the numerator is exp(x)-exp(x) == 0 (time.f90:72-74)
      cx<T> e1 = cexp(par(1) * pk);
      m = cexp((-par(2)) * pk) / (pk * pk) * (e1 - e1) / (e1 + e1);
    } else if (tt == 8) {
      cx<T> eh = cexp((-par(1)) * pk / T(2));
      m = cexp((-par(2)) * pk) * (T(1) - eh) / ((T(1) + eh) * pk);
    } else if (tt < 0 && tt >= -100) {
      const int n = -tt;
      T tf = par(n + 1);
      std::vector<T> Q(n + 1);
      Q[0] = 0;
      for (int i = 1; i <= n; ++i) Q[i] = par(n + 1 + i);
      cx<T> s(0, 0);
      T sq = 0;
      for (int i = 1; i <= n; ++i) {
        s = s + (Q[i] - Q[i - 1]) * cexp((-par(i)) * pk);
        sq += Q[i] - Q[i - 1];
      }
      m = (s - sq * cexp((-tf) * pk)) / pk;
    } else if (tt <= -101) {
      const int n = -tt - 100;
      T tf = par(n + 1);
      std::vector<T> ti(n + 2), y(n + 2), W(n + 2);
      for (int i = 1; i <= n; ++i) { ti[i] = par(i); y[i] = par(n + 1 + i); }
      // NB time.f90:110 reads y(2:n+1) although y has n entries; entry n+1 is
      // out of bounds in the reference.  The oracle uses 0 there (documented).
      y[n + 1] = 0;
      W[0] = 0;
      for (int i = 1; i <= n; ++i) {
        T den = ((i < n) ? ti[i + 1] : tf) - ti[i];
        W[i] = (y[i + 1] - y[i]) / den;
      }
      cx<T> s(0, 0);
      T sw = 0;
      for (int i = 1; i <= n; ++i) {
        s = s + (W[i] - W[i - 1]) * cexp((-ti[i]) * pk);
        sw += W[i] - W[i - 1];
      }
      m = (s - sw * cexp((-tf) * pk)) / (pk * pk);
    } else {
      m = cx<T>(std::numeric_limits<T>::quiet_NaN(), std::numeric_limits<T>::quiet_NaN());
    }
    mult[k] = m;
  }
}

// ---------------------------------------------------------------------------
// cbessel.f90:877-1146 (cbesk) -> 5036-5495 (cbknu), specialised to fnu=0, kode=1, n=2,
// Re z >= 0.  Returns ierr as cbesk does; nz_out = number of underflowed members.
// Includes the exp(-z) underflow branch for Re z > alim (cbessel.f90:5215 -> label 200 :5482,
// scaled Miller recurrence, label 190 :5458-5466 -> ckscl :5499-5611 -> cuchk :5895-5927).
inline int cbesk01(cx<double> z, cx<double> K[2], int *nz_out) {
  typedef double R;
  typedef cx<double> C;
  *nz_out = 0;
  const R tiny_ = std::numeric_limits<R>::min();
  const R spacing0 = tiny_;  // spacing(0.0_dp) == tiny(0.0_dp)
  R xx = z.re, yy = z.im;
  if (std::fabs(yy) * 2.0 < spacing0 && std::fabs(xx) * 2.0 < spacing0) return 1;
  const R tol = std::max(std::numeric_limits<R>::epsilon(), 1.0e-18);
  const int k1m = -1021, k2m = 1024;  // MINEXPONENT, MAXEXPONENT
  const R r1m5 = std::log10(2.0);
  const int kk = std::min(std::abs(k1m), std::abs(k2m));
  const R elim = 2.303 * (kk * r1m5 - 3.0);
  const int k1 = 53 - 1;
  R aa = r1m5 * k1;
  aa = aa * 2.303;
  const R alim = elim + std::max(-aa, -41.45);
  const R az = cabs(z);
  const R fn = 1.0;
  aa = 0.5 / tol;
  R bb = 2147483647.0 * 0.5;
  aa = std::min(aa, bb);
  int ierr = 0;
  if (!(az <= aa)) return 4;
  if (!(fn <= aa)) return 4;
  aa = std::sqrt(aa);
  if (az > aa) ierr = 3;
  const R ufl = tiny_ * 1.0e+3f;  // TINY*1.0E+3 (default-real literal, exact)
  if (!(az >= ufl)) return 2;
  if (!(xx >= 0.0)) return 7;  // left half plane (cacon) not restated: unreachable here
  // ---- cbknu ----
  const R cc1 = 5.77215664901532861e-01;
  // module constants, cbessel.f90:20-25 (computed from atan as in the source)
  const R pi = 4.0 * std::atan(1.0), rthpi = std::sqrt(8.0 * std::atan(1.0)) / 2.0,
          spi = 3.0 / (2.0 * std::atan(1.0)), hpi = 2.0 * std::atan(1.0);
  const R fpi = 1.89769999331517738, tth = 6.66666666666666666e-01;  // cbessel.f90:5066
  const R caz = az;
  const C rz = C(2.0, 0.0) / z;
  const R dnu = 0.0, dnu2 = 0.0;
  C s1, s2;
  int kflag, iflag = 0;
  const R cssv[4] = {0, 1.0 / tol, 1.0, tol}, csrv[4] = {0, tol, 1.0, 1.0 / tol};
  if (caz <= 2.0) {
    // series, cbessel.f90:5098-5200
    C smu = clog(rz);
    C fmu = smu * dnu;
    // cshch(fmu): csh = sinh, cch = cosh of fmu (=0)  cbessel.f90:4896-4925
    C cch(std::cosh(fmu.re) * std::cos(fmu.im), std::sinh(fmu.re) * std::sin(fmu.im));
    R a2 = 1.0 + dnu;
    R t2 = std::exp(-lm<double>::lgamma(a2));
    R fc = 1.0;
    R t1 = 1.0 / (t2 * fc);
    R s = cc1;  // DO k=2,8: ak*=dnu2 (=0); tm=0 -> exits at k=2
    R g1 = -s;
    R g2 = 0.5 * (t1 + t2) * fc;
    g1 = g1 * fc;
    C f = g1 * cch + smu * g2;
    C pt = cexp(fmu);
    C p = C(0.5 / t2, 0.0) * pt;
    C q = C(0.5 / t1, 0.0) / pt;
    s1 = f;
    s2 = p;
    R ak = 1.0, a1 = 1.0, bk = 1.0 - dnu2;
    C ck(1.0, 0.0);
    if (caz >= tol) {
      C cz = z * z * 0.25;
      t1 = 0.25 * caz * caz;
      do {
        f = (f * ak + p + q) / bk;
        p = p / (ak - dnu);
        q = q / (ak + dnu);
        R rk = 1.0 / ak;
        ck = ck * cz * rk;
        s1 = s1 + ck * f;
        s2 = s2 + ck * (p - f * ak);
        a1 = a1 * t1 * rk;
        bk = bk + ak + ak + 1.0;
        ak = ak + 1.0;
      } while (a1 > tol);
    }
    kflag = 2;
    bk = smu.re;
    a1 = 0.0 + 1.0;
    ak = a1 * std::fabs(bk);
    if (ak > alim) kflag = 3;
    C p2 = s2 * C(cssv[kflag], 0.0);
    s2 = p2 * rz;
    s1 = s1 * C(cssv[kflag], 0.0);
  } else {
    // Miller, cbessel.f90:5209-5327
    C coef = C(rthpi, 0.0) / csqrt(z);
    kflag = 2;
    iflag = (xx > alim) ? 1 : 0;   // cbessel.f90:5215 -> 200: koded = 2, values stay scaled by exp(z)
    if (!iflag) {
      R a1 = std::exp(-xx) * cssv[kflag];
      C pt = a1 * C(std::cos(yy), -std::sin(yy));
      coef = coef * pt;
    }
    R ak = std::fabs(std::cos(pi * dnu));
    R fhs = std::fabs(0.25 - dnu2);
    R t1 = (53 - 1) * std::log10(2.0) * 3.321928094;
    t1 = std::max(t1, 12.0);
    t1 = std::min(t1, 60.0);
    R t2 = tth * t1 - 6.0;
    if (std::fabs(xx) * 2.0 < spacing0) t1 = hpi;
    else t1 = std::fabs(std::atan(yy / xx));
    R fk;
    if (t2 <= caz) {
      R etest = ak / (pi * caz * tol);
      fk = 1.0;
      if (!(etest < 1.0)) {
        R fks = 2.0, rk = caz + caz + 2.0, a1 = 0.0, a2 = 1.0;
        bool found = false;
        for (int i = 1; i <= 30; ++i) {
          ak = fhs / fks;
          R bk = rk / (fk + 1.0);
          R tm = a2;
          a2 = bk * a2 - ak * a1;
          a1 = tm;
          rk = rk + 2.0;
          fks = fks + fk + fk + 2.0;
          fhs = fhs + fk + fk;
          fk = fk + 1.0;
          tm = std::fabs(a2) * fk;
          if (etest < tm) { found = true; break; }
        }
        if (!found) return 5;  // nz=-2 -> cbesk ierr=5
        fk = fk + spi * t1 * std::sqrt(t2 / caz);
        fhs = std::fabs(0.25 - dnu2);
      }
    } else {
      R a2 = std::sqrt(caz);
      ak = fpi * ak / (tol * std::sqrt(a2));
      R aa2 = 3.0 * t1 / (1.0 + caz);
      R bb2 = 14.7 * t1 / (28.0 + caz);
      ak = (std::log(ak) + caz * std::cos(aa2) / (1.0 + 0.008 * caz)) / std::cos(bb2);
      fk = 0.12125 * ak * ak / caz + 1.5;
    }
    int k = (int)fk;
    fk = k;
    R fks = fk * fk;
    C p1(0, 0), p2(tol, 0), cs = p2;
    for (int i = 1; i <= k; ++i) {
      R a1 = fks - fk;
      R a2 = (fks + fk) / (a1 + fhs);
      R rk = 2.0 / (fk + 1.0);
      R tt1 = (fk + xx) * rk;
      R tt2 = yy * rk;
      C pt = p2;
      p2 = (p2 * C(tt1, tt2) - p1) * a2;
      p1 = pt;
      cs = cs + p2;
      fks = a1 - fk + 1.0;
      fk = fk - 1.0;
    }
    R tm = cabs(cs);
    C pt = C(1.0 / tm, 0.0);
    s1 = pt * p2;
    cs = conj(cs) * pt;
    s1 = coef * s1 * cs;
    tm = cabs(p2);
    pt = C(1.0 / tm, 0.0);
    p1 = pt * p1;
    p2 = conj(p2) * pt;
    pt = p1 * p2;
    s2 = s1 * (C(1.0, 0.0) + (C(dnu + 0.5, 0.0) - pt) / z);
  }
  if (iflag) {
    // label 100 (inu = 0) -> 190: y = (s1, s2); ckscl with zd = z, n = 2, ascle = bry(1)
    const R ascle = 1.0e+3f * tiny_ / tol;
    C y[2] = {s1, s2};
    int nz = 0, ic = 0;
    for (int i = 1; i <= 2; ++i) {
      const C sv = y[i - 1];
      const R as = cabs(sv);
      const R acs = -xx + std::log(as);
      nz += 1;
      y[i - 1] = C(0.0, 0.0);
      if (acs >= -elim) {
        C cs = (-z) + clog(sv);
        const R aa2 = std::exp(cs.re) / tol;
        cs = aa2 * C(std::cos(cs.im), std::sin(cs.im));
        // cuchk
        int nw = 0;
        const R yr = std::fabs(cs.re), yi = std::fabs(cs.im);
        R st = std::min(yr, yi);
        if (!(st > ascle)) {
          const R ss = std::max(yr, yi);
          st = st / tol;
          if (ss < st) nw = 1;
        }
        if (nw == 0) { y[i - 1] = cs; nz -= 1; ic = i; }
      }
    }
    if (ic <= 1) { y[0] = C(0.0, 0.0); nz = 2; }
    // back in cbknu (:5467-5476): the surviving members are unscaled by csr(1) = tol
    K[0] = C(0.0, 0.0); K[1] = C(0.0, 0.0);
    if (nz == 0) { K[0] = y[0] * C(csrv[1], 0.0); K[1] = y[1] * C(csrv[1], 0.0); }
    else if (nz == 1) { K[1] = y[1] * C(csrv[1], 0.0); }
    *nz_out = nz;
    return ierr;
  }
  // label 100 -> 130 (inu=0, n=2)
  K[0] = s1 * C(csrv[kflag], 0.0);
  K[1] = s2 * C(csrv[kflag], 0.0);
  return ierr;
}

// ---------------------------------------------------------------------------
// laplace_hankel_solutions.f90:122-131  theis
template <class T> inline cx<T> theis(T a, cx<T> p) { return T(2) / (p + a * a); }

// laplace_hankel_solutions.f90:133-202  hantush; zD[0..nz-1], zLay per z; out udp[p*nz + z]
template <class T>
void hantush(T a, const T *zD, const int *zLay, int nz, const std::vector<cx<T>> &p,
             const params &prm, std::vector<cx<T>> &udp) {
  const int np = (int)p.size();
  const T dD1 = T(1) - T(prm.dD), lD1 = T(1) - T(prm.lD);
  udp.resize((size_t)np * nz);
  for (int k = 0; k < np; ++k) {
    cx<T> eta = csqrt((p[k] + a * a) / T(prm.kappa));
    cx<T> ff1 = csinh(eta * T(prm.dD));
    cx<T> ff2 = csinh(eta * lD1);
    cx<T> sh = csinh(eta);
    cx<T> g3 = cexp((-eta) * lD1) - (ff1 + cexp(-eta) * ff2) / sh;
    cx<T> th = theis(a, p[k]);
    for (int z = 0; z < nz; ++z) {
      cx<T> g1 = ccosh(eta * (dD1 - zD[z]));
      cx<T> g2 = (ff1 * ccosh(eta * zD[z]) + ff2 * ccosh(eta * (T(1) - zD[z]))) / sh;
      cx<T> u;
      if (zLay[z] == 1) u = g3 * ccosh(eta * zD[z]);
      else if (zLay[z] == 2) u = T(1) - g2;
      else u = g1 - g2;
      udp[(size_t)k * nz + z] = u * th / T(prm.bD);
    }
  }
}

// laplace_hankel_solutions.f90:204-301  hantushstorage (layer-3 points: evident intent,
// same layer functions as hantush; the reference reads unset ff there)
inline int hantushstorage(double a, const double *zD, const int *zLay, int nz,
                          const std::vector<cx<double>> &p, const params &prm,
                          std::vector<cx<double>> &u) {
  typedef double T;
  const int np = (int)p.size();
  const T PI = consts<T>::pi();
  const T dD1 = 1.0 - prm.dD, lD1 = 1.0 - prm.lD;
  const T CDw = prm.rDw * prm.rDw / (2.0 * (prm.l - prm.d) * prm.Ss);
  const T tDb = PI * prm.rDwobs * prm.rDwobs / (prm.sF * prm.Ss);
  u.resize((size_t)np * nz);
  int worst = 0;
  for (int k = 0; k < np; ++k) {
    cx<T> xi = prm.rDw * csqrt(p[k]);
    cx<T> eta = csqrt((p[k] + a * a) / prm.kappa);
    cx<T> K[2];
    int nzero;
    int ierr = cbesk01(xi, K, &nzero);
    if (ierr > 0 && ierr != 3) worst = ierr;
    cx<T> A0 = 2.0 / (p[k] * CDw * K[0] + xi * K[1]);
    cx<T> uDf = A0 / ((p[k] + a * a) * (p[k] * tDb + 1.0));
    cx<T> ff1 = csinh(eta * prm.dD);
    cx<T> ff2 = csinh(eta * lD1);
    cx<T> sh = csinh(eta);
    cx<T> ff3 = cexp((-eta) * lD1) - (ff1 + cexp(-eta) * ff2) / sh;
    for (int z = 0; z < nz; ++z) {
      cx<T> uDp;
      if (zLay[z] == 3) {
        cx<T> g1 = ccosh(eta * (dD1 - zD[z]));
        cx<T> g2 = (ff1 * ccosh(eta * zD[z]) + ff2 * ccosh(eta * (1.0 - zD[z]))) / sh;
        uDp = g1 - g2;
      } else if (zLay[z] == 1) {
        uDp = ff3 * ccosh(eta * zD[z]);
      } else {
        cx<T> g2 = (ff1 * ccosh(eta * zD[z]) + ff2 * ccosh(eta * (1.0 - zD[z]))) / sh;
        uDp = 1.0 - g2;
      }
      u[(size_t)k * nz + z] = (uDf / prm.bD) * uDp;
    }
  }
  return worst;
}

template <class T> struct soln_ws {
  std::vector<cx<T>> udp, lt;
  std::vector<T> zx;
  std::vector<int> lx;
};

// laplace_hankel_solutions.f90:30-120  lap_hank_soln; fp[p*nz + z]
// lt = lapTime(p) (depends on p only; the reference recomputes it per call, :118)
template <class T>
void lap_hank_soln(T a, T rD, const std::vector<cx<T>> &p, const T *zD, const int *zLay, int nz,
                   const params &prm, const std::vector<cx<T>> &lt, soln_ws<T> &ws,
                   std::vector<cx<T>> &fp) {
  const int np = (int)p.size();
  fp.resize((size_t)np * nz);
  const int model = prm.model;
  if (model == 0) {
    for (int k = 0; k < np; ++k) {
      cx<T> th = theis(a, p[k]);
      for (int z = 0; z < nz; ++z) fp[(size_t)k * nz + z] = th;
    }
  } else if (model == 1) {
    hantush(a, zD, zLay, nz, p, prm, fp);
  } else if (model == 2) {
    if constexpr (std::is_same<T, double>::value) {
      hantushstorage(a, zD, zLay, nz, p, prm, fp);
    } else {
      // the Amos routine is double precision only (complex(DP) K, :219)
      std::vector<cx<double>> pd(np), out;
      std::vector<double> zd(nz);
      for (int k = 0; k < np; ++k) pd[k] = cx<double>((double)p[k].re, (double)p[k].im);
      for (int z = 0; z < nz; ++z) zd[z] = (double)zD[z];
      hantushstorage((double)a, zd.data(), zLay, nz, pd, prm, out);
      for (size_t i = 0; i < out.size(); ++i) fp[i] = cx<T>(out[i].re, out[i].im);
    }
  } else if (model == 6) {
    // laplace_hankel_solutions.f90:404-442  mishraNeumanMalama (MNtype 1; :107-110)
    const T beta0 = T(prm.mn_ak) * T(prm.mn_b);
    const T phiDa = T(prm.mn_psia) / T(prm.mn_b);
    const T phiDk = T(prm.mn_psik) / T(prm.mn_b);
    const T vartheta = beta0 * T(prm.mn_Sy) / (T(prm.Ss) * T(prm.mn_b)) * std::exp(-beta0 * (phiDa - phiDk));
    const T u0 = beta0 / T(2);
    for (int k = 0; k < np; ++k) {
      cx<T> eta1 = csqrt((p[k] * vartheta + a * a) / T(prm.kappa));
      cx<T> q = eta1 / u0;
      cx<T> v = csqrt(T(1) + q * q);
      cx<T> u = u0 * (T(1) - v);
      cx<T> etasq = (p[k] + a * a) / T(prm.kappa);
      cx<T> eta = csqrt(etasq);
      cx<T> Delta0 = eta * csinh(eta) - u * ccosh(eta);
      cx<T> pre = T(2) / (T(prm.kappa) * etasq);
      cx<T> ud = u / Delta0;
      for (int z = 0; z < nz; ++z)
        fp[(size_t)k * nz + z] = pre * (T(1) + ud * ccosh(eta * zD[z]));
    }
  } else {
    // models 3..5, laplace_hankel_solutions.f90:64-93
    const T MAXEXP = consts<T>::maxexp();
    ws.zx.assign(zD, zD + nz);
    ws.zx.push_back(T(1));
    ws.lx.assign(zLay, zLay + nz);
    ws.lx.push_back(3);
    if (model == 4) {
      ws.udp.resize((size_t)np * (nz + 1));
      for (int k = 0; k < np; ++k) {
        cx<T> th = theis(a, p[k]);
        for (int z = 0; z <= nz; ++z) ws.udp[(size_t)k * (nz + 1) + z] = th;
      }
    } else {
      hantush(a, ws.zx.data(), ws.lx.data(), nz + 1, p, prm, ws.udp);
    }
    for (int k = 0; k < np; ++k) {
      cx<T> eta = csqrt((p[k] + a * a) / T(prm.kappa));
      cx<T> xi = eta * T(prm.alphaD) / p[k];
      if (model == 3) {
        cx<T> s(0, 0);
        for (int m = 0; m < prm.moench_M; ++m)
          s = s + T(1) / (T(1) + p[k] * (T(1) / T(prm.moench_gamma[m])));
        xi = xi * T(prm.moench_M) / s;
      }
      const cx<T> *ud = &ws.udp[(size_t)k * (nz + 1)];
      cx<T> top = ud[nz];
      if (eta.re < MAXEXP) {
        cx<T> den = (T(1) + T(prm.beta) * eta * xi) * ccosh(eta) + xi * csinh(eta);
        for (int z = 0; z < nz; ++z)
          fp[(size_t)k * nz + z] = ud[z] - top * ccosh(eta * zD[z]) / den;
      } else {
        cx<T> den = T(1) + T(prm.beta) * eta * xi + xi;
        for (int z = 0; z < nz; ++z)
          fp[(size_t)k * nz + z] = ud[z] - top * cexp(eta * (zD[z] - T(1))) / den;
      }
    }
  }
  // common factor, :118   a*bessel_j0(a*rD)*fp*lapTime
  T aj = a * lm<T>::j0(a * rD);
  for (int k = 0; k < np; ++k)
    for (int z = 0; z < nz; ++z) fp[(size_t)k * nz + z] = aj * fp[(size_t)k * nz + z] * lt[k];
}

// ---------------------------------------------------------------------------
// Quadrature tables fixed for a run (driver.f90:79-91, 121-126, 141-151, 179-183)
template <class T> struct tables {
  int np, N, R, G, nacc;
  std::vector<T> hv;
  std::vector<int> Nv;
  std::vector<std::vector<T>> tsw;  // weights per level 1..R (index j-1)
  std::vector<T> glx, glw;
  void build(const params &prm) {
    np = 2 * prm.M + 1;
    N = (1 << prm.ts_k) - 1;
    R = prm.ts_R;
    nacc = prm.gl_nacc;
    G = prm.gl_ord - 2;
    hv.resize(R); Nv.resize(R); tsw.resize(R);
    for (int m = 1; m <= R; ++m) {
      int kv = prm.ts_k - R + m;
      Nv[m - 1] = (1 << kv) - 1;
      hv[m - 1] = T(4) / T(1 << kv);
      tanh_sinh_setup<T>(kv, T(0), tsw[m - 1], nullptr);
    }
    gauss_lobatto_setup<T>(prm.gl_ord, glx, glw);
  }
};

// One (t,r) iteration of driver.f90:100-231.  ts_scale = the `arg` used for the tanh-sinh
// ABSCISSAE (the reference keeps the arg of the first (t,r): driver.f90:121-126,274).
// infint is carried between calls as in the reference (driver.f90:209-211).
// flags[z] bit0 = infint left stale for at least one p.
template <class T>
void eval_column(const params &prm, const tables<T> &tb, T tD, int sv, T rD, T ts_scale,
                 int nz, const T *zD, const int *zLay, std::vector<cx<T>> &infint,
                 T *s_out, T *ds_out, int32_t *flags, bool inner_parallel) {
  const int np = tb.np, N = tb.N, R = tb.R, G = tb.G, nacc = tb.nacc;
  const T tee = T(prm.tee_mult) * tD;
  std::vector<cx<T>> p, lt;
  dehoog_pvalues<T>(tee, prm, p);
  lap_time<T>(prm, p, lt);
  std::vector<T> j0z(prm.n_j0z);
  for (int i = 0; i < prm.n_j0z; ++i) j0z[i] = T(prm.j0z[i]);  // j0z[j-1] = h%j0z(j)
  const T arg = j0z[sv - 1] / rD;  // driver.f90:120
  // tanh-sinh abscissae of the densest level, integration.f90:62 with s = ts_scale
  std::vector<T> a;
  {
    std::vector<T> wtmp;
    tanh_sinh_setup<T>(prm.ts_k, ts_scale, wtmp, &a);
  }
  const size_t blk = (size_t)np * nz;
  std::vector<cx<T>> fa((size_t)N * blk);
  // driver.f90:129-133 (OpenMP over abscissae, as the reference)
#pragma omp parallel if (inner_parallel)
  {
    soln_ws<T> ws;
    std::vector<cx<T>> fp;
#pragma omp for schedule(static)
    for (int nn = 0; nn < N; ++nn) {
      lap_hank_soln<T>(a[nn], rD, p, zD, zLay, nz, prm, lt, ws, fp);
      std::copy(fp.begin(), fp.end(), fa.begin() + (size_t)nn * blk);
    }
  }
  std::vector<cx<T>> finint(blk), tmp(R);
  for (int k = 0; k < np; ++k)
    for (int z = 0; z < nz; ++z) {
      for (int j = 1; j <= R; ++j) {
        const int step = 1 << (R - j);
        cx<T> sum(0, 0);
        for (int m = 1; m <= tb.Nv[j - 1]; ++m)
          sum = sum + tb.tsw[j - 1][m - 1] * fa[(size_t)(m * step - 1) * blk + (size_t)k * nz + z];
        tmp[j - 1] = (arg / T(2)) * sum;
      }
      finint[(size_t)k * nz + z] = (R > 1) ? extraptozero<T>(tb.hv.data(), tmp.data(), R) : tmp[0];
    }
  // Gauss-Lobatto between J0 zeros, driver.f90:187-203
  std::vector<cx<T>> GLarea((size_t)nacc * blk), GLz((size_t)G * blk);
  for (int j = sv + 1; j <= sv + nacc; ++j) {
    const T lob = j0z[j - 2] / rD, hib = j0z[j - 1] / rD;
    const T width = hib - lob;
#pragma omp parallel if (inner_parallel)
    {
      soln_ws<T> ws;
      std::vector<cx<T>> fp;
#pragma omp for schedule(static)
      for (int m = 0; m < G; ++m) {
        // GLy = (width*x + (hib+lob))/2.0 ; gfortran -O3 -march=native contracts to an FMA
        const T y = std::fma(width, tb.glx[m], hib + lob) / T(2);
        lap_hank_soln<T>(y, rD, p, zD, zLay, nz, prm, lt, ws, fp);
        std::copy(fp.begin(), fp.end(), GLz.begin() + (size_t)m * blk);
      }
    }
    for (size_t i = 0; i < blk; ++i) {
      cx<T> acc(0, 0);
      for (int m = 0; m < G; ++m) acc = acc + GLz[(size_t)m * blk + i] * tb.glw[m];
      GLarea[(size_t)(j - sv - 1) * blk + i] = (width / T(2)) * acc;
    }
  }
  if (infint.size() != blk) infint.assign(blk, cx<T>(0, 0));
  std::vector<cx<T>> series(nacc);
  for (int z = 0; z < nz; ++z) if (flags) flags[z] = 0;
  for (int k = 0; k < np; ++k)
    for (int z = 0; z < nz; ++z) {
      bool any = false;
      for (int j = 0; j < nacc; ++j) {
        series[j] = GLarea[(size_t)j * blk + (size_t)k * nz + z];
        if (cabs(series[j]) > T(0)) any = true;
      }
      if (any) infint[(size_t)k * nz + z] = wynn_epsilon<T>(series.data(), nacc);
      else if (flags) flags[z] |= 1;
    }
  std::vector<cx<T>> col(np);
  for (int z = 0; z < nz; ++z) {
    for (int k = 0; k < np; ++k) col[k] = finint[(size_t)k * nz + z] + infint[(size_t)k * nz + z];
    s_out[z] = dehoog_invlap<T>(tD, tee, col.data(), prm);
    for (int k = 0; k < np; ++k) col[k] = col[k] * p[k];
    ds_out[z] = dehoog_invlap<T>(tD, tee, col.data(), prm) * tD;
  }
}

}  // namespace orc
