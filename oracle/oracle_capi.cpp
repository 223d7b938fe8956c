// TEST INFRASTRUCTURE ONLY -- C entry points of the CPU oracle (PARITY UNPINNED, see
// oracle_math.hpp).  Loaded through ctypes by tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs; never by the product.
#include <cstdint>
#include <cstring>
#include <vector>
#include <cmath>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "oracle_core.hpp"

using namespace orc;

namespace {

template <class T>
int eval_grid_t(const params *prm, int nt, const double *tD, const int32_t *sv, int nr,
                const double *rD, int nz, const double *zD, const int32_t *zLay,
                const double *ts_scale, int carry, int nthreads, double *totint,
                double *totintd, int32_t *flags) {
  tables<T> tb;
  tb.build(*prm);
  std::vector<T> z(nz);
  std::vector<int> lay(zLay, zLay + nz);
  for (int i = 0; i < nz; ++i) z[i] = T(zD[i]);
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  auto scale_of = [&](int i, int k) -> T {
    if (ts_scale) return T(ts_scale[(size_t)k + (size_t)nr * i]);
    return T(prm->j0z[sv[i] - 1]) / T(rD[k]);
  };
  if (carry) {
    // driver.f90:100-276 order: t outer, r inner; infint persists (driver.f90:209-211)
    std::vector<cx<T>> infint;
    std::vector<T> s(nz), ds(nz);
    std::vector<int32_t> fl(nz);
    for (int i = 0; i < nt; ++i)
      for (int k = 0; k < nr; ++k) {
        eval_column<T>(*prm, tb, T(tD[i]), sv[i], T(rD[k]), scale_of(i, k), nz, z.data(),
                       lay.data(), infint, s.data(), ds.data(), fl.data(), true);
        size_t base = (size_t)nz * ((size_t)k + (size_t)nr * i);
        for (int m = 0; m < nz; ++m) {
          totint[base + m] = (double)s[m];
          totintd[base + m] = (double)ds[m];
          if (flags) flags[base + m] = fl[m];
        }
      }
  } else {
    const long ncol = (long)nt * nr;
#pragma omp parallel
    {
      std::vector<cx<T>> infint;
      std::vector<T> s(nz), ds(nz);
      std::vector<int32_t> fl(nz);
#pragma omp for schedule(dynamic, 1)
      for (long c = 0; c < ncol; ++c) {
        int i = (int)(c / nr), k = (int)(c % nr);
        infint.assign((size_t)tb.np * nz, cx<T>(0, 0));
        eval_column<T>(*prm, tb, T(tD[i]), sv[i], T(rD[k]), scale_of(i, k), nz, z.data(),
                       lay.data(), infint, s.data(), ds.data(), fl.data(), false);
        size_t base = (size_t)nz * (size_t)c;
        for (int m = 0; m < nz; ++m) {
          totint[base + m] = (double)s[m];
          totintd[base + m] = (double)ds[m];
          if (flags) flags[base + m] = fl[m];
        }
      }
    }
  }
  return 0;
}

template <class T>
int eval_points_t(const params *prm, int64_t n, const double *tD, const int32_t *sv,
                  const double *rD, const double *zD, const int32_t *zLay,
                  const double *ts_scale, int nthreads, double *s, double *ds, int32_t *flags) {
  tables<T> tb;
  tb.build(*prm);
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
  {
    std::vector<cx<T>> infint;
#pragma omp for schedule(dynamic, 1)
    for (int64_t c = 0; c < n; ++c) {
      infint.assign((size_t)tb.np, cx<T>(0, 0));
      T z = T(zD[c]);
      int lay = zLay[c];
      T so, dso;
      int32_t fl;
      T sc = ts_scale ? T(ts_scale[c]) : T(prm->j0z[sv[c] - 1]) / T(rD[c]);
      eval_column<T>(*prm, tb, T(tD[c]), sv[c], T(rD[c]), sc, 1, &z, &lay, infint, &so, &dso, &fl,
                     false);
      s[c] = (double)so;
      ds[c] = (double)dso;
      if (flags) flags[c] = fl;
    }
  }
  return 0;
}

}  // namespace

extern "C" {

// ulps = 0 disables; see oracle_math.hpp (jitter)
void orc_set_jitter(double ulps, unsigned long long seed) {
  jitter().ulps = ulps;
  jitter().seed = seed;
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

int orc_eval_grid(const params *prm, int nt, const double *tD, const int32_t *sv, int nr,
                  const double *rD, int nz, const double *zD, const int32_t *zLay,
                  const double *ts_scale, int carry, int nthreads, double *totint,
                  double *totintd, int32_t *flags) {
  return eval_grid_t<double>(prm, nt, tD, sv, nr, rD, nz, zD, zLay, ts_scale, carry, nthreads,
                             totint, totintd, flags);
}

// same algorithm in x87 long double (inputs/outputs double): rounding-floor probe only
int orc_eval_grid_ld(const params *prm, int nt, const double *tD, const int32_t *sv, int nr,
                     const double *rD, int nz, const double *zD, const int32_t *zLay,
                     const double *ts_scale, int carry, int nthreads, double *totint,
                     double *totintd, int32_t *flags) {
  return eval_grid_t<long double>(prm, nt, tD, sv, nr, rD, nz, zD, zLay, ts_scale, carry,
                                  nthreads, totint, totintd, flags);
}

int orc_eval_points(const params *prm, int64_t n, const double *tD, const int32_t *sv,
                    const double *rD, const double *zD, const int32_t *zLay,
                    const double *ts_scale, int nthreads, double *s, double *ds, int32_t *flags) {
  return eval_points_t<double>(prm, n, tD, sv, rD, zD, zLay, ts_scale, nthreads, s, ds, flags);
}

int orc_eval_points_ld(const params *prm, int64_t n, const double *tD, const int32_t *sv,
                       const double *rD, const double *zD, const int32_t *zLay,
                       const double *ts_scale, int nthreads, double *s, double *ds,
                       int32_t *flags) {
  return eval_points_t<long double>(prm, n, tD, sv, rD, zD, zLay, ts_scale, nthreads, s, ds,
                                    flags);
}

// ---- components -----------------------------------------------------------
void orc_tanh_sinh(int k, double s, double *w, double *a) {
  std::vector<double> wv, av;
  tanh_sinh_setup<double>(k, s, wv, a ? &av : nullptr);
  std::memcpy(w, wv.data(), wv.size() * sizeof(double));
  if (a) std::memcpy(a, av.data(), av.size() * sizeof(double));
}

void orc_gauss_lobatto(int ord, double *x, double *w) {
  std::vector<double> xv, wv;
  gauss_lobatto_setup<double>(ord, xv, wv);
  std::memcpy(x, xv.data(), xv.size() * sizeof(double));
  std::memcpy(w, wv.data(), wv.size() * sizeof(double));
}

int orc_wynn(const double *series, int n, double *out) {
  std::vector<cx<double>> s(n);
  for (int i = 0; i < n; ++i) s[i] = cx<double>(series[2 * i], series[2 * i + 1]);
  int info;
  cx<double> r = wynn_epsilon<double>(s.data(), n, &info);
  out[0] = r.re; out[1] = r.im;
  return info;
}

void orc_extrap(const double *x, const double *y, int n, double *out) {
  std::vector<cx<double>> yy(n);
  for (int i = 0; i < n; ++i) yy[i] = cx<double>(y[2 * i], y[2 * i + 1]);
  cx<double> r = extraptozero<double>(x, yy.data(), n);
  out[0] = r.re; out[1] = r.im;
}

void orc_pvalues(const params *prm, double tee, double *p) {
  std::vector<cx<double>> pv;
  dehoog_pvalues<double>(tee, *prm, pv);
  for (size_t i = 0; i < pv.size(); ++i) { p[2 * i] = pv[i].re; p[2 * i + 1] = pv[i].im; }
}

double orc_dehoog(const params *prm, double t, double tee, const double *fp) {
  const int np = 2 * prm->M + 1;
  std::vector<cx<double>> f(np);
  for (int i = 0; i < np; ++i) f[i] = cx<double>(fp[2 * i], fp[2 * i + 1]);
  return dehoog_invlap<double>(t, tee, f.data(), *prm);
}

void orc_lap_time(const params *prm, int np, const double *p, double *out) {
  std::vector<cx<double>> pv(np), m;
  for (int i = 0; i < np; ++i) pv[i] = cx<double>(p[2 * i], p[2 * i + 1]);
  lap_time<double>(*prm, pv, m);
  for (int i = 0; i < np; ++i) { out[2 * i] = m[i].re; out[2 * i + 1] = m[i].im; }
}

int orc_cbesk01(double zr, double zi, double *out, int *nz) {
  cx<double> K[2];
  int ierr = cbesk01(cx<double>(zr, zi), K, nz);
  out[0] = K[0].re; out[1] = K[0].im; out[2] = K[1].re; out[3] = K[1].im;
  return ierr;
}

// lap_hank_soln at one abscissa for the p-vector of time tD; out[(p*nz+z)*2 + {0,1}]
void orc_soln(const params *prm, double a, double rD, double tD, int nz, const double *zD,
              const int32_t *zLay, double *out) {
  std::vector<cx<double>> p, lt, fp;
  dehoog_pvalues<double>(prm->tee_mult * tD, *prm, p);
  lap_time<double>(*prm, p, lt);
  soln_ws<double> ws;
  std::vector<int> lay(zLay, zLay + nz);
  lap_hank_soln<double>(a, rD, p, zD, lay.data(), nz, *prm, lt, ws, fp);
  for (size_t i = 0; i < fp.size(); ++i) { out[2 * i] = fp[i].re; out[2 * i + 1] = fp[i].im; }
}

// driver_io.f90:628-647  zeros of J0 by Newton from (i+3/4)pi
void orc_j0_zeros(int terms, double *j0z) {
  const double PIEP = 4.0 * std::atan(1.0);
  for (int i = 0; i < terms; ++i) {
    double x = (i + 0.75) * PIEP;
    for (;;) {
      double dx = ::j0(x) / ::j1(x);
      x = x + dx;
      if (std::fabs(dx) < (std::nextafter(std::fabs(x), INFINITY) - std::fabs(x))) break;  // spacing(x)
    }
    j0z[i] = x;
  }
}

// driver_io.f90:658-664  split index between finite and infinite integrals
void orc_split_index(int nt, const double *tD, int j0s0, int j0s1, int32_t *sv) {
  int zrange = std::max(j0s0, j0s1) - std::min(j0s0, j0s1);
  double mn = INFINITY, mxv = -INFINITY;
  for (int i = 0; i < nt; ++i) {
    double l = std::log10(tD[i]);
    mn = std::min(mn, l); mxv = std::max(mxv, l);
  }
  int minlsp = (int)std::floor(mn), maxlsp = (int)std::ceil(mxv);
  int spRange = maxlsp - minlsp + 1;
  for (int i = 0; i < nt; ++i)
    sv[i] = std::min(j0s0, j0s1) + (int)(zrange * ((maxlsp - std::log10(tD[i])) / spRange));
}

// driver_io.f90:572-586  layer of each z
void orc_zlay(int nz, const double *zD, double lD, double dD, int32_t *lay) {
  for (int i = 0; i < nz; ++i) {
    if (zD[i] <= 0.0 || zD[i] < (1.0 - lD)) lay[i] = 1;
    else if ((zD[i] - 1.0) >= 0.0 || zD[i] < (1.0 - dD)) lay[i] = 2;
    else lay[i] = 3;
  }
}

}  // extern "C"
