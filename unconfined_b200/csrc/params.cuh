// Plain-data structures shared by the host planner (capi.cu) and the kernels.
#pragma once
#include <cstdint>

namespace unc {

#define UNC_MAX_NACC 32
#define UNC_WARPS 8
#define UNC_THREADS (UNC_WARPS * 32)

struct DevParams {
  int model, M, np, N, R, G, nacc, gl_rounds, nts_pad, time_type, n_time_par, moench_M, n_j0z;
  double alpha, log_tol, tee_mult, kappa, alphaD, beta, lD, dD, bD, rDw, CDw, tDb, lD1, dD1;
  double mn_vartheta, mn_u0;   // model 6 / MNtype 1 (laplace_hankel_solutions.f90:424-431)
  const double *ts_T;          // [N]   tanh(u2)+1           (integration.f90:62)
  const double *ts_wc;         // [N]   Richardson-combined tanh-sinh weights
  const double *gl_x;          // [G]   Gauss-Lobatto interior nodes
  const double *gl_w;          // [G]
  const double *j0z;           // [n_j0z]
  const double *time_par;      // [n_time_par]
  const double *moench_gamma;  // [moench_M]
};

struct Job {
  long long ncol;   // number of (t,r) columns (grid) or points
  int nz;           // z-values per column
  long long tdiv;   // column c uses tD[c / tdiv], sv[c / tdiv]
  long long rmod;   // and rD[c % rmod]
  int zstride;      // z of column c starts at zD + c*zstride (0: shared grid z, 1: points)
  const double *tD;
  const int *sv;
  const double *rD;
  const double *zD;
  const int *zLay;
  const double *ts_scale;  // per column, or NULL (fresh abscissae)
  double *s, *ds;
  int *flags;
};

}  // namespace unc
