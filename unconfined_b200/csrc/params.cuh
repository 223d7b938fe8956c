// Plain-data structures shared by the host planner (capi.cu) and the kernels.
#pragma once
#include <cstdint>

namespace unc {

#if defined(UNC_BUDGET_NEVILLE) && !defined(UNC_BUDGET_SEQSUM)
#define UNC_BUDGET_SEQSUM
#endif
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
  const double *ts_lw;         // error-budget builds only (UNC_BUDGET_NEVILLE): the R levels' own weights, concatenated
  int ts_k;                    // ditto: densest tanh-sinh level
};

struct Job {
  long long ncol;   // number of (t,r) columns (grid) or points
  int nz;           // z-values per column
  long long col0;   // grid jobs of one shard: local column c is global column c + col0
  long long tdiv;   // column c uses tD[(c+col0) / tdiv], sv[(c+col0) / tdiv]
  long long rmod;   // and rD[(c+col0) % rmod]
  int zstride;      // z of column c starts at zD + c*zstride (0: shared grid z, 1: points)
  const double *tD;
  const int *sv;
  const double *rD;
  const double *zD;
  const int *zLay;
  const double *ts_scale;  // per local column, or NULL (fresh abscissae)
  double *s, *ds;
  int *flags;
  unsigned long long *smask;  // per point, or NULL: bit p set = infint(p,z) was stale (driver.f90:209)
  unsigned int *nstale;       // or NULL: incremented once per point with a non-zero mask
  // carry post-pass (lh_point_kernel<1> only; capi.cu carry_postpass): CTA e works on point fix_list[e]
  int fix_mode;               // 0 normal; 3 normal, units taken in the order of fix_list (cost-ordered points);
                              // 1 source: Wynn result of the p in need[] -> fix_val;
                              // 2 destination: stale infint(p) taken from fix_val[fix_src[e*np+p]]
  long long fix_n;            // number of entries
  const int *fix_list;        // [entries] point index c*nz+z
  const unsigned long long *fix_need;  // mode 1: [npoints] which p of a source point are wanted
  const int *fix_src;         // mode 2: [entries*np] source entry of each stale p, or -1 (none: 0)
  double *fix_val;            // [source entries * np * 2] carried infint values (re, im)
  double *fix_s, *fix_ds;     // mode 2: compact outputs [entries] (also written to s/ds if non-NULL)
};

}  // namespace unc
