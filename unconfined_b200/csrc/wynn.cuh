// Wynn-epsilon acceleration (integration.f90:125-189) for the kernels; host-compilable so
// that tests/hostcheck can compare its edge semantics with the oracle on the CPU.
#pragma once
#include "params.cuh"
#include "cmath.cuh"
#include "fast.cuh"

namespace unc {

// ---------------------------------------------------------------------------
// integration.f90:125-189  wynn_epsilon on nacc terms (two live columns, in place)
__host__ __device__ __noinline__ cplx wynn_dev(const cplx *series, int nacc) {
  cplx X[UNC_MAX_NACC + 1], Y[UNC_MAX_NACC + 1];  // 1-based; X: odd columns (starts as col -1), Y: even
  int ns = nacc;
  cplx run = mk(0.0, 0.0);
  for (int i = 1; i <= nacc; ++i) {
    if (!is_finite_c(series[i - 1])) {
      ns = i - 1;
      break;
    }
    run = run + series[i - 1];
    Y[i] = run;
    X[i] = mk(0.0, 0.0);
  }
  if (ns < nacc && ns < 4) return mk(-999999.875, 0.0);  // real(4) literal -999999.9
  const double eps = 2.220446049250313e-16;
  for (int j = 0; j <= ns - 2; ++j) {
    cplx *cur = (j & 1) ? X : Y;   // column j
    cplx *oth = (j & 1) ? Y : X;   // column j-1 -> becomes j+1
    for (int m = 1; m <= ns - (j + 1); ++m) {
      const cplx a = cur[m + 1], b = cur[m];
      const double dr = a.re - b.re, di = a.im - b.im;
      // abs(denom) > epsilon  <=>  |denom|^2 > eps^2 (no under/overflow in this range);
      // 1/denom = conj(denom)/|denom|^2 (one division; rounding differs by ~1 ulp from Smith)
      const double n2 = fma(dr, dr, di * di);
      if (n2 > eps * eps) {
        const double inv = rcp_fast(n2);
        const cplx o = oth[m + 1];
        oth[m] = mk(fma(dr, inv, o.re), fma(-di, inv, o.im));
      } else return a;
    }
  }
  return Y[2];
}

// Column order as wynn_dev, but a column is processed without the early exit inside it (the
// first |denom| <= epsilon of the column is remembered and returned after the column; what
// the later entries of that column hold no longer matters) and in blocks of four rows whose
// loads are issued together: the table lives in L2-backed local memory and one dependent
// round trip per entry is what the plain loop costs.
template <int NMAX>
__host__ __device__ __noinline__ cplx wynn_blk_t(const cplx *series, int nacc) {
  cplx X[NMAX + 4], Y[NMAX + 4];  // 1-based; X: odd columns (starts as col -1), Y: even
  int ns = nacc;
  cplx run = mk(0.0, 0.0);
  for (int i = 1; i <= nacc; ++i) {
    if (!is_finite_fastc(series[i - 1])) {
      ns = i - 1;
      break;
    }
    run = run + series[i - 1];
    Y[i] = run;
    X[i] = mk(0.0, 0.0);
  }
  if (ns < nacc && ns < 4) return mk(-999999.875, 0.0);  // real(4) literal -999999.9
  const double eps2 = 2.220446049250313e-16 * 2.220446049250313e-16;
  constexpr int UB = 4;
  for (int j = 0; j <= ns - 2; ++j) {
    cplx *cur = (j & 1) ? X : Y;   // column j
    cplx *oth = (j & 1) ? Y : X;   // column j-1 -> becomes j+1
    const int mmax = ns - (j + 1);
    bool hit = false;
    cplx ret = mk(0.0, 0.0);
    cplx b = cur[1];
    for (int m0 = 1; m0 <= mmax; m0 += UB) {
      cplx a[UB], o[UB];
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        const int m = min(m0 + u, mmax);      // clamped: stays inside the initialised rows
        a[u] = cur[m + 1];
        o[u] = oth[m + 1];
      }
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        if (m0 + u <= mmax) {
          const double dr = a[u].re - b.re, di = a[u].im - b.im;
          const double n2 = fma(dr, dr, di * di);
          const double inv = rcp_fast(n2);
          oth[m0 + u] = mk(fma(dr, inv, o[u].re), fma(-di, inv, o[u].im));
          if (!(n2 > eps2) && !hit) { hit = true; ret = a[u]; }   // integration.f90:169-177
          b = a[u];
        }
      }
    }
    if (hit) return ret;
  }
  return Y[2];
}

__host__ __device__ __forceinline__ cplx wynn_blk(const cplx *series, int nacc) {
#ifdef UNC_WYNN_SMALL
  if (nacc <= 12) return wynn_blk_t<12>(series, nacc);   // smaller thread-local columns
#endif
  return wynn_blk_t<UNC_MAX_NACC>(series, nacc);
}

// Anti-diagonal ("moving lozenge") order with ONE column-indexed array D[j*stride], meant for
// shared memory: D[j] = eps(n-j, j) of the last completed anti-diagonal n.  Exit semantics as
// wynn_reg: the reference returns at the first |denom| <= epsilon in (column, row) order, which
// in this order is the cancel with the smallest column seen, earlier rows first; once a cancel
// is known only smaller columns still matter (jlim).
__host__ __device__ __forceinline__ cplx wynn_loz(const cplx *series, int nacc, cplx *D, int stride) {
  int ns = nacc;
  for (int i = 0; i < nacc; ++i)
    if (!is_finite_fastc(series[i])) { ns = i; break; }
  if (ns < nacc && ns < 4) return mk(-999999.875, 0.0);  // real(4) literal -999999.9
  const double eps2 = 2.220446049250313e-16 * 2.220446049250313e-16;
  cplx run = mk(0.0, 0.0), best = mk(0.0, 0.0), keep_odd = mk(0.0, 0.0);
  int jlim = 1 << 30;  // no cancel seen yet
  for (int n = 1; n <= ns; ++n) {
    const cplx term = series[n - 1];
    run = mk(run.re + term.re, run.im + term.im);   // eps(n,0) = sum(series(1:n))
    cplx cur = run;
    cplx pm1 = mk(0.0, 0.0);                          // eps(:,-1) = 0
    const int jtop = min(n - 2, jlim - 1);
    for (int j = 0; j <= jtop; ++j) {
      const cplx a = D[j * stride];                   // eps(n-1-j, j)
      if (n == ns && j == ns - 3) keep_odd = a;       // eps(2, ns-3)
      D[j * stride] = cur;                            // eps(n-j, j)
      const double dr = cur.re - a.re, di = cur.im - a.im;
      const double n2 = fma(dr, dr, di * di);         // abs(denom) > epsilon(1.0)
      if (n2 > eps2) {
        const double inv = rcp_fast(n2);
        cur = mk(fma(dr, inv, pm1.re), fma(-di, inv, pm1.im));   // eps(n-j-1, j+1)
        pm1 = a;
      } else {
        best = cur;                                   // the reference returns eps(m+1, j)
        jlim = j;
        break;
      }
    }
    if (n - 1 < jlim) D[(n - 1) * stride] = cur;
  }
  if (jlim < (1 << 30)) return best;
  if (ns & 1) return keep_odd;
  return D[(ns - 2) * stride];
}

// The same algorithm with the epsilon table held in REGISTERS (north_star): anti-diagonal
// ("moving lozenge") order needs only one entry per column, D[j] = eps(n-j, j) of the last
// completed anti-diagonal n, instead of two full columns in local memory (which misses L1
// here because shared memory takes most of it).  The reference fills the table column by
// column and returns at the first |denom| <= epsilon in (column, row) order; in diagonal
// order that is the cancel with the smallest column seen, earlier rows first, and once a
// cancel is known only smaller columns still matter (jlim).  Fully unrolled for up to
// NMAX terms with lane-private predicates.
template <int NMAX>
__host__ __device__ __forceinline__ cplx wynn_reg(const cplx *series, int nacc) {
  int ns = nacc;
  for (int i = 0; i < nacc; ++i)
    if (!is_finite_c(series[i])) { ns = i; break; }
  if (ns < nacc && ns < 4) return mk(-999999.875, 0.0);  // real(4) literal -999999.9
  const double eps2 = 2.220446049250313e-16 * 2.220446049250313e-16;
  cplx D[NMAX];
#pragma unroll
  for (int j = 0; j < NMAX; ++j) D[j] = mk(0.0, 0.0);
  cplx run = mk(0.0, 0.0), best = mk(0.0, 0.0), keep_odd = mk(0.0, 0.0);
  int jlim = NMAX;  // no cancel seen yet
#pragma unroll
  for (int n = 1; n <= NMAX; ++n) {
    if (n <= ns) {
      const cplx term = series[n - 1];
      run = mk(run.re + term.re, run.im + term.im);  // eps(n,0) = sum(series(1:n))
      cplx cur = run;
      cplx pm1 = mk(0.0, 0.0);                        // eps(:,-1) = 0
#pragma unroll
      for (int j = 0; j <= n - 2; ++j) {
        if (j < jlim) {
          const cplx a = D[j];                        // eps(n-1-j, j)
          if (n == ns && j == ns - 3) keep_odd = a;   // eps(2, ns-3)
          D[j] = cur;                                 // eps(n-j, j)
          const double dr = cur.re - a.re, di = cur.im - a.im;
          const double n2 = fma(dr, dr, di * di);     // abs(denom) > epsilon(1.0)
          if (n2 > eps2) {
            const double inv = rcp_fast(n2);
            cur = mk(fma(dr, inv, pm1.re), fma(-di, inv, pm1.im));  // eps(n-j-1, j+1)
            pm1 = a;
          } else {
            best = cur;                               // reference returns eps(m+1, j)
            jlim = j;
          }
        }
      }
      if (n - 1 < jlim) D[n - 1] = cur;
    }
  }
  if (jlim < NMAX) return best;
  if (ns & 1) return keep_odd;
  cplx r = mk(0.0, 0.0);
#pragma unroll
  for (int j = 0; j < NMAX; ++j) if (j == ns - 2) r = D[j];
  return r;
}

__host__ __device__ __forceinline__ cplx wynn_any(const cplx *series, int nacc) {
#ifdef UNC_WYNN_REG
  if (nacc <= 12) return wynn_reg<12>(series, nacc);
#endif
  return wynn_dev(series, nacc);
}

}  // namespace unc
