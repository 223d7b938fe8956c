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

// wynn_loz with TWO anti-diagonals in flight.  A diagonal is one dependent chain (difference,
// |.|^2, reciprocal with two Newton steps, FMA: ~10 FP64 instructions deep per entry, 66 entries
// for 12 terms), during which the FP64 pipe idles two thirds of the time.  Entry j of diagonal
// n+1 needs entry j of diagonal n (what wynn_loz parks in D[j]) and its own entry j-1, so the
// diagonals n ("A") and n+1 ("B") run in lockstep, B one column behind A and fed from A through a
// register: iteration i does A's column i and B's column i-1, two independent chains.  A never
// writes D (B overwrites every column it would have written; columns B does not reach are dead,
// see below).  The arithmetic of every entry is that of wynn_loz, bit for bit.
// Exit semantics: a step (diagonal, column j) is executed iff j < jlim at that moment; within an
// iteration A's step is settled before B's.  B is always behind A, so this admits exactly the
// steps of the sequential order (all of A, then all of B) whose cancels can still win: the
// reference returns the cancel with the smallest column, earliest row first.
__host__ __device__ __forceinline__ cplx wynn_loz2(const cplx *series, int nacc, cplx *D, int stride) {
  int ns = nacc;
  for (int i = 0; i < nacc; ++i)
    if (!is_finite_fastc(series[i])) { ns = i; break; }
  if (ns < nacc && ns < 4) return mk(-999999.875, 0.0);  // real(4) literal -999999.9
  const double eps2 = 2.220446049250313e-16 * 2.220446049250313e-16;
  cplx run = mk(0.0, 0.0), best = mk(0.0, 0.0), keep_odd = mk(0.0, 0.0);
  int jlim = 1 << 30;  // no cancel seen yet
  // diagonals n0..ns go in pairs (so that the last one, which captures eps(2, ns-3), is a "B")
  const int n0 = ((ns - 1) & 1) ? 3 : 2;
  for (int n = 1; n < n0 && n <= ns; ++n) {   // diagonal 1 (and 2): as wynn_loz
    const cplx term = series[n - 1];
    run = mk(run.re + term.re, run.im + term.im);
    cplx cur = run;
    if (n == 2) {
      const cplx a = D[0];
      if (n == ns && 0 == ns - 3) keep_odd = a;
      D[0] = cur;
      const double dr = cur.re - a.re, di = cur.im - a.im;
      const double n2 = fma(dr, dr, di * di);
      if (n2 > eps2) {
        const double inv = rcp_fast(n2);
        cur = mk(fma(dr, inv, 0.0), fma(-di, inv, 0.0));   // eps(:,-1) = 0
      } else {
        best = cur;
        jlim = 0;
      }
    }
    if (n - 1 < jlim) D[(n - 1) * stride] = cur;
  }
  for (int n = n0; n + 1 <= ns; n += 2) {
    const cplx tA = series[n - 1], tB = series[n];
    const cplx runA = mk(run.re + tA.re, run.im + tA.im);
    run = mk(runA.re + tB.re, runA.im + tB.im);
    cplx curA = runA, pm1A = mk(0.0, 0.0), curB = run, pm1B = mk(0.0, 0.0);
    cplx sav = mk(0.0, 0.0);        // A's value at column i-1 = what wynn_loz would have left in D[i-1]
    bool liveA = true, liveB = true;
    const bool lastB = (n + 1 == ns);
    for (int i = 0; i <= n; ++i) {
      const cplx cA = curA;
      const bool doA = liveA && i <= n - 2 && i < jlim;
      // both chains' arithmetic first (independent), the bookkeeping after
      const cplx aA = doA ? D[i * stride] : mk(0.0, 0.0);
      const double drA = cA.re - aA.re, diA = cA.im - aA.im;
      const double n2A = fma(drA, drA, diA * diA);
      const double invA = rcp_fast(n2A);
      const cplx newA = mk(fma(drA, invA, pm1A.re), fma(-diA, invA, pm1A.im));
      const cplx aB = sav;
      const double drB = curB.re - aB.re, diB = curB.im - aB.im;
      const double n2B = fma(drB, drB, diB * diB);
      const double invB = rcp_fast(n2B);
      const cplx newB = mk(fma(drB, invB, pm1B.re), fma(-diB, invB, pm1B.im));
      // selects only (no branches): both chains stay in one basic block and ptxas interleaves them
      const bool okA = doA && (n2A > eps2), cxA = doA && !(n2A > eps2);
      curA = okA ? newA : curA;
      pm1A = okA ? aA : pm1A;
      best = cxA ? cA : best;
      jlim = cxA ? i : jlim;
      liveA = okA;
      const int j = i - 1;
      const bool doB = liveB && j >= 0 && j < jlim;      // j <= n-1 = (n+1)-2 by the loop bound
      const bool okB = doB && (n2B > eps2), cxB = doB && !(n2B > eps2);
      keep_odd = (doB && lastB && j == ns - 3) ? aB : keep_odd;
      if (doB) D[j * stride] = curB;
      best = cxB ? curB : best;
      jlim = cxB ? j : jlim;
      curB = okB ? newB : curB;
      pm1B = okB ? aB : pm1B;
      liveB = okB || (j < 0);
      sav = cA;
      if (!liveA && !liveB) break;
    }
    if (n < jlim) D[n * stride] = curB;
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
