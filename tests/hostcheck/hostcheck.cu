// TEST TOOL -- compiles the product's fast-path device math (unconfined_b200/csrc/fast.cuh)
// for the HOST so that CPU tests can compare the algebra of ap_terms_fast/eval_z_fast
// with the oracle at single (a,p,z) points.  Not linked into libunconfined_b200.so.
#include <cstring>
#include "../../unconfined_b200/csrc/fast.cuh"
#include "../../unconfined_b200/csrc/wynn.cuh"

extern "C" {

// f(z_i) = fp(a,p,z_i) of laplace_hankel_solutions.f90:30-116 without the common factor (:118).
// Returns 0 if the fast path declined (Re eta beyond its bound), 1 otherwise.
static double g_mn_vartheta = 0.0, g_mn_u0 = 1.0;
// run constants of model 6 (DevParams::mn_vartheta, mn_u0) for the next hc_fast_soln calls
void hc_set_mn(double vartheta, double u0) { g_mn_vartheta = vartheta; g_mn_u0 = u0; }

int hc_fast_soln(int model, double kappa, double alphaD, double beta, double lD, double dD, double bD,
                 int moench_M, double aux_re, double aux_im, double aux2_re, double aux2_im, double a,
                 double p_re, double p_im, int nz, const double *z, const int *lay, double *out,
                 double *eta_out) {
  unc::DevParams P;
  std::memset(&P, 0, sizeof P);
  P.mn_vartheta = g_mn_vartheta; P.mn_u0 = g_mn_u0;
  P.model = model; P.kappa = kappa; P.alphaD = alphaD; P.beta = beta;
  P.lD = lD; P.dD = dD; P.bD = bD; P.lD1 = 1.0 - lD; P.dD1 = 1.0 - dD; P.moench_M = moench_M;
  int mask = 0;
  for (int i = 0; i < nz; ++i) mask |= 1 << (lay[i] - 1);
  unc::cplx eta;
  unc::Coef co[3];
  bool ok = unc::ap_terms_fast(P, unc::mk(p_re, p_im), unc::mk(aux_re, aux_im),
                               unc::mk(aux2_re, aux2_im), a * a, 1.0, mask,
                               unc::fast_eta_max(P, mask, 1.0), &eta, co);
  eta_out[0] = eta.re; eta_out[1] = eta.im;
  if (!ok) return 0;
  for (int i = 0; i < nz; ++i) {
    unc::cplx f = unc::eval_z_fast(eta, co[lay[i] - 1], z[i]);
    out[2 * i] = f.re; out[2 * i + 1] = f.im;
  }
  return 1;
}

void hc_exp_pm(double x, double *out) {
  unc::rexp e = unc::exp_pm(x);
  out[0] = e.ep; out[1] = e.em; out[2] = e.ch; out[3] = e.sh;
}

// which: 0 = dispatcher used by the kernels, 1 = local-memory version, 2 = register version, 3 = blocked,
// 4 = single-array lozenge (shared-memory version; here on a plain buffer with stride 3),
// 5 = the same with two anti-diagonals in lockstep (the grid kernel's)
void hc_wynn(const double *series, int n, int which, double *out) {
  unc::cplx s[UNC_MAX_NACC];
  for (int i = 0; i < n; ++i) s[i] = unc::mk(series[2 * i], series[2 * i + 1]);
  unc::cplx Dbuf[3 * UNC_MAX_NACC];
  if (which == 5) { unc::cplx r5 = unc::wynn_loz2(s, n, Dbuf, 3); out[0] = r5.re; out[1] = r5.im; return; }
  if (which == 4) { unc::cplx r4 = unc::wynn_loz(s, n, Dbuf, 3); out[0] = r4.re; out[1] = r4.im; return; }
  unc::cplx r = which == 1 ? unc::wynn_dev(s, n) : (which == 2 ? unc::wynn_reg<12>(s, n) : (which == 3 ? unc::wynn_blk(s, n) : unc::wynn_any(s, n)));
  out[0] = r.re; out[1] = r.im;
}

void hc_sincos(double y, double *out) { unc::sincos_q(y, &out[0], &out[1]); }

}  // extern "C"
