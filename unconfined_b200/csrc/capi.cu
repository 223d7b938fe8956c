// C ABI of unconfined_b200 (include/unconfined_b200.h): host planner + launches.
//
// Host side = what the reference computes once per run on the host before its loop
// nest (driver.f90:79-91 level tables, integration.f90:31-120 quadrature set-up), plus
// sharding of the (t,r) columns over the GPUs of one box and the copies in/out.
// There is deliberately no CPU evaluation path here.
#include <cuda_runtime.h>
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <limits>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/unconfined_b200.h"
#include "kernels.cuh"

namespace {

thread_local std::string g_err;
std::mutex g_mutex;                 // calls are serialised (threading)
std::atomic<long long> g_launches{0};
int g_device = 0;

int fail(int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CK(call)                                                                      \
  do {                                                                                \
    cudaError_t e_ = (call);                                                          \
    if (e_ != cudaSuccess)                                                            \
      return fail(UNC_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                  __FILE__, __LINE__);                                                \
  } while (0)

// ---- host set-up tables -----------------------------------------------------
// integration.f90:45-62: weights of tanh-sinh level k (renormalised to sum 2) and,
// for the densest level, T = tanh(u2)+1 so that a = T*s/2.
void tanh_sinh_level(int k, std::vector<double> &w, std::vector<double> *T) {
  const int N = (1 << k) - 1, r = (N - 1) / 2;
  const double h = 4.0 / (double)(1 << k);
  const double piov2 = 2.0 * std::atan(1.0);
  w.resize(N);
  if (T) T->resize(N);
  double sum = 0.0;
  for (int i = -r; i <= r; ++i) {
    const double u1 = piov2 * std::cosh(h * i), u2 = piov2 * std::sinh(h * i);
    const double c = std::cosh(u2);
    w[i + r] = u1 / (c * c);
    if (T) (*T)[i + r] = std::tanh(u2) + 1.0;
  }
  for (int i = 0; i < N; ++i) sum += w[i];
  for (int i = 0; i < N; ++i) w[i] = 2.0 * w[i] / sum;
}

// integration.f90:192-237 is linear in y for fixed x: return the coefficient of each
// y_j in the value extrapolated to x=0 (Neville on unit vectors, long double).
std::vector<long double> richardson_coeffs(const std::vector<double> &x) {
  const int n = (int)x.size();
  std::vector<long double> coef(n);
  for (int u = 0; u < n; ++u) {
    std::vector<long double> c(n, 0.0L), d(n, 0.0L);
    c[u] = d[u] = 1.0L;
    int ns = 1;
    for (int i = 2; i <= n; ++i) if (x[i - 1] < x[ns - 1]) ns = i;
    long double y = c[ns - 1];
    ns -= 1;
    for (int m = 1; m <= n - 1; ++m) {
      for (int i = 1; i <= n - m; ++i) {
        long double den = (c[i] - d[i - 1]) / ((long double)x[i - 1] - (long double)x[i + m - 1]);
        d[i - 1] = (long double)x[i + m - 1] * den;
        c[i - 1] = (long double)x[i - 1] * den;
      }
      long double dy;
      if (2 * ns < n - m) dy = c[ns];
      else { dy = d[ns - 1]; ns -= 1; }
      y += dy;
    }
    coef[u] = y;
  }
  return coef;
}

// integration.f90:70-120: interior Gauss-Lobatto nodes and weights
void gauss_lobatto(int ord, std::vector<double> &gx, std::vector<double> &gw) {
  const int N = ord - 1, N1 = N + 1;
  const double pi = 4.0 * std::atan(1.0);
  std::vector<double> x(ord), xold(ord, 2.0), Pn(ord), Pn1(ord);
  for (int i = 0; i <= N; ++i) x[i] = std::cos(pi * i / N);
  for (;;) {
    double mx = 0.0;
    for (int i = 0; i < ord; ++i) mx = std::max(mx, std::fabs(x[i] - xold[i]));
    if (!(mx > std::numeric_limits<double>::epsilon())) break;
    xold = x;
    for (int i = 0; i < ord; ++i) {
      // Legendre recurrence up to P_N (Pn) and P_{N-1} (Pn1)
      double pkm1 = 1.0, pk = x[i];
      for (int k = 2; k <= N; ++k) {
        double pkp1 = ((2 * k - 1) * x[i] * pk - (k - 1) * pkm1) / k;
        pkm1 = pk;
        pk = pkp1;
      }
      Pn[i] = pk;
      Pn1[i] = pkm1;
      x[i] = xold[i] - (x[i] * pk - pkm1) / (N1 * pk);
    }
  }
  gx.resize(ord - 2);
  gw.resize(ord - 2);
  for (int i = 1; i < ord - 1; ++i) {
    gx[i - 1] = x[i];
    gw[i - 1] = 2.0 / ((double)(N * N1) * (Pn[i] * Pn[i]));
  }
}

struct HostPlan {
  unc::DevParams P;
  std::vector<double> blob;  // ts_T | ts_wc | gl_x | gl_w | j0z | time_par | moench_gamma
  size_t off_T, off_wc, off_glx, off_glw, off_j0z, off_tp, off_mg;
};

int make_plan(const unc_params *prm, HostPlan &hp) {
  if (!prm) return fail(UNC_ERR_BAD_ARG, "prm is NULL");
  if (prm->model < 0 || prm->model > 6) return fail(UNC_ERR_BAD_ARG, "invalid model %d", prm->model);
  if (prm->model == 6) {
    if (prm->mn_type == 0 || prm->mn_type == 2)
      return fail(UNC_ERR_UNSUPPORTED, "model 6 MNtype %d (ARB quad precision / finite-difference "
                  "Mishra-Neuman) is out of scope; only MNtype 1 is supported", prm->mn_type);
    if (prm->mn_type != 1) return fail(UNC_ERR_BAD_ARG, "invalid MNtype %d (driver_io.f90:272)", prm->mn_type);
    if (!(prm->mn_b > 0.0) || !(prm->Ss > 0.0) || !(prm->mn_ak > 0.0))
      return fail(UNC_ERR_BAD_ARG, "model 6 needs mn_b, Ss, mn_ak > 0");
  }
  if (prm->M < 2) return fail(UNC_ERR_BAD_ARG, "de Hoog M must be >= 2 (driver_io.f90:308)");
  if (prm->M > 31) return fail(UNC_ERR_UNSUPPORTED, "de Hoog M > 31 not supported (2M+1 <= 63)");
  if (prm->ts_R < 1 || prm->ts_k - prm->ts_R < 2)
    return fail(UNC_ERR_BAD_ARG, "tanh-sinh needs R>=1 and k-R>=2 (driver_io.f90:317-328)");
  if (prm->ts_k > 12) return fail(UNC_ERR_UNSUPPORTED, "tanh-sinh k > 12 not supported");
  if (prm->gl_nacc < 2 || prm->gl_nacc > UNC_MAX_NACC)
    return fail(UNC_ERR_UNSUPPORTED, "gl_nacc must be in 2..%d", UNC_MAX_NACC);
  if (prm->gl_ord < 3) return fail(UNC_ERR_BAD_ARG, "gl_ord must be >= 3");
  if (!prm->j0z || prm->n_j0z < prm->gl_nacc + 2)
    return fail(UNC_ERR_BAD_ARG, "j0z must hold at least nacc+2 zeros");
  if (prm->time_type == 0 || prm->time_type > 8)
    return fail(UNC_ERR_BAD_ARG, "invalid time_type %d (time.f90:47-121)", prm->time_type);
  {
    int need = 2;
    if (prm->time_type < 0 && prm->time_type >= -100) need = 2 * (-prm->time_type) + 1;
    if (prm->time_type <= -101) need = 2 * (-prm->time_type - 100) + 1;
    if (!prm->time_par || prm->n_time_par < need)
      return fail(UNC_ERR_BAD_ARG, "time_par needs %d entries for time_type %d", need, prm->time_type);
  }
  if (prm->model == 3 && (prm->moench_M < 1 || !prm->moench_gamma))
    return fail(UNC_ERR_BAD_ARG, "model 3 needs moench_M >= 1 (driver_io.f90:144-149)");
  if (!(prm->tol > 0.0)) return fail(UNC_ERR_BAD_ARG, "tol must be > 0");

  unc::DevParams &P = hp.P;
  std::memset(&P, 0, sizeof P);
  P.model = prm->model;
  P.M = prm->M;
  P.np = 2 * prm->M + 1;            // driver.f90:79
  P.N = (1 << prm->ts_k) - 1;       // driver.f90:81
  P.R = prm->ts_R;
  P.G = prm->gl_ord - 2;
  P.nacc = prm->gl_nacc;
  P.gl_rounds = (P.nacc * P.G + 31) / 32;
  if (P.gl_rounds > P.G)
    return fail(UNC_ERR_UNSUPPORTED, "nacc*(ord-2)/32 must not exceed ord-2 (nacc <= 32)");
  P.nts_pad = ((P.N + 31) / 32) * 32;
  P.time_type = prm->time_type;
  P.n_time_par = prm->n_time_par;
  P.moench_M = prm->model == 3 ? prm->moench_M : 0;
  P.n_j0z = prm->n_j0z;
  P.alpha = prm->alpha;
  P.log_tol = std::log(prm->tol);   // invlap.f90:77,166 (glibc log, as the reference)
  P.tee_mult = prm->tee_mult;
  P.kappa = prm->kappa;
  P.alphaD = prm->alphaD;
  P.beta = prm->beta;
  P.lD = prm->lD; P.dD = prm->dD; P.bD = prm->bD; P.rDw = prm->rDw;
  P.lD1 = 1.0 - prm->lD;            // laplace_hankel_solutions.f90:159-160
  P.dD1 = 1.0 - prm->dD;
  if (prm->model == 6) {
    // laplace_hankel_solutions.f90:424-431 (run constants of mishraNeumanMalama)
    const double beta0 = prm->mn_ak * prm->mn_b;
    const double phiDa = prm->mn_psia / prm->mn_b;
    const double phiDk = prm->mn_psik / prm->mn_b;
    P.mn_vartheta = beta0 * prm->mn_Sy / (prm->Ss * prm->mn_b) * std::exp(-beta0 * (phiDa - phiDk));
    P.mn_u0 = beta0 / 2.0;
  }
  if (prm->model == 2) {
    const double PI = 4.0 * std::atan(1.0);
    P.CDw = prm->rDw * prm->rDw / (2.0 * (prm->l - prm->d) * prm->Ss);   // :248
    P.tDb = PI * prm->rDwobs * prm->rDwobs / (prm->sF * prm->Ss);        // :251
  }
  // quadrature tables
  std::vector<double> T, wc(P.N, 0.0), w;
  std::vector<double> hv(P.R);
  std::vector<std::vector<double>> lw(P.R);
  for (int m = 1; m <= P.R; ++m) {
    int kv = prm->ts_k - P.R + m;                    // driver.f90:86-91
    hv[m - 1] = 4.0 / (double)(1 << kv);
    tanh_sinh_level(kv, lw[m - 1], m == P.R ? &T : nullptr);
  }
  std::vector<long double> rc(P.R, 1.0L);
  if (P.R > 1) rc = richardson_coeffs(hv);
  for (int nn = 1; nn <= P.N; ++nn) {
    long double acc = 0.0L;
    for (int j = 1; j <= P.R; ++j) {
      int step = 1 << (P.R - j);                     // driver.f90:150
      if (nn % step == 0) acc += rc[j - 1] * (long double)lw[j - 1][nn / step - 1];
    }
    wc[nn - 1] = (double)acc;
  }
  std::vector<double> gx, gw;
  gauss_lobatto(prm->gl_ord, gx, gw);
  auto push = [&](const double *src, size_t n) {
    size_t off = hp.blob.size();
    hp.blob.insert(hp.blob.end(), src, src + n);
    if (hp.blob.size() % 2) hp.blob.push_back(0.0);
    return off;
  };
  hp.blob.clear();
  hp.off_T = push(T.data(), T.size());
  hp.off_wc = push(wc.data(), wc.size());
  hp.off_glx = push(gx.data(), gx.size());
  hp.off_glw = push(gw.data(), gw.size());
  hp.off_j0z = push(prm->j0z, prm->n_j0z);
  hp.off_tp = push(prm->time_par, prm->n_time_par);
  double zero = 0.0;
  hp.off_mg = P.moench_M ? push(prm->moench_gamma, P.moench_M) : push(&zero, 1);
  return UNC_OK;
}

// ---- per-device context -------------------------------------------------------
struct DevBuf {
  void *ptr = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return UNC_OK;
    if (ptr) cudaFree(ptr);
    ptr = nullptr; cap = 0;
    size_t want = std::max(bytes, (size_t)4096);
    CK(cudaMalloc(&ptr, want));
    cap = want;
    return UNC_OK;
  }
  void release() { if (ptr) cudaFree(ptr); ptr = nullptr; cap = 0; }
};

struct DevCtx {
  bool init = false;
  cudaStream_t stream = nullptr;
  DevBuf tables, in, out, scratch, counter;
  int sm_count = 0;
  std::vector<double> blob_cached;
  bool smem_set[32] = {false};
};
DevCtx g_ctx[16];

int ensure_ctx(int dev) {
  if (dev < 0 || dev >= 16) return fail(UNC_ERR_BAD_ARG, "device %d out of range", dev);
  CK(cudaSetDevice(dev));
  DevCtx &c = g_ctx[dev];
  if (!c.init) {
    CK(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    CK(cudaDeviceGetAttribute(&c.sm_count, cudaDevAttrMultiProcessorCount, dev));
    c.init = true;
  }
  return UNC_OK;
}

int upload_tables(int dev, HostPlan &hp, cudaStream_t st) {
  DevCtx &c = g_ctx[dev];
  size_t bytes = hp.blob.size() * sizeof(double);
  int rc = c.tables.ensure(bytes);
  if (rc) return rc;
  if (c.blob_cached != hp.blob) {
    // tables are a few KB: synchronous semantics are fine, but stay on the job's stream
    CK(cudaMemcpyAsync(c.tables.ptr, hp.blob.data(), bytes, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    c.blob_cached = hp.blob;
  }
  const double *b = (const double *)c.tables.ptr;
  hp.P.ts_T = b + hp.off_T;
  hp.P.ts_wc = b + hp.off_wc;
  hp.P.gl_x = b + hp.off_glx;
  hp.P.gl_w = b + hp.off_glw;
  hp.P.j0z = b + hp.off_j0z;
  hp.P.time_par = b + hp.off_tp;
  hp.P.moench_gamma = b + hp.off_mg;
  return UNC_OK;
}

template <int ZT>
int launch_zt(int dev, const unc::DevParams &P, const unc::Job &J, cudaStream_t st) {
  const int na = P.nts_pad + P.gl_rounds * 32;
  const size_t smem = unc::smem_bytes(P.np, P.nacc, na, ZT);
  if (smem > 227 * 1024) return fail(UNC_ERR_UNSUPPORTED, "shared memory need %zu B exceeds 227 KB", smem);
  DevCtx &c = g_ctx[dev];
  if (!c.smem_set[ZT]) {
    CK(cudaFuncSetAttribute(unc::lh_point_kernel<ZT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            227 * 1024));
    c.smem_set[ZT] = true;
  }
  const long long ntiles = (J.nz + ZT - 1) / ZT;
  const long long nblk = J.ncol * ntiles;
  if (nblk <= 0) return UNC_OK;
  if (nblk > 2147483647LL) return fail(UNC_ERR_UNSUPPORTED, "too many work items (%lld)", nblk);
  unc::lh_point_kernel<ZT><<<(unsigned)nblk, UNC_THREADS, smem, st>>>(P, J);
  g_launches++;
  CK(cudaGetLastError());
  return UNC_OK;
}

template <int ZL>
int launch_grid_zl(int dev, const unc::DevParams &P, const unc::Job &J, cudaStream_t st) {
  const int NA = P.N + P.nacc * P.G;
  const size_t smem = unc::grid_smem_bytes(P.np, (NA + 31) & ~31, ZL);
  if (smem > 227 * 1024) return fail(UNC_ERR_UNSUPPORTED, "shared memory need %zu B exceeds 227 KB", smem);
  DevCtx &c = g_ctx[dev];
  if (!c.smem_set[10 + ZL]) {
    CK(cudaFuncSetAttribute(unc::lh_grid_kernel<ZL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    c.smem_set[10 + ZL] = true;
  }
  const long long nblk = J.ncol * ((J.nz + 32 * ZL - 1) / (32 * ZL));
  if (nblk <= 0) return UNC_OK;
  if (nblk > 2147483647LL) return fail(UNC_ERR_UNSUPPORTED, "too many work items (%lld)", nblk);
  unc::lh_grid_kernel<ZL><<<(unsigned)nblk, UNC_THREADS, smem, st>>>(P, J);
  g_launches++;
  CK(cudaGetLastError());
  return UNC_OK;
}

// 128 z per work item, persistent CTAs drawing items from an atomic counter; totlap in a
// per-CTA global scratch slot (kernels.cuh: lh_grid4_kernel)
template <int NW>
int launch_grid4_nw(int dev, const unc::DevParams &P, const unc::Job &J, cudaStream_t st) {
  const int NA = P.N + P.nacc * P.G;
  const size_t smem = unc::grid4_smem_bytes(P.np, (NA + 31) & ~31, NW);
  if (smem > 227 * 1024) return fail(UNC_ERR_UNSUPPORTED, "shared memory need %zu B exceeds 227 KB", smem);
  DevCtx &c = g_ctx[dev];
  if (!c.smem_set[NW % 9]) {
    CK(cudaFuncSetAttribute(unc::lh_grid4_kernel<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    c.smem_set[NW % 9] = true;
  }
  const long long nitems = J.ncol * ((J.nz + 127) / 128);
  if (nitems <= 0) return UNC_OK;
  if (nitems > 4000000000LL) return fail(UNC_ERR_UNSUPPORTED, "too many work items (%lld)", nitems);
  int occ = 2;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, unc::lh_grid4_kernel<NW>, NW * 32, smem));
  if (occ < 1) occ = 1;
  const int grid = (int)std::min<long long>(nitems, (long long)c.sm_count * occ);
  int rc = c.scratch.ensure((size_t)grid * 2 * P.np * 128 * sizeof(unc::cplx));   // two totlap slots per CTA
  if (rc) return rc;
  const bool fresh_counter = c.counter.ptr == nullptr;
  rc = c.counter.ensure(256);
  if (rc) return rc;
  if (fresh_counter) CK(cudaMemsetAsync(c.counter.ptr, 0, 256, st));   // the kernel re-arms it itself
  unc::lh_grid4_kernel<NW><<<grid, NW * 32, smem, st>>>(P, J, (unc::cplx *)c.scratch.ptr,
                                                         (unsigned int *)c.counter.ptr);
  g_launches++;
  CK(cudaGetLastError());
  return UNC_OK;
}

// second generation: eight z-slots per lane, two Laplace parameters per warp (lh_grid8_kernel)
template <int NW>
int launch_grid8_nw(int dev, const unc::DevParams &P, const unc::Job &J, cudaStream_t st) {
  const int NA = P.N + P.nacc * P.G;
  const size_t smem = unc::grid8_smem_bytes(P.np, (NA + 31) & ~31, NW);
  if (smem > 227 * 1024) return fail(UNC_ERR_UNSUPPORTED, "shared memory need %zu B exceeds 227 KB", smem);
  DevCtx &c = g_ctx[dev];
  if (!c.smem_set[20]) {
    CK(cudaFuncSetAttribute(unc::lh_grid8_kernel<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
#ifdef UNC_CARVEOUT   // measured (ms/step): default 111.7 = 60 (132 KB shared, 124 KB L1); 100 (28 KB L1) 115.8
    CK(cudaFuncSetAttribute(unc::lh_grid8_kernel<NW>, cudaFuncAttributePreferredSharedMemoryCarveout, UNC_CARVEOUT));
#endif
    c.smem_set[20] = true;
  }
  const long long nitems = J.ncol * ((J.nz + 127) / 128);
  if (nitems <= 0) return UNC_OK;
  if (nitems > 4000000000LL) return fail(UNC_ERR_UNSUPPORTED, "too many work items (%lld)", nitems);
  int occ = 2;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, unc::lh_grid8_kernel<NW>, NW * 32, smem));
  if (occ < 1) occ = 1;
  const int grid = (int)std::min<long long>(nitems, (long long)c.sm_count * occ);
  int rc = c.scratch.ensure((size_t)grid * 2 * P.np * 128 * sizeof(unc::cplx));   // two totlap slots per CTA
  if (rc) return rc;
  const bool fresh_counter = c.counter.ptr == nullptr;
  rc = c.counter.ensure(256);
  if (rc) return rc;
  if (fresh_counter) CK(cudaMemsetAsync(c.counter.ptr, 0, 256, st));   // the kernel re-arms it itself
  unc::lh_grid8_kernel<NW><<<grid, NW * 32, smem, st>>>(P, J, (unc::cplx *)c.scratch.ptr,
                                                         (unsigned int *)c.counter.ptr);
  g_launches++;
  CK(cudaGetLastError());
  return UNC_OK;
}

int launch_grid4(int dev, const unc::DevParams &P, const unc::Job &J, cudaStream_t st) {
#ifdef UNC_GRID4_NW
  return launch_grid4_nw<UNC_GRID4_NW>(dev, P, J, st);   // experiments (tools/run_variants.sh)
#else
  return launch_grid4_nw<8>(dev, P, J, st);
#endif
}

int launch_grid(int dev, const unc::DevParams &P, const unc::Job &J, cudaStream_t st) {
  const char *force = getenv("UNC_FORCE_KERNEL");
  const bool no4 = force && !strcmp(force, "grid2");
  const bool only4 = force && !strcmp(force, "grid4");
#ifdef UNC_GRID8_NW
  if (J.nz >= 96 && !no4 && !only4) return launch_grid8_nw<UNC_GRID8_NW>(dev, P, J, st);   // experiments
#else
  if (J.nz >= 96 && !no4 && !only4) return launch_grid8_nw<8>(dev, P, J, st);
#endif
  if (J.nz >= 96 && !no4) return launch_grid4(dev, P, J, st);
  // two z per lane (64 z per CTA) halves the per-(a,p) work per point; keep one z per lane
  // for short columns and when the larger totlap tile would not fit twice per SM
  if (J.nz > 32 && P.np <= 53) return launch_grid_zl<2>(dev, P, J, st);
  return launch_grid_zl<1>(dev, P, J, st);
}

// kernel selection: lanes<->z (grid kernel) once a column has enough z to fill most of a
// warp; otherwise lanes<->abscissae (point kernel).  UNC_FORCE_KERNEL=point|grid overrides.
int launch(int dev, const unc::DevParams &P, const unc::Job &J, cudaStream_t st) {
  const char *force = getenv("UNC_FORCE_KERNEL");
  bool grid = J.nz >= 12;
  // a small contour grid (fewer column CTAs than SMs) is a latency problem: the point kernel
  // spreads it over nz/4 times as many CTAs (hantush-contours deck, 30 r x 20 z: 6.9 ms -> <1 ms)
  if (grid && J.nz < 96 && J.ncol * ((J.nz + 31) / 32) < (long long)g_ctx[dev].sm_count) grid = false;
  if (force && !strcmp(force, "point")) grid = false;
  if (force && (!strcmp(force, "grid") || !strcmp(force, "grid2") || !strcmp(force, "grid4"))) grid = true;
  if (grid) return launch_grid(dev, P, J, st);
  if (J.nz >= 4) return launch_zt<4>(dev, P, J, st);
  if (J.nz >= 2) return launch_zt<2>(dev, P, J, st);
  return launch_zt<1>(dev, P, J, st);
}

int check_common(const double *tD, const int32_t *sv, const double *rD, const double *zD,
                 const int32_t *zLay, const double *o1, const double *o2) {
  if (!tD || !sv || !rD || !zD || !zLay || !o1 || !o2) return fail(UNC_ERR_BAD_ARG, "NULL array argument");
  return UNC_OK;
}

int validate_sv(const unc_params *prm, long long n, const int32_t *sv) {
  for (long long i = 0; i < n; ++i)
    if (sv[i] < 1 || sv[i] + prm->gl_nacc > prm->n_j0z)
      return fail(UNC_ERR_BAD_ARG, "sv[%lld]=%d needs j0z(1..%d) but n_j0z=%d", i, sv[i],
                  sv[i] + prm->gl_nacc, prm->n_j0z);
  return UNC_OK;
}

int device_count() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

// Host-array evaluation of a contiguous range of columns on one device.
struct Shard {
  int dev;
  long long c0, c1;  // columns [c0,c1)
  int rc = 0;
  std::string err;
};

struct HostJob {
  bool grid;
  long long ncol;
  int nt, nr, nz;
  const double *tD; const int32_t *sv; const double *rD; const double *zD; const int32_t *zLay;
  const double *ts;
  double *s, *ds; int32_t *flags;
};

int run_shard(const unc_params *prm, const HostJob &hj, Shard &sh) {
  HostPlan hp;
  int rc = make_plan(prm, hp);
  if (rc) return rc;
  rc = ensure_ctx(sh.dev);
  if (rc) return rc;
  DevCtx &c = g_ctx[sh.dev];
  cudaStream_t st = c.stream;
  rc = upload_tables(sh.dev, hp, st);
  if (rc) return rc;
  const long long nc = sh.c1 - sh.c0;
  if (nc <= 0) return UNC_OK;
  const int nz = hj.nz;
  // input layout on device: tD | rD | zD | ts | sv | zLay
  size_t n_t, n_r, n_z;
  long long t_first = 0, t_last = 0;
  if (hj.grid) {
    t_first = sh.c0 / hj.nr;
    t_last = (sh.c1 - 1) / hj.nr;
    n_t = (size_t)(t_last - t_first + 1);
    n_r = hj.nr;
    n_z = nz;
  } else {
    n_t = n_r = n_z = (size_t)nc;
  }
  const size_t n_ts = hj.ts ? (size_t)nc : 0;
  size_t bytes_in = (n_t + n_r + n_z + n_ts) * sizeof(double) + (n_t + n_z) * sizeof(int32_t) + 64;
  rc = c.in.ensure(bytes_in);
  if (rc) return rc;
  rc = c.out.ensure((size_t)nc * nz * (2 * sizeof(double) + sizeof(int32_t)) + 64);
  if (rc) return rc;
  double *d_tD = (double *)c.in.ptr;
  double *d_rD = d_tD + n_t;
  double *d_zD = d_rD + n_r;
  double *d_ts = d_zD + n_z;
  int32_t *d_sv = (int32_t *)(d_ts + n_ts);
  int32_t *d_lay = d_sv + n_t;
  double *d_s = (double *)c.out.ptr;
  double *d_ds = d_s + (size_t)nc * nz;
  int32_t *d_fl = (int32_t *)(d_ds + (size_t)nc * nz);
  const long long toff = hj.grid ? t_first : sh.c0;
  const long long roff = hj.grid ? 0 : sh.c0;
  const long long zoff = hj.grid ? 0 : sh.c0;
  CK(cudaMemcpyAsync(d_tD, hj.tD + toff, n_t * sizeof(double), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_sv, hj.sv + toff, n_t * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_rD, hj.rD + roff, n_r * sizeof(double), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_zD, hj.zD + zoff, n_z * sizeof(double), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_lay, hj.zLay + zoff, n_z * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  if (hj.ts) CK(cudaMemcpyAsync(d_ts, hj.ts + sh.c0, n_ts * sizeof(double), cudaMemcpyHostToDevice, st));
  unc::Job J;
  J.ncol = nc;
  J.nz = nz;
  J.tD = d_tD; J.sv = d_sv; J.rD = d_rD; J.zD = d_zD; J.zLay = d_lay;
  J.ts_scale = hj.ts ? d_ts : nullptr;
  J.s = d_s; J.ds = d_ds; J.flags = hj.flags ? d_fl : nullptr;
  if (hj.grid) {
    // local column lc = c - c0; global c = lc + c0: t = c / nr, r = c % nr.  Shift so the
    // kernel's (lc / tdiv, lc % rmod) addressing works: pad by starting at column
    // c0 - t_first*nr inside the first time row.
    J.tdiv = hj.nr;
    J.rmod = hj.nr;
    J.zstride = 0;
  } else {
    J.tdiv = 1;
    J.rmod = nc;
    J.zstride = 1;
  }
  if (hj.grid && (sh.c0 % hj.nr) != 0) {
    // shards are cut on time-row boundaries by the caller whenever possible; otherwise
    // run the partial rows one by one
    long long c = sh.c0;
    while (c < sh.c1) {
      long long row_end = std::min(sh.c1, (c / hj.nr + 1) * hj.nr);
      unc::Job Jr = J;
      Jr.ncol = row_end - c;
      Jr.tdiv = 1LL << 40;  // single time row
      Jr.tD = d_tD + (c / hj.nr - t_first);
      Jr.sv = d_sv + (c / hj.nr - t_first);
      Jr.rD = d_rD + (c % hj.nr);
      Jr.rmod = 1LL << 40;
      Jr.ts_scale = hj.ts ? d_ts + (c - sh.c0) : nullptr;
      Jr.s = d_s + (c - sh.c0) * nz;
      Jr.ds = d_ds + (c - sh.c0) * nz;
      Jr.flags = hj.flags ? d_fl + (c - sh.c0) * nz : nullptr;
      rc = launch(sh.dev, hp.P, Jr, st);
      if (rc) return rc;
      c = row_end;
    }
  } else {
    rc = launch(sh.dev, hp.P, J, st);
    if (rc) return rc;
  }
  CK(cudaMemcpyAsync(hj.s + sh.c0 * nz, d_s, (size_t)nc * nz * sizeof(double), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(hj.ds + sh.c0 * nz, d_ds, (size_t)nc * nz * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (hj.flags)
    CK(cudaMemcpyAsync(hj.flags + sh.c0 * nz, d_fl, (size_t)nc * nz * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return UNC_OK;
}

int run_host(const unc_params *prm, const HostJob &hj, int ngpu) {
  std::lock_guard<std::mutex> lk(g_mutex);
  {
    HostPlan probe;   // parameter validation is host-only and comes before any device work
    int rc = make_plan(prm, probe);
    if (rc) return rc;
  }
  if (ngpu < 0) return fail(UNC_ERR_BAD_ARG, "ngpu < 0");
  const int avail = device_count();
  if (avail <= 0) return fail(UNC_ERR_NO_DEVICE, "no CUDA device available (there is no CPU fallback)");
  int use = ngpu == 0 ? avail : std::min(ngpu, avail);
  if (hj.ncol < use) use = (int)std::max<long long>(1, hj.ncol);
  if (use == 1) {
    Shard sh{g_device, 0, hj.ncol};
    return run_shard(prm, hj, sh);
  }
  // contiguous equal split of the columns; grid shards cut on time rows
  // when there are enough rows, otherwise anywhere
  std::vector<Shard> shards(use);
  for (int g = 0; g < use; ++g) {
    long long c0 = hj.ncol * g / use, c1 = hj.ncol * (g + 1) / use;
    shards[g].dev = g;
    shards[g].c0 = c0;
    shards[g].c1 = c1;
  }
  std::vector<std::thread> th;
  for (int g = 0; g < use; ++g)
    th.emplace_back([&, g]() {
      shards[g].rc = run_shard(prm, hj, shards[g]);
      if (shards[g].rc) shards[g].err = g_err;
    });
  for (auto &t : th) t.join();
  for (int g = 0; g < use; ++g)
    if (shards[g].rc) {
      g_err = shards[g].err;
      return shards[g].rc;
    }
  return UNC_OK;
}

}  // namespace

extern "C" {

#ifdef UNC_PROFILE
int unc_debug_profile(unsigned long long *out, int reset) {
  unsigned long long z[16] = {0};
  if (out) cudaMemcpyFromSymbol(out, unc::g_prof, sizeof z);
  if (reset) cudaMemcpyToSymbol(unc::g_prof, z, sizeof z);
  return 0;
}
#endif

const char *unc_version(void) { return "unconfined_b200 0.1 (sm_100a)"; }
const char *unc_last_error(void) { return g_err.c_str(); }

int unc_device_count(int32_t *ngpu) {
  if (!ngpu) return fail(UNC_ERR_BAD_ARG, "NULL");
  *ngpu = device_count();
  return UNC_OK;
}

int unc_set_device(int32_t device) {
  if (device < 0 || device >= device_count()) return fail(UNC_ERR_NO_DEVICE, "device %d not available", device);
  g_device = device;
  return UNC_OK;
}

int unc_device_info(int32_t *ngpu, double *fp64_peak_flops) {
  int n = device_count();
  if (ngpu) *ngpu = n;
  if (fp64_peak_flops) {
    *fp64_peak_flops = 0.0;
    if (n > 0) {
      cudaDeviceProp pr;
      CK(cudaGetDeviceProperties(&pr, g_device));
      int khz = 0;
      CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, g_device));
      // 64 FP64 FMA lanes per SM per clock on sm_100 (2 flops each)
      *fp64_peak_flops = (double)pr.multiProcessorCount * 64.0 * 2.0 * (double)khz * 1e3;
    }
  }
  return UNC_OK;
}

int unc_kernel_launch_count(int64_t *n) {
  if (!n) return fail(UNC_ERR_BAD_ARG, "NULL");
  *n = g_launches.load();
  return UNC_OK;
}

int unc_measure_fp64_peak(double *flops) {
  if (!flops) return fail(UNC_ERR_BAD_ARG, "NULL");
  std::lock_guard<std::mutex> lk(g_mutex);
  if (device_count() <= 0) return fail(UNC_ERR_NO_DEVICE, "no CUDA device available");
  int rc = ensure_ctx(g_device);
  if (rc) return rc;
  cudaStream_t st = g_ctx[g_device].stream;
  cudaDeviceProp pr;
  CK(cudaGetDeviceProperties(&pr, g_device));
  const int blocks = pr.multiProcessorCount * 8, threads = 256, iters = 1 << 16;
  double *d = nullptr;
  CK(cudaMalloc(&d, (size_t)blocks * threads * sizeof(double)));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  unc::fp64_peak_kernel<<<blocks, threads, 0, st>>>(d, 1024);
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    CK(cudaEventRecord(e0, st));
    unc::fp64_peak_kernel<<<blocks, threads, 0, st>>>(d, iters);
    g_launches++;
    CK(cudaEventRecord(e1, st));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    double f = (double)blocks * threads * (double)iters * 8.0 * 2.0 / (ms * 1e-3);
    best = std::max(best, f);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  *flops = best;
  return UNC_OK;
}

int unc_shutdown(void) {
  std::lock_guard<std::mutex> lk(g_mutex);
  for (int d = 0; d < 16; ++d) {
    DevCtx &c = g_ctx[d];
    if (!c.init) continue;
    cudaSetDevice(d);
    c.tables.release(); c.in.release(); c.out.release(); c.scratch.release(); c.counter.release();
    cudaStreamDestroy(c.stream);
    c = DevCtx();
  }
  return UNC_OK;
}

int unc_eval_grid_ex(const unc_params *prm, int32_t nt, const double *tD, const int32_t *sv,
                     int32_t nr, const double *rD, int32_t nz, const double *zD,
                     const int32_t *zLay, const double *ts_abscissa_scale, int32_t ngpu,
                     double *totint, double *totintd, int32_t *flags) {
  if (nt < 0 || nr < 0 || nz < 0) return fail(UNC_ERR_BAD_ARG, "negative size");
  if (nt == 0 || nr == 0 || nz == 0) return UNC_OK;
  int rc = check_common(tD, sv, rD, zD, zLay, totint, totintd);
  if (rc) return rc;
  if (!prm) return fail(UNC_ERR_BAD_ARG, "prm is NULL");
  rc = validate_sv(prm, nt, sv);
  if (rc) return rc;
  HostJob hj{true, (long long)nt * nr, nt, nr, nz, tD, sv, rD, zD, zLay, ts_abscissa_scale,
             totint, totintd, flags};
  return run_host(prm, hj, ngpu);
}

int unc_eval_grid(const unc_params *prm, int32_t nt, const double *tD, const int32_t *sv,
                  int32_t nr, const double *rD, int32_t nz, const double *zD,
                  const int32_t *zLay, const double *ts_abscissa_scale, int32_t ngpu,
                  double *totint, double *totintd) {
  return unc_eval_grid_ex(prm, nt, tD, sv, nr, rD, nz, zD, zLay, ts_abscissa_scale, ngpu, totint,
                          totintd, nullptr);
}

int unc_eval_points_ex(const unc_params *prm, int64_t n, const double *tD, const int32_t *sv,
                       const double *rD, const double *zD, const int32_t *zLay,
                       const double *ts_abscissa_scale, int32_t ngpu, double *s, double *ds,
                       int32_t *flags) {
  if (n < 0) return fail(UNC_ERR_BAD_ARG, "negative size");
  if (n == 0) return UNC_OK;
  int rc = check_common(tD, sv, rD, zD, zLay, s, ds);
  if (rc) return rc;
  if (!prm) return fail(UNC_ERR_BAD_ARG, "prm is NULL");
  rc = validate_sv(prm, n, sv);
  if (rc) return rc;
  HostJob hj{false, (long long)n, 0, 0, 1, tD, sv, rD, zD, zLay, ts_abscissa_scale, s, ds, flags};
  return run_host(prm, hj, ngpu);
}

int unc_eval_points(const unc_params *prm, int64_t n, const double *tD, const int32_t *sv,
                    const double *rD, const double *zD, const int32_t *zLay,
                    const double *ts_abscissa_scale, int32_t ngpu, double *s, double *ds) {
  return unc_eval_points_ex(prm, n, tD, sv, rD, zD, zLay, ts_abscissa_scale, ngpu, s, ds, nullptr);
}

int unc_eval_grid_device(const unc_params *prm, int32_t nt, const double *d_tD,
                         const int32_t *d_sv, int32_t nr, const double *d_rD, int32_t nz,
                         const double *d_zD, const int32_t *d_zLay,
                         const double *d_ts_abscissa_scale, double *d_totint, double *d_totintd,
                         int32_t *d_flags, void *stream) {
  if (nt < 0 || nr < 0 || nz < 0) return fail(UNC_ERR_BAD_ARG, "negative size");
  if (nt == 0 || nr == 0 || nz == 0) return UNC_OK;
  int rc = check_common(d_tD, d_sv, d_rD, d_zD, d_zLay, d_totint, d_totintd);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(g_mutex);
  if (device_count() <= 0) return fail(UNC_ERR_NO_DEVICE, "no CUDA device available");
  HostPlan hp;
  rc = make_plan(prm, hp);
  if (rc) return rc;
  rc = ensure_ctx(g_device);
  if (rc) return rc;
  cudaStream_t st = stream ? (cudaStream_t)stream : (cudaStream_t)0;
  rc = upload_tables(g_device, hp, st);
  if (rc) return rc;
  unc::Job J;
  J.ncol = (long long)nt * nr;
  J.nz = nz;
  J.tdiv = nr; J.rmod = nr; J.zstride = 0;
  J.tD = d_tD; J.sv = d_sv; J.rD = d_rD; J.zD = d_zD; J.zLay = d_zLay;
  J.ts_scale = d_ts_abscissa_scale;
  J.s = d_totint; J.ds = d_totintd; J.flags = d_flags;
  return launch(g_device, hp.P, J, st);
}

int unc_eval_points_device(const unc_params *prm, int64_t n, const double *d_tD,
                           const int32_t *d_sv, const double *d_rD, const double *d_zD,
                           const int32_t *d_zLay, const double *d_ts_abscissa_scale, double *d_s,
                           double *d_ds, int32_t *d_flags, void *stream) {
  if (n < 0) return fail(UNC_ERR_BAD_ARG, "negative size");
  if (n == 0) return UNC_OK;
  int rc = check_common(d_tD, d_sv, d_rD, d_zD, d_zLay, d_s, d_ds);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(g_mutex);
  if (device_count() <= 0) return fail(UNC_ERR_NO_DEVICE, "no CUDA device available");
  HostPlan hp;
  rc = make_plan(prm, hp);
  if (rc) return rc;
  rc = ensure_ctx(g_device);
  if (rc) return rc;
  cudaStream_t st = stream ? (cudaStream_t)stream : (cudaStream_t)0;
  rc = upload_tables(g_device, hp, st);
  if (rc) return rc;
  unc::Job J;
  J.ncol = n;
  J.nz = 1;
  J.tdiv = 1; J.rmod = n; J.zstride = 1;
  J.tD = d_tD; J.sv = d_sv; J.rD = d_rD; J.zD = d_zD; J.zLay = d_zLay;
  J.ts_scale = d_ts_abscissa_scale;
  J.s = d_s; J.ds = d_ds; J.flags = d_flags;
  return launch(g_device, hp.P, J, st);
}

// driver_io.f90:628-647
int unc_j0_zeros(int32_t terms, double *j0z) {
  if (terms < 0 || (terms > 0 && !j0z)) return fail(UNC_ERR_BAD_ARG, "bad arguments");
  const double PIEP = 4.0 * std::atan(1.0);
  for (int i = 0; i < terms; ++i) {
    double x = (i + 0.75) * PIEP;
    for (int it = 0; it < 100; ++it) {
      double dx = ::j0(x) / ::j1(x);
      x = x + dx;
      double sp = std::nextafter(std::fabs(x), std::numeric_limits<double>::infinity()) - std::fabs(x);
      if (std::fabs(dx) < sp) break;
    }
    j0z[i] = x;
  }
  return UNC_OK;
}

// driver_io.f90:658-664
int unc_split_index(int32_t nt, const double *tD, int32_t j0s_a, int32_t j0s_b, int32_t *sv) {
  if (nt < 0 || (nt > 0 && (!tD || !sv))) return fail(UNC_ERR_BAD_ARG, "bad arguments");
  const int lo = std::min(j0s_a, j0s_b), zrange = std::max(j0s_a, j0s_b) - lo;
  double mn = INFINITY, mx = -INFINITY;
  for (int i = 0; i < nt; ++i) {
    double l = std::log10(tD[i]);
    mn = std::min(mn, l);
    mx = std::max(mx, l);
  }
  const int minlsp = (int)std::floor(mn), maxlsp = (int)std::ceil(mx);
  const int spRange = maxlsp - minlsp + 1;
  for (int i = 0; i < nt; ++i)
    sv[i] = lo + (int)(zrange * ((maxlsp - std::log10(tD[i])) / spRange));
  return UNC_OK;
}

// driver_io.f90:572-586
int unc_zlay(int32_t nz, const double *zD, double lD, double dD, int32_t *zLay) {
  if (nz < 0 || (nz > 0 && (!zD || !zLay))) return fail(UNC_ERR_BAD_ARG, "bad arguments");
  for (int i = 0; i < nz; ++i) {
    if (zD[i] <= 0.0 || zD[i] < (1.0 - lD)) zLay[i] = 1;
    else if ((zD[i] - 1.0) >= 0.0 || zD[i] < (1.0 - dD)) zLay[i] = 2;
    else zLay[i] = 3;
  }
  return UNC_OK;
}

}  // extern "C"
