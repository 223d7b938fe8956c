// Microbenchmark of the grid4 hot loop (eval4_recur over staged per-(a,p) terms) in isolation:
// how close to the FP64 pipe limit can this instruction stream run, as a function of
// warps/CTA, CTAs/SM and unrolling?  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include "../../unconfined_b200/csrc/kernels.cuh"

using namespace unc;

#ifndef UNROLL
#define UNROLL 1
#endif

constexpr int kUnroll = UNROLL;

#ifndef MODE
#define MODE 0
#endif
#ifndef BEXC
#define BEXC -1
#endif
#ifndef BK0Z
#define BK0Z false
#endif
template <int NW>
__global__ void __launch_bounds__(NW * 32, MINB) hot(const StageEnt4 *g_stage, double *out, int reps, DevParams P) {
  extern __shared__ __align__(16) unsigned char smem[];
  StageEnt4 *stage = (StageEnt4 *)smem + (threadIdx.x >> 5) * 32;
  const int lane = threadIdx.x & 31;
  stage[lane] = g_stage[lane];
  __syncwarp();
  cplx acc[4];
  for (int k = 0; k < 4; ++k) acc[k] = mk(0, 0);
  const double z0 = 0.001 * lane;
  const int l0 = reps < 0 ? 1 : 0, l1 = reps < -1 ? 1 : 0, l2 = reps < -2 ? 1 : 0, l3 = (lane >= 29 && reps > 0) ? 1 : 0;
  (void)l0; (void)l1; (void)l2; (void)l3;
#if MODE == 2
  // warp-specialised mix without hand-off: every fourth warp only computes per-(a,p) terms,
  // the others only run the hot loop (do the two instruction streams disturb each other?)
  if (((threadIdx.x >> 5) & 3) == 3) {
    double sink = 0.0;
    for (int r = 0; r < reps * 6; ++r) {
      StageEnt4 e;
      const cplx pp = mk(0.3 + 1e-6 * r, 0.7 + 0.01 * lane);
      bool ok = ap_terms_fast(P, pp, mk(0, 0), mk(0, 0), 1.0 + 0.37 * lane + 1e-3 * r, 1e-3, 3, 690.0, &e.eta, e.co);
      sink += ok ? e.co[0].cp.re + e.co[1].cm.im + e.eta.re : 0.0;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = sink;
    return;
  }
#endif
#if MODE == 3
  // eight z-slots per lane, two p per warp (half-warps read different stage entries)
  {
    StageEnt8 *st8 = (StageEnt8 *)smem + (threadIdx.x >> 5) * 32;
    {
      StageEnt8 t;
      const StageEnt4 &g = g_stage[lane];
      t.eta = g.eta; t.co[0] = g.co[0]; t.co[1] = g.co[1]; t.co[2] = g.co[2];
      t.sp = g.sp; t.sm = g.sm; t.spx = g.sp; t.smx = g.sm;
      st8[lane] = t;
    }
    __syncwarp();
    cplx a8[8];
    for (int k = 0; k < 8; ++k) a8[k] = mk(0, 0);
    const int half = lane >> 4;
    const int Lx = (lane & 15) >= 13 ? 1 : 0;
    for (int r = 0; r < reps; ++r) hot8_run<BEXC, BK0Z>(st8 + half * 16, 0, 16, z0, 0, Lx, a8);
    double t = 0; for (int k = 0; k < 8; ++k) t += a8[k].re + a8[k].im;
    out[blockIdx.x * blockDim.x + threadIdx.x] = t;
    return;
  }
#endif
  for (int r = 0; r < reps; ++r) {
#if MODE == 1
    {
      StageEnt4 e;
      const cplx pp = mk(0.3 + 1e-6 * r, 0.7 + 0.01 * lane);
      bool ok = ap_terms_fast(P, pp, mk(0, 0), mk(0, 0), 1.0 + 0.37 * lane + 1e-3 * r, 1e-3, 3, 690.0, &e.eta, e.co);
      const cbundle S = cexp_bundle(e.eta.re * 0.01, e.eta.im * 0.01);
      e.sp = S.ep; e.sm = S.em;
      if (ok) stage[lane] = e;
      __syncwarp();
    }
#endif
#pragma unroll kUnroll
    for (int j = 0; j < 32; ++j) {
      const StageEnt4 &e = stage[j];
#ifdef MIXED
      eval4_recur<0>(e, e.co[l0], e.co[l1], e.co[l2], e.co[l3], z0, acc);
#else
      const Coef c0 = e.co[0];
      eval4_recur<0>(e, c0, c0, c0, c0, z0, acc);
#endif
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc[0].re + acc[1].im + acc[2].re + acc[3].im;
}

int main(int argc, char **argv) {
  const int NW = NWARPS;
  int reps = argc > 1 ? atoi(argv[1]) : 2000;
  StageEnt4 h[32];
  for (int i = 0; i < 32; ++i) {
    h[i].eta = mk(1.0 + 0.1 * i, 0.3 + 0.01 * i);
    for (int L = 0; L < 3; ++L) { h[i].co[L].k0 = mk(1e-3, 2e-3); h[i].co[L].cp = mk(1e-4, -1e-4); h[i].co[L].cm = mk(2e-4, 1e-4); }
    h[i].sp = mk(1.0001, 1e-4); h[i].sm = mk(0.9999, -1e-4);
  }
  StageEnt4 *d; double *o;
  cudaMalloc(&d, sizeof h); cudaMemcpy(d, h, sizeof h, cudaMemcpyHostToDevice);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int grid = sms * MINB;
  cudaMalloc(&o, sizeof(double) * grid * NW * 32);
  size_t smem = NW * 32 * sizeof(StageEnt8) + PADSMEM;
  cudaFuncSetAttribute(hot<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  int occ; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, hot<NW>, NW * 32, smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  DevParams P; memset(&P, 0, sizeof P); P.model = 5; P.kappa = 0.5288; P.alphaD = 1e-5; P.beta = 2.0; P.lD = 0.018; P.dD = 8.4e-6; P.bD = P.lD - P.dD; P.lD1 = 1 - P.lD; P.dD1 = 1 - P.dD;
  hot<NW><<<grid, NW * 32, smem>>>(d, o, 10, P);
  cudaEventRecord(e0);
  hot<NW><<<grid, NW * 32, smem>>>(d, o, reps, P);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double absc = (double)grid * NW * reps * 32;            // warp-abscissae
  double fp64 = absc * 112 * 32;                          // FP64 thread-instructions
  printf("NW=%d MINB=%d UNROLL=%d occ=%d  %.2f ms  %.1f cycles/warp-abscissa/SMSP  FP64 pipe %.1f%% (112 instr/abscissa)\n",
         NW, MINB, UNROLL, occ, ms, ms * 1e-3 * 1.965e9 * sms * 4 / absc, 100 * fp64 / (ms * 1e-3) / (sms * 64.0 * 1.965e9));
  return 0;
}
