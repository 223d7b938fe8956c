// lh_grid8_kernel: the persistent kernel for contour grids (nz >= 33), the benchmarked kernel.
// Included at the end of kernels.cuh (it uses item_tables, ap_terms_stage8, hot8_chunk, slow8_run,
// finish8 and dehoog_lane from there).
#pragma once
namespace unc {

// Descriptor of a work item of lh_grid8_kernel (one (t,r) column x up to 128 z)
struct ItemMeta {
  long long col;
  double tD, Dz, eta_max;
  int z0, nzv, valid, lay_mask, zuni, pad_;
};

__host__ __device__ inline size_t grid8_smem_bytes(int np, int na_seq, int NW) {
  size_t b = 0;
  b += (size_t)NW * 32 * sizeof(StageEnt8);                                           // per-warp stage
  b += 3 * sizeof(ItemMeta) + 64;                                                     // descriptors, sync words
  return (b + 15) & ~(size_t)15;
}

// global scratch of one CTA: three totlap slots, two table slots, three stale-mask slots
__host__ __device__ inline size_t grid8_tab_bytes(int np, int na_seq) {
  return (size_t)4 * np * sizeof(cplx) + (size_t)2 * na_seq * sizeof(double);
}
__host__ __device__ inline size_t grid8_scratch_bytes(int np, int na_seq, int grid) {
  return (size_t)grid * ((size_t)3 * np * 128 * sizeof(cplx) + 2 * grid8_tab_bytes(np, na_seq) +
                         3 * 128 * sizeof(unsigned long long));
}

#ifndef UNC_GRID8_NW
#define UNC_GRID8_NW 16     // warps per CTA (one CTA per SM); see the measurements below
#endif
#ifndef UNC_GRID8_MINB
#define UNC_GRID8_MINB 1
#endif

// tables of the item in table slot `slot` (0/1) of lh_grid8_kernel's shared memory
__device__ __forceinline__ void grid8_tables(unsigned char *smem, int slot, int np, int na_seq, PTab &T,
                                             double *&a2, double *&wj) {
  unsigned char *q = smem + (size_t)slot * ((size_t)4 * np * sizeof(cplx) + (size_t)2 * na_seq * sizeof(double));
  T.p = (cplx *)q; q += np * sizeof(cplx);
  T.lt = (cplx *)q; q += np * sizeof(cplx);
  T.aux = (cplx *)q; q += np * sizeof(cplx);
  T.aux2 = (cplx *)q; q += np * sizeof(cplx);
  a2 = (double *)q; q += na_seq * sizeof(double);
  wj = (double *)q;
}

// wait until a CTA-level counter in shared memory reaches a target (polled by lane 0)
__device__ __forceinline__ void wait_ge(volatile int *w, int target, int lane) {
  (void)lane;
  while (*w < target) __nanosleep(40);   // every lane polls the same word: the warp stays converged
  __syncwarp();
  __threadfence_block();
}

// T(kn) of lh_grid8_kernel: fetch item kn from the global counter and build its tables and
// descriptor; executed by one warp, kept out of line (once per item).
template <int ZL, int ZB, int GL, int NDMAX>
__device__ __noinline__ void grid8_tjob(const DevParams &P, const Job &J, int kn, long long nitems, int nzb, int np2,
                                        unsigned char *tabs, unsigned long long *s_flag, ItemMeta *s_meta,
                                        volatile int *s_sync, unsigned int *g_counter, int lane) {
  {
    volatile int *s_ready = s_sync + 1, *s_pdone = s_sync + 4, *s_ddone = s_sync + 6;
    const int hl = lane & 15;
    const int np = P.np;
    const int na_seq = (P.N + P.nacc * P.G + 31) & ~31;
    const int ms = kn % 3, ts = kn & 1;
    bool valid = (kn == 0) ? true : (s_meta[(kn - 1) % 3].valid != 0);
    // the descriptor / stale-mask slot still belongs to item kn-3 until its inversions are done
    if (kn >= 3 && s_meta[ms].valid) wait_ge(&s_ddone[ms], (kn / 3) * NDMAX, lane);
    long long item = 0;
    if (valid) {
      unsigned int it = 0;
      if (lane == 0) it = atomicAdd(g_counter, 1u);
      it = __shfl_sync(0xffffffffu, it, 0);
      item = it;
      valid = item < nitems;
    }
    if (valid) {
      wait_ge(&s_pdone[ts], (kn / 2) * np2, lane);   // p-jobs of item kn-2 no longer read the table slot
      const long long col = item / nzb;
      const int z0 = (int)(item % nzb) * ZB;
      const int nzv = min(ZB, J.nz - z0);
      const long long tcol = col + J.col0;
      const double tD = J.tD[tcol / J.tdiv];
      const int sv = J.sv[tcol / J.tdiv];
      const double rD = J.rD[tcol % J.rmod];
      const double arg = P.j0z[sv - 1] / rD;                      // driver.f90:120
      const double tscale = J.ts_scale ? J.ts_scale[col] : arg;  // driver.f90:121-126
      const long long zbase = (J.zstride ? col * (long long)J.nz : 0) + z0;
      for (int i = lane; i < ZB; i += 32) __stcg(&s_flag[ms * 128 + i], 0ull);
      PTab T;
      double *a2, *wj;
      grid8_tables(tabs, ts, np, na_seq, T, a2, wj);
      item_tables(P, T, a2, wj, tD, sv, rD, tscale, lane, 32);
      double myz[ZL];
      int mylay[ZL];
      bool zvalid[ZL];
#pragma unroll
      for (int k = 0; k < ZL; ++k) {
        const int zi = hl + GL * k;
        zvalid[k] = zi < nzv;
        myz[k] = zvalid[k] ? J.zD[zbase + zi] : 0.0;
        mylay[k] = zvalid[k] ? J.zLay[zbase + zi] : 0;
      }
      int mm = 0;
      float za = 0.f;
#pragma unroll
      for (int k = 0; k < ZL; ++k)
        if (zvalid[k]) { mm |= 1 << (mylay[k] - 1); za = fmaxf(za, (float)fabs(myz[k]) * 1.0000002f); }
      for (int o = 16; o > 0; o >>= 1) {
        mm |= __shfl_xor_sync(0xffffffffu, mm, o);
        za = fmaxf(za, __shfl_xor_sync(0xffffffffu, za, o));
      }
      // equally spaced slots?  D from lane 0 (slots 0,1 are always valid when nz >= 32)
      const double D = __shfl_sync(0xffffffffu, myz[1] - myz[0], 0);
      const double tol = 4.0 * 2.220446049250313e-16 * (double)za;
      bool uni = __shfl_sync(0xffffffffu, (int)(zvalid[0] && zvalid[1]), 0) != 0;
#pragma unroll
      for (int k = 0; k + 1 < ZL; ++k)
        if (zvalid[k] && zvalid[k + 1] && !(fabs((myz[k + 1] - myz[k]) - D) <= tol)) uni = false;
      uni = __all_sync(0xffffffffu, uni);
      if (lane == 0) {
        ItemMeta d;
        d.col = col; d.tD = tD; d.Dz = D; d.eta_max = fast_eta_max(P, mm, (double)za);
        d.z0 = z0; d.nzv = nzv; d.valid = 1; d.lay_mask = mm; d.zuni = uni ? 1 : 0; d.pad_ = 0;
        s_meta[ms] = d;
      }
    } else if (lane == 0) {
      s_meta[ms].valid = 0;
      if (s_sync[9] < 0) s_sync[9] = kn;   // first round without an item
    }
    __syncwarp();
    __threadfence_block();
    if (lane == 0) s_ready[ms] = kn + 1;
  }
}

// ---------------------------------------------------------------------------
// Grid kernel for contour grids (nz >= 33): persistent CTAs (one 16-warp CTA per SM) draw work
// items (one (t,r) column x up to 128 z) from a global atomic counter, so the few expensive
// columns (literal path at small rD) do not leave SMs idle.  Lanes <-> z with EIGHT z-slots per
// lane and TWO Laplace parameters per warp (lanes 0-15 <-> p = 2*job, lanes 16-31 <-> p =
// 2*job+1; z = z0 + hl + 16 k): the per-(a,p) terms are staged in shared memory 16 abscissae at
// a time and shared by 128 z, the per-abscissa exponential of slot 0 (48 of the FP64
// instructions) by eight z, and the slots advance the products cp*e^{eta z}, cm*e^{-eta z}
// themselves (eval8_scaled).  Requires equally spaced z within an item, checked per item on the
// actual z (tolerance 4 ulp of max|z|: the induced error |eta|*4ulp is the size of the
// reference's own rounding of the product eta*zD); any other z-list takes the exact per-slot
// evaluation.  totlap (np x 128 complex = 108 KB for M=26) and the item's tables (p, lapTime,
// a^2, weight*a*J0: 15 KB, read once per (a,p)) live in global scratch slots owned by the CTA
// (L1/L2-resident): shared memory holds only the per-warp stage, which leaves 124 KB of L1 to
// the thread-local area and q-d arrays.
//
// Job flow without CTA-wide barriers.  The warps of a CTA claim jobs from ONE linear sequence
// (a shared-memory counter); "round" k of the sequence holds, in this order,
//   P(k,j)   j < ceil(np/2): quadrature + Wynn of two Laplace parameters of item k,
//   T(k+1)   fetch item k+1 from the global counter and build its tables (one warp),
//   D(k-1,j) j < 8: de Hoog inversions (value and derivative) of 32 (z,kind) pairs of item k-1.
// Dependencies are only on EARLIER jobs of the sequence, which some warp has already claimed
// and runs to completion, so the waits below cannot deadlock:
//   P(k,.)   needs T(k) (tables, descriptor: `ready`) and D(k-3,.) done (totlap slot k%3 free);
//   T(k+1)   needs P(k-1,.) done (table slot (k+1)%2 free) and D(k-2,.) done (descriptor and
//            stale-mask slot (k+1)%3 free);
//   D(k-1,.) needs P(k-1,.) done.
// Tables are double-buffered; totlap, descriptors and stale masks triple-buffered, so that no
// job waits on a job of the round just before it.  A warp that runs out of p-jobs of item k
// goes on with the tables of item k+1, the inversions of item k-1 and then the p-jobs of item
// k+1 while the others finish.
//
// Measured on C5a (ms per 2^20-point step, same box, round 2), against the former design (two
// 8-warp CTAs per SM, one item per CTA-wide barrier round, tables in shared memory): 91.8.
//   this job flow, two 8-warp CTAs, tables double-buffered in SHARED memory   98.7
//     (91 KB of shared memory per CTA leave 60 KB of L1 instead of 92 KB: the L1 share of the
//      thread-local arrays is what decides here, not the barrier)
//   the same with the tables in global scratch                               91.4
//   one CTA per SM with 12 / 14 / 16 / 18 warps                   99.6 / 95.9 / 89.3 / 90.0
// and on a 1024-column shard (one GPU's share of an 8-GPU split) 11.47 against 11.99 ms: with 16
// warps on one item the last wave of a launch is half as long.  Polling is done by every lane
// (see wait_ge): with lane 0 polling alone the warps came back split and ran 16 % more warp
// instructions.
template <int NW>
__global__ void __launch_bounds__(NW * 32, UNC_GRID8_MINB)
lh_grid8_kernel(const __grid_constant__ DevParams P, const __grid_constant__ Job J,
                cplx *__restrict__ g_tot, unsigned int *__restrict__ g_counter) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int ZL = 8, ZB = 128, GL = 16;   // 8 slots per lane, 16 lanes per Laplace parameter
  constexpr int NDMAX = 2 * ZB / 32;          // D-job slots per item
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int half = lane >> 4, hl = lane & 15;   // which of the warp's two p; z lane
  const int np = P.np, nacc = P.nacc, G = P.G, N = P.N;
  const int NA = N + nacc * G;
  const int na_seq = (NA + 31) & ~31;
  const int nzb = (J.nz + ZB - 1) / ZB;
  const long long nitems = J.ncol * (long long)nzb;
  const int np2 = (np + 1) / 2;               // p-jobs per item (two Laplace parameters each)
  const int JPR = np2 + 1 + NDMAX;            // jobs per round

  // the items' tables (p, lapTime, a^2, weights: 15 KB, read once per (a,p)) live in a global
  // scratch slot of the CTA (L1/L2-resident), two items deep: shared memory is left to the stage
  const size_t tab_bytes = grid8_tab_bytes(np, na_seq);
  unsigned char *tabs = (unsigned char *)(g_tot + (size_t)gridDim.x * 3 * np * ZB) + (size_t)blockIdx.x * 2 * tab_bytes;
  // stale masks (one 64-bit mask of p per z, three items deep): rarely touched, kept in the CTA's
  // global scratch behind the tables (written with atomics, read with L2 loads)
  unsigned long long *s_flag = (unsigned long long *)((unsigned char *)(g_tot + (size_t)gridDim.x * 3 * np * ZB) +
                                                      (size_t)gridDim.x * 2 * tab_bytes) + (size_t)blockIdx.x * 3 * 128;
  unsigned char *sp = smem_raw;
  cplx *s_stage = (cplx *)sp; sp += (size_t)NW * 32 * sizeof(StageEnt8);   // [warp][st8_idx(field, abscissa, half)]
  ItemMeta *s_meta = (ItemMeta *)sp; sp += 3 * sizeof(ItemMeta);
  volatile int *s_sync = (volatile int *)sp;
  // s_sync: [0] next job, [1..3] ready (round + 1 of the descriptor in slot r%3), [4..5] p-jobs done per
  // table slot (cumulative), [6..7] D-job slots done per totlap slot (cumulative), [8] first round
  // without an item (-1 while unknown), [9] warps that have left
  volatile int *s_ready = s_sync + 1, *s_pdone = s_sync + 4, *s_ddone = s_sync + 6;
  cplx *tot_base = g_tot + (size_t)blockIdx.x * 3 * np * ZB;  // this CTA's three totlap slots [p][z]

  if (tid < 16) s_sync[tid] = (tid == 9) ? -1 : 0;
  if (tid < 3) s_meta[tid].valid = 0;
  __syncthreads();

  cplx *stage = s_stage + (size_t)warp * 32 * ST8_NF + st8_idx(0, 0, half);   // field 0 of this half-warp's abscissa 0

  if (warp == 0) grid8_tjob<ZL, ZB, GL, NDMAX>(P, J, 0, nitems, nzb, np2, tabs, s_flag, s_meta, s_sync, g_counter, lane);
  __syncthreads();

  for (;;) {
    int g = 0;
    if (lane == 0) g = atomicAdd((int *)&s_sync[0], 1);
    __syncwarp();
    g = __shfl_sync(0xffffffffu, g, 0);
    const int k = g / JPR, j = g - k * JPR;
    // the descriptor of round k, or the knowledge that the items ended before it
    {
      bool gone = false;
      for (;;) {          // polled by every lane alike (see wait_ge)
        const int e = s_sync[9];
        if (e >= 0 && k > e) { gone = true; break; }
        if (s_ready[k % 3] >= k + 1) break;
        __nanosleep(40);
      }
      gone = __any_sync(0xffffffffu, gone);
      __threadfence_block();
      if (gone) break;
    }
    if (j == np2) {   // ---- T(k+1) ----------------------------------------------------------
      grid8_tjob<ZL, ZB, GL, NDMAX>(P, J, k + 1, nitems, nzb, np2, tabs, s_flag, s_meta, s_sync, g_counter, lane);
      continue;
    }
    if (j > np2) {
      // ---- D-job: de Hoog for 32 (z, value|derivative) pairs of item k-1 --------------------
      const int kk = k - 1;
      if (kk < 0) continue;
      const ItemMeta pm = s_meta[kk % 3];
      if (!pm.valid) continue;
      wait_ge(&s_pdone[kk & 1], (kk / 2 + 1) * np2, lane);
      const cplx *tot_prev = tot_base + (size_t)(kk % 3) * np * ZB;
      const unsigned long long *flag_prev = s_flag + (kk % 3) * 128;
      const int idx = (j - np2 - 1) * 32 + lane;
      if (idx < 2 * pm.nzv) {
        const int deriv = idx >= pm.nzv ? 1 : 0;
        const int zi = idx - deriv * pm.nzv;
        const unsigned long long fl = __ldcg(&flag_prev[zi]);
        const double ptee = P.tee_mult * pm.tD;
#ifdef UNC_SKIP_DEHOOG
        double v = tot_prev[zi].re;
#else
        double v = dehoog_lane(P, tot_prev + zi, ZB, deriv != 0, pm.tD, ptee);
#endif
        const long long o = pm.col * (long long)J.nz + pm.z0 + zi;
        if (deriv) J.ds[o] = v * pm.tD;  // driver.f90:228
        else {
          J.s[o] = v;
          if (J.flags) J.flags[o] = fl != 0ull ? 1 : 0;
          if (J.smask) J.smask[o] = fl;
          if (J.nstale && fl != 0ull) atomicAdd(J.nstale, 1u);
        }
      }
      __syncwarp();
      if (lane == 0) atomicAdd((int *)&s_ddone[kk % 3], 1);
      continue;
    }
    // ---- p-job: Hankel quadrature + Wynn for two Laplace parameters (one per half-warp),
    //      128 z each.  np odd: the upper half of the last job repeats p = np-1 and stores nothing.
    const ItemMeta im = s_meta[k % 3];
    if (!im.valid) continue;
    if (k >= 3) wait_ge(&s_ddone[k % 3], (k / 3) * NDMAX, lane);   // totlap slot: item k-3 inverted
    PTab T;
    double *s_a2, *s_wj;
    grid8_tables(tabs, k & 1, np, na_seq, T, s_a2, s_wj);
    cplx *tot = tot_base + (size_t)(k % 3) * np * ZB;
    unsigned long long *flag_cur = s_flag + (k % 3) * 128;
    const int nzv = im.nzv, lay_mask = im.lay_mask;
    const double eta_max = im.eta_max, Dz = im.Dz;
    const bool zuni = im.zuni != 0;
    const long long zbase = (J.zstride ? im.col * (long long)J.nz : 0) + im.z0;
    const int L0 = __ffs(lay_mask) - 1;
    double z_first;
    int Lc, kx, Lx, Lb = 0;
    // slots of this lane beyond the column's last z: their results are dropped, so the early stop
    // below must not wait for them (they continue the grid past the top and mostly overflow; a
    // column of 32..127 z, or the last block of a longer one, otherwise never stops early:
    // 56 instead of 23 ms for 2048 columns of 48 z)
    int padmask = 0;
    bool hot_ok, k0z;
    {
      double myz[ZL];
      int myL[ZL];
#pragma unroll
      for (int kq = 0; kq < ZL; ++kq) {
        const int zi = hl + GL * kq;
        const bool v = zi < nzv;
        myz[kq] = v ? J.zD[zbase + zi] : 0.0;
        myL[kq] = (v ? J.zLay[zbase + zi] : L0 + 1) - 1;   // padding slots mimic a present layer
      }
      const bool v0 = hl < nzv;
#pragma unroll
      for (int kq = 0; kq < ZL; ++kq)
        if (!(hl + GL * kq < nzv)) {                          // ... and the uniform grid
          myz[kq] = zuni ? myz[0] + kq * Dz : 0.5;
          if (!v0) myz[kq] = 0.5;
          padmask |= 1 << kq;
        }
      z_first = myz[0];
      // slots whose lanes are not all on the layer of (slot 0, lane 0); exactly one such slot
      // gets the cheaper "exception" loop
      Lc = __shfl_sync(0xffffffffu, myL[0], 0);
      int offmask = 0;
#pragma unroll
      for (int kq = 0; kq < ZL; ++kq)
        if (!__all_sync(0xffffffffu, myL[kq] == Lc)) offmask |= 1 << kq;
      kx = (offmask != 0 && (offmask & (offmask - 1)) == 0) ? __ffs(offmask) - 1 : -1;
      hot_ok = (offmask & (offmask - 1)) == 0;   // at most one slot off the common layer
      Lx = myL[0];
#pragma unroll
      for (int kq = 1; kq < ZL; ++kq) if (kq == kx) Lx = myL[kq];
      // ... whose lanes off the common layer all lie on ONE other layer Lb (padding lanes, whose
      // results are dropped, join the common layer)
      Lb = Lc;
      if (kx >= 0) {
        if (!(hl + GL * kx < nzv)) Lx = Lc;
        const unsigned offl = __ballot_sync(0xffffffffu, Lx != Lc);
        if (offl) Lb = __shfl_sync(0xffffffffu, Lx, __ffs(offl) - 1);
        if (!__all_sync(0xffffffffu, Lx == Lc || Lx == Lb)) hot_ok = false;
      }
      // k0 of the common layer is exactly zero below/above the screen of the Hantush-type models
      k0z = (P.model == 1 || P.model == 2 || P.model == 3 || P.model == 5) && Lc != 1;
    }
    const bool pvalid = 2 * j + half < np;
    const int pi = min(2 * j + half, np - 1);
    int stale = 0;
    const cplx pp = T.p[pi], aux = T.aux[pi], aux2 = T.aux2[pi];
    // areas[k][0] = tanh-sinh part (finint), areas[k][1..nacc] = Gauss-Lobatto interval areas:
    // one thread-local array instead of a second register set for the finite part
    cplx areas[ZL][UNC_MAX_NACC + 1];
    cplx acc[ZL];
#pragma unroll
    for (int kq = 0; kq < ZL; ++kq) { acc[kq] = mk(0.0, 0.0); areas[kq][0] = mk(0.0, 0.0); }
    int seg = 0;
    int next_b = N;
    // Wynn only uses the areas before the first non-finite one (integration.f90:140-160) and
    // driver.f90:209 only asks whether SOME area is finite and non-zero.  Once that is settled
    // for every z of the warp (dead: a non-finite area seen; anyf: a finite non-zero one seen)
    // the remaining, ever more expensive, overflowing abscissae cannot change the result.
    const cplx lt_chk = T.lt[pi];
    const bool lt_ok = is_finite_fastc(lt_chk) && (lt_chk.re != 0.0 || lt_chk.im != 0.0);
    int dead = 0, anyf = 0;
    bool done = false;
    for (int base = 0; base < NA && !done; base += GL) {
      int ok = 1;
      {
        const int idx = base + hl;
        if (idx < NA)
          ok = ap_terms_stage8(P, pp, aux, aux2, s_a2[idx], s_wj[idx], lay_mask, eta_max, zuni, Dz, kx, Lb,
                               stage + hl * ST8_JSTEP);
      }
      const unsigned okball = __ballot_sync(0xffffffffu, ok);
      const bool all_ok = okball == 0xffffffffu;
      __syncwarp();
      const int cnt = min(GL, NA - base);
      int jj = 0;
      while (jj < cnt) {
        const int jend = min(cnt, next_b - base);
        if (all_ok && zuni && hot_ok) {
          // hot loop over the whole chunk (segment ends are handled inside the call)
          const int seg0 = seg;
#ifndef UNC_SKIP_HOT
#define UNC_H8(KXV, KZV) seg = hot8_chunk<KXV, KZV>(stage, cnt, z_first, Lc, Lx, acc, &areas[0][0], seg, next_b - base, NA - base, G)
          if (kx < 0) {
            if (k0z) UNC_H8(-1, true); else UNC_H8(-1, false);
          } else if (k0z) {
            switch (kx) {
              case 0: UNC_H8(0, true); break;
              case 1: UNC_H8(1, true); break;
              case 2: UNC_H8(2, true); break;
              case 3: UNC_H8(3, true); break;
              case 4: UNC_H8(4, true); break;
              case 5: UNC_H8(5, true); break;
              case 6: UNC_H8(6, true); break;
              default: UNC_H8(7, true); break;
            }
          } else {
            switch (kx) {
              case 0: UNC_H8(0, false); break;
              case 1: UNC_H8(1, false); break;
              case 2: UNC_H8(2, false); break;
              case 3: UNC_H8(3, false); break;
              case 4: UNC_H8(4, false); break;
              case 5: UNC_H8(5, false); break;
              case 6: UNC_H8(6, false); break;
              default: UNC_H8(7, false); break;
            }
          }
#undef UNC_H8
#else
          while (next_b - base <= cnt && next_b < NA) { seg += 1; next_b += G; }
          next_b -= (seg - seg0) * G;
#endif
          const int hret = seg;
          seg = hret & 0xff;
          next_b += (seg - seg0) * G;
          jj = cnt;
          // fate of the intervals closed inside this chunk (see the comment at `dead`)
          if (lt_ok) {
            for (int sidx = max(seg0, 1); sidx < seg && !done; ++sidx) {
              int cur_bad = 0;
              if (((hret >> 8) & 3) == 1) {       // one interval closed: its fate came back in registers
                cur_bad = (hret >> 12) & 0xff;
                anyf |= (hret >> 20) & 0xff;
              } else
              {
#pragma unroll
                for (int kq = 0; kq < ZL; ++kq) {
                  const cplx a = areas[kq][sidx];
                  const bool f = is_finite_fastc(a);
                  if (!f) cur_bad |= 1 << kq;
                  if (f && (a.re != 0.0 || a.im != 0.0)) anyf |= 1 << kq;
                }
              }
              dead |= cur_bad;
              if (__all_sync(0xffffffffu, ((dead & anyf) | padmask) == (1 << ZL) - 1)) {
                const double nanv = __longlong_as_double(0x7ff8000000000000LL);
#pragma unroll
                for (int kq = 0; kq < ZL; ++kq) {
                  for (int q = sidx + 1; q <= nacc; ++q) areas[kq][q] = mk(nanv, nanv);
                  acc[kq] = mk(nanv, nanv);
                }
                seg = nacc;   // the final store below rewrites areas[nacc] with NaN
                done = true;
              }
            }
          }
          continue;
        } else {
          // rare: abscissae beyond the fast-path bound, z-lists that are not equally spaced,
          // more than one slot off the common layer -- kept out of line (and re-reading its
          // z from global memory) so that it does not weigh on the registers of the common path
          slow8_run(P, T, pi, stage, (okball >> (half * GL)) & 0xffffu, base, jj, jend, s_wj, s_a2, J.zD + zbase, J.zLay + zbase, nzv, hl,
                    L0, zuni, Dz, acc);
          jj = jend;
        }
        const bool seg_end = (base + jj == next_b && next_b < NA);
        if (seg >= 1 && lt_ok && (seg_end || !all_ok)) {
          // at an interval end: record its fate; inside an interval that already went
          // non-finite for everybody (only looked at after chunks with literal nodes): stop
          int cur_bad = 0;
#pragma unroll
          for (int kq = 0; kq < ZL; ++kq) {
            const bool f = is_finite_fastc(acc[kq]);
            if (!f) cur_bad |= 1 << kq;
            if (seg_end && f && (acc[kq].re != 0.0 || acc[kq].im != 0.0)) anyf |= 1 << kq;
          }
          if (seg_end) dead |= cur_bad;
          const int settled = (dead | cur_bad) & anyf;
          if (__all_sync(0xffffffffu, (settled | padmask) == (1 << ZL) - 1)) {
            const double nanv = __longlong_as_double(0x7ff8000000000000LL);
#pragma unroll
            for (int kq = 0; kq < ZL; ++kq) {
              if (seg_end) areas[kq][seg] = acc[kq];
              for (int q = seg_end ? seg : seg - 1; q < nacc; ++q) areas[kq][q + 1] = mk(nanv, nanv);
              acc[kq] = mk(nanv, nanv);
            }
            seg = nacc;   // the final store below rewrites series[nacc-1] with NaN
            done = true;
            break;
          }
        }
        if (seg_end) {
#pragma unroll
          for (int kq = 0; kq < ZL; ++kq) {
            areas[kq][seg] = acc[kq];
            acc[kq] = mk(0.0, 0.0);
          }
          seg += 1;
          next_b += G;
        }
      }
      __syncwarp();
    }
#pragma unroll
    for (int kq = 0; kq < ZL; ++kq) areas[kq][seg] = acc[kq];
    // lapTime, Wynn-epsilon on the interval areas, totlap = finint + infint for the 8 slots
#ifndef UNC_WYNN_LOCALMEM
    stale = finish8(&areas[0][0], nacc, T.lt[pi], pvalid ? tot + (size_t)pi * ZB + hl : nullptr,
                    s_stage + (size_t)warp * 32 * ST8_NF + lane);
    __syncwarp();   // the scratch becomes the stage of the next job again
#else
    stale = finish8(&areas[0][0], nacc, T.lt[pi], pvalid ? tot + (size_t)pi * ZB + hl : nullptr, nullptr);
#endif
#pragma unroll
    for (int kq = 0; kq < ZL; ++kq) if (pvalid && (stale & (1 << kq))) atomicOr(&flag_cur[GL * kq + hl], 1ull << pi);
    __threadfence_block();   // totlap rows and stale masks before the completion count
    __syncwarp();
    if (lane == 0) atomicAdd((int *)&s_pdone[k & 1], 1);
  }
  // the last warp of the last CTA re-arms both global counters, so that a profiler's replay of the
  // launch starts from zero as well (the host also clears them before every launch)
  if (lane == 0 && atomicAdd((int *)&s_sync[10], 1) == NW - 1) {
    __threadfence();
    if (atomicAdd(g_counter + 1, 1u) == gridDim.x - 1) {
      g_counter[0] = 0u;
      g_counter[1] = 0u;
      __threadfence();
    }
  }
}



}  // namespace unc
