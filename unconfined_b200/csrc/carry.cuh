// Stale-infint carry (driver.f90:205-214): when every Gauss-Lobatto area of a (p,z) is zero or
// NaN the reference does not assign infint(p,z), which therefore keeps the value of the last
// (t,r) iteration that did assign it (t outer, r inner; the array is allocated once,
// driver.f90:66, and never initialised: 0 is used before the first assignment).  The main
// kernels write, per output point, the 64-bit mask of its stale p; these helper kernels turn
// the masks into work lists for the two fix-up passes of lh_point_kernel<1> (Job::fix_mode):
//   sources       (column c', z) whose Wynn result some later point inherits, for the p wanted;
//   destinations  every point with a non-zero mask, re-inverted with the inherited values.
#pragma once
#include <cstdint>

namespace unc {

// list[k] = index of the k-th point (arbitrary order) with mask != 0; slot[idx] = k or -1
__global__ void carry_list_kernel(const unsigned long long *__restrict__ mask, long long npts,
                                  int *__restrict__ list, int *__restrict__ slot,
                                  unsigned int *__restrict__ count) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= npts) return;
  int k = -1;
  if (mask[idx] != 0ull) {
    k = (int)atomicAdd(count, 1u);
    if (list) list[k] = (int)idx;
  }
  if (slot) slot[idx] = k;
}

// One thread per (z,p) walks the columns in the reference's (t,r) order: for a stale (c,z,p)
// the source is the last earlier column where (z,p) was assigned (src_col, -1 = none), and
// that column is marked as needed for p.
__global__ void carry_scan_kernel(const unsigned long long *__restrict__ mask, long long ncol, int nz,
                                  int np, const int *__restrict__ slot, int *__restrict__ src_col,
                                  unsigned long long *__restrict__ need) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nz * np) return;
  const int z = t % nz, p = t / nz;
  long long last = -1;
  for (long long c = 0; c < ncol; ++c) {
    const unsigned long long m = mask[c * nz + z];
    if ((m >> p) & 1ull) {
      src_col[(size_t)slot[c * nz + z] * np + p] = (int)last;
      if (last >= 0) atomicOr(&need[last * nz + z], 1ull << p);
    } else {
      last = c;
    }
  }
}

// src_col (column index) -> entry index in the source list; entries of non-stale p are -1
__global__ void carry_link_kernel(const int *__restrict__ list, long long nf, int nz, int np,
                                  const unsigned long long *__restrict__ mask,
                                  const int *__restrict__ src_slot, int *__restrict__ src) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nf * np) return;
  const long long e = i / np;
  const int p = (int)(i - e * np);
  const long long pt = list[e];
  const int z = (int)(pt % nz);
  int v = -1;
  if ((mask[pt] >> p) & 1ull) {
    const int c = src[i];
    if (c >= 0) v = src_slot[(long long)c * nz + z];
  }
  src[i] = v;
}

// ---- cost-ordered unit lists for large scattered point sets (lh_point_kernel, fix_mode 3) -------
// The cost of a point is set by its radius: small rD puts quadrature nodes beyond the fast-path
// bound (literal path, ~15x per node).  Mixed at random, nearly every SM runs the large literal
// code next to the fast path all the time (instruction-cache misses were 30 % of the point
// kernel's stalls on C5b); ordered by radius -- smallest, i.e. most expensive, first -- the heavy
// points run together and first, the rest of the launch runs the compact fast path only, and
// the tail of the launch consists of cheap units.  Half-octave bins of rD; the order inside a
// bin is arbitrary (results are written by point index, so the output does not depend on it).
__device__ __forceinline__ int cost_bin(double rD) {
  const int hi = __double2hiint(rD);
  const int e = ((hi >> 20) & 0x7ff) - 1023;
  const int b = 2 * (e + 24) + ((hi >> 19) & 1);
  return min(max(b, 0), 63);
}

__global__ void cost_hist_kernel(const double *__restrict__ rD, long long n, unsigned int *__restrict__ bins) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) atomicAdd(&bins[cost_bin(rD[i])], 1u);
}

__global__ void cost_scan_kernel(unsigned int *bins) {   // bins[0..63] counts -> bins[64..127] cursors
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    unsigned int run = 0;
    for (int b = 0; b < 64; ++b) { bins[64 + b] = run; run += bins[b]; }
  }
}

__global__ void cost_scatter_kernel(const double *__restrict__ rD, long long n, unsigned int *__restrict__ bins,
                                    int *__restrict__ list) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) list[atomicAdd(&bins[64 + cost_bin(rD[i])], 1u)] = (int)i;
}

}  // namespace unc
