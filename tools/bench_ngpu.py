#!/usr/bin/env python
"""Strong scaling of the path the product ships to a single-process caller (the Fortran driver):
ONE unc_eval_grid_ex call with host arrays on ONE 2^20-point C5a grid, ngpu = 1, 2, 4, 8.
The library splits the (t,r) columns over the devices (one host thread + stream per device) and
copies the results into disjoint slices of the caller's arrays; no collective.

  python tools/bench_ngpu.py OUT.json [reps]

Wall-clock of the synchronous call (it covers H2D, kernels on every device and D2H), best and
median of `reps`; bitwise comparison of every N against N=1.  Diagnostic/measurement tooling.
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import unconfined_b200 as ub  # noqa: E402


def pin(a):
    """Page-lock a numpy array in place (the library's D2H copies are then asynchronous)."""
    try:
        rt = C.CDLL("libcudart.so.12")
    except OSError:
        try:
            rt = C.CDLL("/usr/local/cuda/lib64/libcudart.so")
        except OSError:
            return False
    return rt.cudaHostRegister(C.c_void_p(a.ctypes.data), C.c_size_t(a.nbytes), 1) == 0


def main():
    out_path = sys.argv[1]
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 7
    navail = ub.device_count()
    d, t, r, z = bench.c5a_grid(0)
    p, tD, sv, rD, zD, lay = bench.derive(d, t, r, z, ub)
    prm = ub.Params(p)
    nt, nr, nz = len(tD), len(rD), len(zD)
    npts = nt * nr * nz
    s = np.empty((nt, nr, nz)); ds = np.empty((nt, nr, nz))
    pinned = pin(s) and pin(ds)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))  # noqa: E731
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))  # noqa: E731
    lib = ub.lib()

    def call(n, ts=None):
        rc = lib.unc_eval_grid_ex(C.byref(prm.s), nt, dp(tD), ip(sv), nr, dp(rD), nz, dp(zD), ip(lay),
                                  None if ts is None else dp(ts), n, dp(s), dp(ds), None)
        if rc != 0:
            raise RuntimeError(ub.last_error())

    res = {"workload": f"C5a {nr}r x {nz}z x {nt}t = {npts} points, ONE grid, unc_eval_grid_ex(ngpu=N), host arrays "
                       f"({'page-locked' if pinned else 'pageable'})",
           "devices_visible": navail, "reps": reps, "runs": []}
    ref = None
    t1 = None
    for n in (1, 2, 3, 4, 8):
        if n > navail:
            continue
        for _ in range(2):
            call(n)
        times = []
        for _ in range(reps):
            t0 = time.perf_counter(); call(n); times.append(time.perf_counter() - t0)
        if ref is None:
            ref = (s.copy(), ds.copy())
        same = bool(np.array_equal(s, ref[0], equal_nan=True) and np.array_equal(ds, ref[1], equal_nan=True))
        best, med = min(times), float(np.median(times))
        if n == 1:
            t1 = best
        res["runs"].append({"ngpu": n, "best_ms": 1e3 * best, "median_ms": 1e3 * med,
                            "points_per_s": npts / best, "speedup_vs_1": t1 / best,
                            "efficiency": t1 / best / n, "bitwise_equal_to_ngpu1": same})
        print(res["runs"][-1], flush=True)
    # reference-compatible mode (stale abscissae + stale-infint carry) on all devices vs one
    ts = np.full((nt, nr), p["j0z"][sv[0] - 1] / rD[0])
    call(1, ts); a1 = (s.copy(), ds.copy())
    t0 = time.perf_counter(); call(1, ts); tc1 = time.perf_counter() - t0
    call(0, ts)
    t0 = time.perf_counter(); call(0, ts); tcn = time.perf_counter() - t0
    res["reference_compatible_mode"] = {
        "ngpu1_ms": 1e3 * tc1, "all_gpus_ms": 1e3 * tcn,
        "bitwise_equal": bool(np.array_equal(s, a1[0], equal_nan=True) and np.array_equal(ds, a1[1], equal_nan=True))}
    print(res["reference_compatible_mode"], flush=True)
    json.dump(res, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
