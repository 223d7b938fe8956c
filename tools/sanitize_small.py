"""Small workloads through every kernel and the carry post-pass (a quick smoke; compute-sanitizer is
closed on this pool, so bad accesses are hunted with small cases and the oracle comparison of tests/):
  python tools/sanitize_small.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import unconfined_b200 as ub  # noqa: E402

d, t, r, z = bench.c5a_grid(0, nr=3, nz=128, nt=1)
p, tD, sv, rD, zD, lay = bench.derive(d, t, r, z, ub)
prm = ub.Params(p)
for force in (None, "grid2", "point"):
    ub.force_kernel(force)
    zz, ll = (zD, lay) if force != "point" else (zD[:5], lay[:5])
    s, ds, fl = ub.eval_grid(prm, tD, sv, rD, zz, ll, want_flags=True)
    print(force or "grid8", s.shape, float(np.nanmax(np.abs(s))), int(fl.sum()))
ub.force_kernel(None)
zz = np.linspace(0, 1, 150)
s, ds = ub.eval_grid(prm, tD, sv, rD[:2], zz, ub.zlay(zz, p["lD"], p["dD"]))
print("two z-blocks", s.shape)
# reference-compatible mode with stale points (carry post-pass: list, scan, link, two point-kernel passes)
rr = np.array([rD[2], 1e-3, rD[1], 2e-3])
for nz in (5, 70, 128):
    zz = np.linspace(0, 1, nz)
    s, ds, fl = ub.eval_grid(prm, np.array([tD[0], 3 * tD[0]]), np.array([sv[0], sv[0]], np.int32), rr, zz,
                             ub.zlay(zz, p["lD"], p["dD"]), ts_scale=float(p["j0z"][sv[0] - 1] / rr[0]), want_flags=True)
    print("carry nz", nz, s.shape, int(fl.sum()), int(np.isnan(s).sum()))
print("ok")
