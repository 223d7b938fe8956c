"""Small workloads through every kernel (a quick smoke; compute-sanitizer is closed on this pool):
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
for force in (None, "grid4", "grid2", "point"):
    if force:
        os.environ["UNC_FORCE_KERNEL"] = force
    else:
        os.environ.pop("UNC_FORCE_KERNEL", None)
    zz, ll = (zD, lay) if force != "point" else (zD[:5], lay[:5])
    s, ds, fl = ub.eval_grid(prm, tD, sv, rD, zz, ll, want_flags=True)
    print(force or "grid8", s.shape, float(np.nanmax(np.abs(s))), int(fl.sum()))
os.environ.pop("UNC_FORCE_KERNEL", None)
s, ds = ub.eval_grid(prm, tD, sv, rD[:2], np.linspace(0, 1, 150), ub.zlay(np.linspace(0, 1, 150), p["lD"], p["dD"]))
print("two z-blocks", s.shape)
print("ok")
