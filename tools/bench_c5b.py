"""Timing of the scattered-point path (SURVEY 8(d) "C5b": every point its own (r,z,t), point
kernel, lanes <-> abscissae).  Not the headline bench (bench.py times C5a); prints one JSON line.
  python tools/bench_c5b.py [log2_points] [steps]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import unconfined_b200 as ub  # noqa: E402

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 18
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n = 1 << lg
d = bench.c5_deck()
rng = np.random.default_rng(20261018)
Lc = d["b"]; Tc = Lc ** 2 / (d["Kr"] / d["Ss"])
rD = 10 ** rng.uniform(-2, 1, n); zD = rng.uniform(0, 1, n); tD = 10 ** rng.uniform(-1, 7, n)
p, _, _, _, _, _ = bench.derive(d, np.array([1.0]), np.array([1.0]), np.array([0.0]), ub)
sv = ub.split_index(tD, d["j0s"]); lay = ub.zlay(zD, p["lD"], p["dD"])
prm = ub.Params(p)
dev = torch.device("cuda", 0)
g = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dev, dtype=dt)  # noqa: E731
a = [g(tD, torch.float64), g(sv, torch.int32), g(rD, torch.float64), g(zD, torch.float64), g(lay, torch.int32)]
s = torch.empty(n, dtype=torch.float64, device=dev); ds = torch.empty_like(s)
for _ in range(2):
    ub.eval_points_device(prm, *a, s, ds)
torch.cuda.synchronize()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
for e0, e1 in ev:
    e0.record(); ub.eval_points_device(prm, *a, s, ds); e1.record()
torch.cuda.synchronize()
ms = float(np.mean([e0.elapsed_time(e1) for e0, e1 in ev]))
F = bench.flops_per_point(p, lay, 1)
peak = ub.measure_fp64_peak()
print(json.dumps({"workload": f"C5b 2^{lg} scattered points, Malama partial penetration", "ms_per_step": ms,
                  "points_per_s": n / (ms * 1e-3), "flops_per_point": F,
                  "roofline_frac": F * n / (ms * 1e-3) / peak, "finite": float(torch.isfinite(s).float().mean())}))
