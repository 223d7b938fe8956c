"""How much do layer boundaries inside the z-range cost the 128-z kernel?  The C5a grid (screen in the
top 2 % of the aquifer: one slot straddles a boundary) against the same grid with the screen in
the middle third (three layers, two boundary slots); ms per call, device-resident, nt = 2.
Measurement tooling.   usage: bench_layers.py OUT.txt"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, unconfined_b200 as ub
dev = torch.device("cuda", 0)
g = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dev, dtype=dt)
rows = []
for what, frac in (("C5a (screen at the top)", None), ("screen in the middle third", (1.0 / 3.0, 2.0 / 3.0))):
    d, t, r, z = bench.c5a_grid(0, nr=1024, nz=128, nt=2)
    if frac:
        d = dict(d, d=frac[0] * d["b"], l=frac[1] * d["b"])
    p, tD, sv, rD, zD, lay = bench.derive(d, t, r, z, ub)
    prm = ub.Params(p)
    ins = (g(tD, torch.float64), g(sv, torch.int32), g(rD, torch.float64), g(zD, torch.float64), g(lay, torch.int32))
    n = len(tD) * len(rD) * len(zD)
    s = torch.empty(n, dtype=torch.float64, device=dev); ds = torch.empty_like(s)
    for _ in range(2): ub.eval_grid_device(prm, *ins, s, ds)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): ub.eval_grid_device(prm, *ins, s, ds)
    e1.record(); torch.cuda.synchronize()
    rows.append(f"{what:32s} layers {sorted(set(lay.tolist()))}  {e0.elapsed_time(e1) / 3:8.2f} ms per {n} points   finite {float(torch.isfinite(s).double().mean()):.3f}")
    print(rows[-1], flush=True)
open(sys.argv[1], "w").write("\n".join(rows) + "\n")
