"""Which grid kernel for mid-size columns?  C5a-style grid with nz = 32..128 through the lanes<->z
kernel (lh_grid_kernel) and through the 128-z persistent kernel (padding slots idle); ms per call,
device-resident.  Measurement tooling."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, unconfined_b200 as ub
dev = torch.device("cuda", 0)
g = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dev, dtype=dt)
rows = []
for nz in (32, 40, 48, 64, 80, 96, 128):
    d, t, r, z = bench.c5a_grid(0, nr=1024, nz=nz, nt=2)
    p, tD, sv, rD, zD, lay = bench.derive(d, t, r, z, ub)
    prm = ub.Params(p)
    ins = (g(tD, torch.float64), g(sv, torch.int32), g(rD, torch.float64), g(zD, torch.float64), g(lay, torch.int32))
    n = len(tD) * len(rD) * nz
    res = {}
    for kern in ("grid2", "grid8"):
        ub.force_kernel(kern)
        s = torch.empty(n, dtype=torch.float64, device=dev); ds = torch.empty_like(s)
        for _ in range(2): ub.eval_grid_device(prm, *ins, s, ds)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): ub.eval_grid_device(prm, *ins, s, ds)
        e1.record(); torch.cuda.synchronize()
        res[kern] = (e0.elapsed_time(e1) / 3, s.cpu().numpy())
    ok = np.isfinite(res["grid2"][1])
    rel = np.abs(res["grid8"][1][ok] - res["grid2"][1][ok]) / np.maximum(np.abs(res["grid2"][1][ok]), 1e-300)
    rows.append(f"nz={nz:4d}  lanes<->z {res['grid2'][0]:8.2f} ms   128-z kernel {res['grid8'][0]:8.2f} ms   median rel diff {np.median(rel):.1e}")
    print(rows[-1], flush=True)
ub.force_kernel(None)
open(sys.argv[1], "w").write("\n".join(rows) + "\n")
