"""Wall time of the reference's own deck configurations (BASELINE.json configs[0..3] and the
other shipped decks) through the C ABI with host buffers (H2D + kernels + D2H), next to the CPU
oracle port on all host cores.  These are 100-600-point jobs: latency, not throughput."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import unconfined_b200 as ub  # noqa: E402
from oracle import oracle  # noqa: E402
from helpers import load_deck, stale_scale  # noqa: E402

DECKS = ["theis-input.dat", "hantush-input.dat", "cape-cod-neuman74.in", "cape-cod-moench.in",
         "malama-partpen-input.dat", "hantush-storage-input.dat", "hantush-contours-input.dat",
         "mishra-neuman-malama.in"]
cores = len(os.sched_getaffinity(0))
print(f"{'deck':32s} {'points':>7s} {'GPU ms':>9s} {'CPU ms':>10s} {'ratio':>8s}   (CPU: oracle port, {cores} threads)")
for name in DECKS:
    d, pd = load_deck(name)
    args = (d["tD"], d["sv"], d["rD"], d["zD"], d["zLay"])
    sc = stale_scale(d)
    prm = ub.Params(pd)
    ub.eval_grid(prm, *args, ts_scale=sc)
    ts = []
    for _ in range(7):
        t0 = time.perf_counter(); s, ds = ub.eval_grid(prm, *args, ts_scale=sc); ts.append(time.perf_counter() - t0)
    g = float(np.median(ts)) * 1e3
    po = oracle.Params(pd)
    t0 = time.perf_counter(); oracle.eval_grid(po, *args, ts_scale=sc, carry=False, nthreads=cores); c = (time.perf_counter() - t0) * 1e3
    print(f"{name:32s} {s.size:7d} {g:9.2f} {c:10.1f} {c / g:8.0f}x")
