"""GPU-vs-oracle parity report over the shipped decks and a scattered sample.

Run on a GPU box:  python tools/parity_report.py [--out gpurun_out/parity.txt]
(The oracle is used here as the checker only.)
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import oracle, deck  # noqa: E402
import unconfined_b200 as ub  # noqa: E402

DECKS = ["theis-input.dat", "hantush-input.dat", "cape-cod-neuman74.in", "cape-cod-moench.in",
         "malama-partpen-input.dat", "malama-fullpen-input.dat", "hantush-storage-input.dat",
         "hantush-fullpen-test.in", "theis-contours-input.dat", "hantush-contours-input.dat",
         "mishra-neuman-malama.in"]


def relerr(a, b):
    a = np.asarray(a); b = np.asarray(b)
    both_nan = np.isnan(a) & np.isnan(b)
    same = (a == b) | both_nan
    den = np.maximum(np.abs(b), 1e-300)
    e = np.abs(a - b) / den
    e[same] = 0.0
    e[np.isnan(e)] = np.inf
    return e


def report(name, sg, dg, so, do, fl, out):
    es, ed = relerr(sg, so), relerr(dg, do)
    msg = (f"{name:34s} n={sg.size:6d} max_rel s={es.max():.3e} ds={ed.max():.3e} "
           f"n(s>1e-9)={int((es > 1e-9).sum())} n(ds>1e-9)={int((ed > 1e-9).sum())} "
           f"stale_flags={int((fl != 0).sum())}")
    print(msg)
    out.write(msg + "\n")
    bad = np.argwhere(es > 1e-9)
    for idx in bad[:5]:
        i = tuple(idx)
        m2 = f"    worst-ish at {i}: gpu={sg[i]!r} oracle={so[i]!r}"
        print(m2); out.write(m2 + "\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "parity.txt"))
    ap.add_argument("--scatter", type=int, default=256)
    ap.add_argument("--c5a", type=int, default=1024)
    args = ap.parse_args()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    out = open(args.out, "w")
    for name in DECKS:
        d = deck.read_deck(os.path.join(ROOT, "configs", name))
        pd = deck.params_dict(d)
        po, pg = oracle.Params(pd), ub.Params(pd)
        stale = d["j0z"][d["sv"][0] - 1] / d["rD"][0]   # reference-compatible abscissa scale
        for mode, sc in (("ref-stale", stale), ("fresh", None)):
            if mode == "fresh" and d["timeseries"] and d["j0s"][0] == d["j0s"][1]:
                continue  # identical to stale
            t0 = time.time()
            so, do, fo = oracle.eval_grid(po, d["tD"], d["sv"], d["rD"], d["zD"], d["zLay"],
                                          ts_scale=sc, carry=False)
            t1 = time.time()
            sg, dg, fg = ub.eval_grid(pg, d["tD"], d["sv"], d["rD"], d["zD"], d["zLay"],
                                      ts_scale=sc, want_flags=True)
            t2 = time.time()
            report(f"{name}[{mode}]", sg, dg, so, do, fg, out)
            if not np.array_equal(fo, fg):
                m = f"    FLAG MISMATCH oracle {int((fo != 0).sum())} gpu {int((fg != 0).sum())}"
                print(m); out.write(m + "\n")
            print(f"    oracle {t1 - t0:.2f}s gpu {t2 - t1:.3f}s")
    # scattered sample, C5b style
    d = deck.read_deck(os.path.join(ROOT, "configs", "malama-partpen-input.dat"))
    rng = np.random.default_rng(20261018)
    n = args.scatter
    rD = 10 ** rng.uniform(-2, 1, n); zD = rng.uniform(0, 1, n); tD = 10 ** rng.uniform(-1, 7, n)
    d["j0s"] = (2, 2)
    sv = oracle.split_index(tD, d["j0s"])
    lay = oracle.zlay(zD, d["lD"], d["dD"])
    pd = deck.params_dict(d)
    po, pg = oracle.Params(pd), ub.Params(pd)
    t0 = time.time()
    so, do, fo = oracle.eval_points(po, tD, sv, rD, zD, lay)
    t1 = time.time()
    sg, dg, fg = ub.eval_points(pg, tD, sv, rD, zD, lay, want_flags=True)
    t2 = time.time()
    report("scatter C5b", sg, dg, so, do, fg, out)
    print(f"    oracle {t1 - t0:.2f}s gpu {t2 - t1:.3f}s; flag mismatch {int((fo != fg).sum())}")
    # random sample of the benchmark grid itself (C5a, lh_grid8_kernel) against the oracle
    import bench
    dd, t, r, z = bench.c5a_grid(0)
    p, tDg, svg, rDg, zDg, layg = bench.derive(dd, t, r, z, ub)
    sg, dg, fg = ub.eval_grid(ub.Params(p), tDg, svg, rDg, zDg, layg, want_flags=True)
    rng = np.random.default_rng(42)
    n = args.c5a
    it, ir, iz = rng.integers(0, len(tDg), n), rng.integers(0, len(rDg), n), rng.integers(0, len(zDg), n)
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
    from helpers import NOISE_K, RTOL, oracle_with_noise
    po = oracle.Params(p)
    pts = (tDg[it], svg[it], rDg[ir], zDg[iz], layg[iz])
    so, do, sps, spd = oracle_with_noise(po, pts, points=True, nsamples=3)
    fo = oracle.eval_points(po, *pts)[2]
    keep = (fo == 0) & np.isfinite(so)
    gs, gd = sg[it, ir, iz][keep], dg[it, ir, iz][keep]
    es = relerr(gs, so[keep]); ed = relerr(gd, do[keep])
    ns = sps[keep] / np.maximum(np.abs(so[keep]), 1e-300); nd = spd[keep] / np.maximum(np.abs(do[keep]), 1e-300)
    us = np.abs(gs - so[keep]) / (RTOL * np.abs(so[keep]) + NOISE_K * sps[keep] + 1e-300)
    ud = np.abs(gd - do[keep]) / (RTOL * np.abs(do[keep]) + NOISE_K * spd[keep] + 1e-300)
    well = NOISE_K * sps[keep] <= RTOL * np.abs(so[keep])
    q = lambda e: " ".join(f"{np.quantile(e, x):.2e}" for x in (0.5, 0.9, 0.99, 1.0))  # noqa: E731
    m = (f"C5a grid sample (lh_grid8_kernel) n={int(keep.sum())} of {n}, flag mismatches {int((fo != fg[it, ir, iz]).sum())}\n"
         f"    quantiles 50/90/99/100%\n"
         f"    |gpu-oracle|/|oracle|                          s [{q(es)}]  ds [{q(ed)}]\n"
         f"    oracle's own noise (libm jitter, x87)/|oracle| s [{q(ns)}]  ds [{q(nd)}]\n"
         f"    |gpu-oracle| / (1e-9|oracle| + {NOISE_K:g} noise)       s [{q(us)}]  ds [{q(ud)}]   (<= 1 is the tests' bar)\n"
         f"    well-conditioned points ({NOISE_K:g} noise <= 1e-9|s|): {int(well.sum())}; on these |gpu-oracle|/|oracle| s [{q(es[well])}]"
         f"  within 1e-9 outright: {float((es[well] <= 1e-9).mean()):.3f}")
    print(m); out.write(m + "\n")
    try:
        pk = ub.measure_fp64_peak()
        n_, nominal = ub.device_info()
        m = f"fp64 DFMA peak measured {pk / 1e12:.2f} TFLOP/s (nominal {nominal / 1e12:.2f})"
        print(m); out.write(m + "\n")
    except Exception as e:  # noqa: BLE001
        print("peak measurement failed:", e)
    out.close()


if __name__ == "__main__":
    main()
