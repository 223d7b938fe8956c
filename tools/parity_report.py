"""GPU-vs-oracle parity report: every shipped deck (reference-compatible mode: stale tanh-sinh
abscissae AND stale-infint carry on both sides; contour decks also with fresh abscissae), a
scattered C5b sample and a random sample of the C5a benchmark grid.  Nothing is masked: flagged
(stale-infint) points are compared like all others.  For s and for ds the report gives the
outright max relative difference, the number of points beyond 1e-9, and the same restricted to
"quiet" points, where the oracle's own rounding-noise spread (libm jitter <= 2 ulp, x87 long
double) is below 1e-10 relative -- points at which 1e-9 is a meaningful bar.

Run on a GPU box:  python tools/parity_report.py [--out gpurun_out/parity.txt]
(The oracle is used here as the checker only.)
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import oracle, deck  # noqa: E402
import unconfined_b200 as ub  # noqa: E402
from helpers import NOISE_K, RTOL, oracle_with_noise  # noqa: E402

DECKS = sorted(f for f in os.listdir(os.path.join(ROOT, "configs")) if f.endswith("-input.dat") or f.endswith(".in"))
BASELINE = ("theis-input.dat", "hantush-input.dat", "cape-cod-neuman74.in", "cape-cod-moench.in")


def relerr(a, b):
    a = np.asarray(a, float); b = np.asarray(b, float)
    same = (a == b) | (np.isnan(a) & np.isnan(b))
    with np.errstate(all="ignore"):
        e = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
    e[same] = 0.0
    e[np.isnan(e)] = np.inf
    return e


def stats(g, r, sp):
    e = relerr(g, r)
    with np.errstate(all="ignore"):
        quiet = np.isfinite(r) & (sp < 1e-10 * np.abs(r))
        bar = np.abs(np.asarray(g) - r) / (RTOL * np.abs(r) + NOISE_K * sp + 1e-300)
    bar[(np.asarray(g) == r) | (np.isnan(g) & np.isnan(r))] = 0.0
    return (f"max {e.max():.2e} >1e-9: {int((e > 1e-9).sum()):4d}/{e.size:<5d} quiet {int(quiet.sum()):5d} "
            f"max@quiet {(e[quiet].max() if quiet.any() else 0.0):.2e}  max/bar {np.nanmax(bar):.2f}")


def line(out, msg):
    print(msg, flush=True)
    out.write(msg + "\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "parity.txt"))
    ap.add_argument("--scatter", type=int, default=256)
    ap.add_argument("--c5a", type=int, default=1024)
    args = ap.parse_args()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    out = open(args.out, "w")
    line(out, "GPU (C ABI, host arrays) against the CPU oracle; columns per quantity: outright max relative difference, "
              "points beyond 1e-9, quiet points (oracle noise < 1e-10), max at quiet points, max in units of the tests' bar")
    ub.force_kernel(None)
    for name in DECKS:
        d = deck.read_deck(os.path.join(ROOT, "configs", name))
        pd = deck.params_dict(d)
        if pd["model"] == 6 and pd.get("mn_type", 1) != 1:
            continue
        po, pg = oracle.Params(pd), ub.Params(pd)
        args_ = (d["tD"], d["sv"], d["rD"], d["zD"], d["zLay"])
        stale = d["j0z"][d["sv"][0] - 1] / d["rD"][0]   # reference-compatible abscissa scale
        for mode, sc in (("reference-compatible", stale), ("fresh abscissae", None)):
            if sc is None and d["timeseries"] and d["j0s"][0] == d["j0s"][1]:
                continue  # identical to the reference-compatible run
            carry = sc is not None
            so, do, sps, spd = oracle_with_noise(po, args_, ts_scale=sc, carry=carry, nsamples=4)
            fo = oracle.eval_grid(po, *args_, ts_scale=sc, carry=carry)[2]
            sg, dg, fg = ub.eval_grid(pg, *args_, ts_scale=sc, want_flags=True)
            tag = "*" if name in BASELINE else " "
            line(out, f"{tag}{name:30s} [{mode:20s}] model {pd['model']} n={sg.size:4d} stale-infint points {int((fg != 0).sum()):4d}"
                      f"{'' if np.array_equal(fo, fg) else '  FLAG MISMATCH'}")
            line(out, f"      s : {stats(sg, so, sps)}")
            line(out, f"      ds: {stats(dg, do, spd)}")
    # scattered sample, C5b style
    d = deck.read_deck(os.path.join(ROOT, "configs", "malama-partpen-input.dat"))
    rng = np.random.default_rng(20261018)
    n = args.scatter
    rD = 10 ** rng.uniform(-2, 1, n); zD = rng.uniform(0, 1, n); tD = 10 ** rng.uniform(-1, 7, n)
    d["j0s"] = (2, 2)
    sv = oracle.split_index(tD, d["j0s"])
    lay = oracle.zlay(zD, d["lD"], d["dD"])
    pd = dict(deck.params_dict(d), j0z=oracle.j0_zeros(2 + d["gl_nacc"] + 1))
    po, pg = oracle.Params(pd), ub.Params(pd)
    pts = (tD, sv, rD, zD, lay)
    so, do, sps, spd = oracle_with_noise(po, pts, points=True, nsamples=3)
    fo = oracle.eval_points(po, *pts)[2]
    sg, dg, fg = ub.eval_points(pg, *pts, want_flags=True)
    line(out, f" scattered C5b sample (point kernel) n={n} stale-infint points {int((fg != 0).sum())}, flag mismatches {int((fo != fg).sum())}")
    line(out, f"      s : {stats(sg, so, sps)}")
    line(out, f"      ds: {stats(dg, do, spd)}")
    # random sample of the benchmark grid itself (C5a, lh_grid8_kernel) against the oracle
    import bench
    dd, t, r, z = bench.c5a_grid(0)
    p, tDg, svg, rDg, zDg, layg = bench.derive(dd, t, r, z, ub)
    sg, dg, fg = ub.eval_grid(ub.Params(p), tDg, svg, rDg, zDg, layg, want_flags=True)
    rng = np.random.default_rng(42)
    n = args.c5a
    it, ir, iz = rng.integers(0, len(tDg), n), rng.integers(0, len(rDg), n), rng.integers(0, len(zDg), n)
    po = oracle.Params(p)
    pts = (tDg[it], svg[it], rDg[ir], zDg[iz], layg[iz])
    so, do, sps, spd = oracle_with_noise(po, pts, points=True, nsamples=3)
    fo = oracle.eval_points(po, *pts)[2]
    gs, gd = sg[it, ir, iz], dg[it, ir, iz]
    line(out, f" C5a benchmark grid, random sample (lh_grid8_kernel) n={n}, flag mismatches {int((fo != fg[it, ir, iz]).sum())}")
    line(out, f"      s : {stats(gs, so, sps)}")
    line(out, f"      ds: {stats(gd, do, spd)}")
    for nm, g_, r_, sp_ in (("s", gs, so, sps), ("ds", gd, do, spd)):
        with np.errstate(all="ignore"):
            well = np.isfinite(r_) & (NOISE_K * sp_ <= RTOL * np.abs(r_))
        e = relerr(g_, r_)
        q = " ".join(f"{np.quantile(e[well], x):.2e}" for x in (0.5, 0.9, 0.99, 1.0)) if well.any() else "-"
        line(out, f"      {nm}: well-conditioned points ({NOISE_K:g} x noise <= 1e-9|{nm}|): {int(well.sum())}; |gpu-oracle|/|oracle| "
                  f"quantiles 50/90/99/100% [{q}]; within 1e-9 outright: {float((e[well] <= 1e-9).mean()) if well.any() else 0:.3f}")
    out.close()


if __name__ == "__main__":
    main()
