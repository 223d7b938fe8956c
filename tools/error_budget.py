#!/usr/bin/env python
"""Error budget of the GPU path against the CPU oracle on the four BASELINE decks.

  python tools/error_budget.py build      # (CPU box) one library per approximation switch
  python tools/error_budget.py run OUT    # (GPU box) table of max |gpu-oracle|/|oracle| per variant

Each variant replaces ONE approximation of the product build by the reference's own operation
(csrc/fast.cuh "Error-budget switches"); `product` is the shipped build.  The oracle side is
computed once (with its noise envelope) and shared.  Test/diagnostic tooling only.
"""
import json
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
BDIR = os.path.join(ROOT, "tools", "budget")
VARIANTS = {"product": [], "ieee_div": ["-DUNC_BUDGET_IEEE_DIV"], "libm": ["-DUNC_BUDGET_LIBM"],
            "literal": ["-DUNC_BUDGET_LITERAL"], "seqsum": ["-DUNC_BUDGET_SEQSUM"],
            "neville": ["-DUNC_BUDGET_NEVILLE"], "no_fma_contraction": ["-fmad=false"],
            "all_reference_like": ["-DUNC_BUDGET_IEEE_DIV", "-DUNC_BUDGET_LIBM", "-DUNC_BUDGET_LITERAL",
                                   "-DUNC_BUDGET_NEVILLE", "-fmad=false"]}
DECKS = ["theis-input.dat", "hantush-input.dat", "cape-cod-neuman74.in", "cape-cod-moench.in"]


def build():
    from unconfined_b200.build import NVCC_FLAGS, CSRC, _nvcc
    os.makedirs(BDIR, exist_ok=True)

    def one(item):
        name, defs = item
        so = os.path.join(BDIR, f"lib_{name}.so")
        r = subprocess.run([_nvcc()] + NVCC_FLAGS + defs + ["-o", so, os.path.join(CSRC, "capi.cu")],
                           capture_output=True, text=True)
        return name, r.returncode, r.stderr[-2000:]
    with ThreadPoolExecutor(max_workers=4) as ex:
        for name, rc, err in ex.map(one, VARIANTS.items()):
            print(name, "ok" if rc == 0 else "FAILED\n" + err, flush=True)
            if rc:
                raise SystemExit(1)


def child(name, npz):
    """One variant, own process (the library path is fixed at first use)."""
    import numpy as np
    import unconfined_b200.api as api
    api._SO = os.path.join(BDIR, f"lib_{name}.so")
    import unconfined_b200 as ub
    from helpers import load_deck, stale_scale
    ref = np.load(npz)
    out = {}
    ub.force_kernel("point")     # the decks' kernel; pinned so that every variant runs the same one
    for dk in DECKS:
        d, pd = load_deck(dk)
        sg, dg = ub.eval_grid(ub.Params(pd), d["tD"], d["sv"], d["rD"], d["zD"], d["zLay"], ts_scale=stale_scale(d))
        row = {}
        for q, g in (("s", sg), ("ds", dg)):
            r, sp = ref[f"{dk}:{q}"], ref[f"{dk}:sp_{q}"]
            rel = np.abs(g - r) / np.abs(r)
            quiet = sp < 1e-10 * np.abs(r)
            row[q] = {"max_rel": float(np.nanmax(rel)), "n_over_1e-9": int((rel > 1e-9).sum()),
                      "n": int(rel.size), "quiet_frac": float(quiet.mean()),
                      "max_rel_quiet": float(np.nanmax(rel[quiet])) if quiet.any() else None,
                      "median_rel": float(np.nanmedian(rel))}
        out[dk] = row
    print("RESULT " + json.dumps(out))


def run(out_path):
    import numpy as np
    from oracle import oracle
    from helpers import load_deck, stale_scale, oracle_with_noise
    npz = os.path.join(BDIR, "oracle_ref.npz")
    ref = {}
    for dk in DECKS:
        d, pd = load_deck(dk)
        args = (d["tD"], d["sv"], d["rD"], d["zD"], d["zLay"])
        so, do, sps, spd = oracle_with_noise(oracle.Params(pd), args, ts_scale=stale_scale(d), carry=True)
        ref[f"{dk}:s"], ref[f"{dk}:ds"], ref[f"{dk}:sp_s"], ref[f"{dk}:sp_ds"] = so, do, sps, spd
    np.savez(npz, **ref)
    res = {}
    for name in VARIANTS:
        if not os.path.exists(os.path.join(BDIR, f"lib_{name}.so")):
            continue
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "child", name, npz],
                           capture_output=True, text=True)
        line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
        if not line:
            res[name] = {"error": (r.stdout + r.stderr)[-600:]}
            continue
        res[name] = json.loads(line[0][7:])
    lines = ["error budget: max |gpu-oracle|/|oracle| on the BASELINE decks (point kernel), per build variant",
             "variant              deck                      s: max_rel  >1e-9   ds: max_rel  >1e-9  quiet_ds  max_rel_quiet_ds"]
    for name, r in res.items():
        if "error" in r:
            lines.append(f"{name:20s} ERROR {r['error']}")
            continue
        for dk, row in r.items():
            s, ds = row["s"], row["ds"]
            mq = ds["max_rel_quiet"]
            lines.append(f"{name:20s} {dk:24s} {s['max_rel']:10.2e} {s['n_over_1e-9']:4d}/{s['n']:<4d} "
                         f"{ds['max_rel']:10.2e} {ds['n_over_1e-9']:4d}/{ds['n']:<4d} {ds['quiet_frac']:6.2f}  "
                         f"{mq if mq is None else format(mq, '.2e')}")
    txt = "\n".join(lines)
    print(txt)
    with open(out_path, "w") as f:
        f.write(txt + "\n")
    with open(os.path.splitext(out_path)[0] + ".json", "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build()
    elif sys.argv[1] == "child":
        child(sys.argv[2], sys.argv[3])
    else:
        run(sys.argv[2])
