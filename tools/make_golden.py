"""Write tests/golden/oracle_<deck>.npz: oracle outputs on the patched decks.

Self-generated regression fixtures (the reference ships no golden outputs and cannot be
compiled here).  Each file holds the inputs actually fed to the evaluators (tD, sv, rD,
zD, zLay, ts_scale) and the oracle's s, ds, flags in reference-compatible mode
(tanh-sinh abscissae of the first (t,r), driver.f90:121-126; infint not carried).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle, deck  # noqa: E402

DECKS = ["theis-input.dat", "hantush-input.dat", "cape-cod-neuman74.in", "cape-cod-moench.in",
         "hantush-storage-input.dat", "hantush-contours-input.dat", "mishra-neuman-malama.in"]

for name in DECKS:
    d = deck.read_deck(os.path.join(ROOT, "configs", name))
    prm = oracle.Params(deck.params_dict(d))
    stale = d["j0z"][d["sv"][0] - 1] / d["rD"][0]
    s, ds, fl = oracle.eval_grid(prm, d["tD"], d["sv"], d["rD"], d["zD"], d["zLay"], ts_scale=stale,
                                 carry=False)
    out = os.path.join(ROOT, "tests", "golden", "oracle_" + name.replace(".", "_") + ".npz")
    np.savez_compressed(out, tD=d["tD"], sv=d["sv"], rD=d["rD"], zD=d["zD"], zLay=d["zLay"],
                        ts_scale=stale, s=s, ds=ds, flags=fl)
    print(out, s.shape)
