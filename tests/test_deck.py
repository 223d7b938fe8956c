"""Deck handling: the patched decks parse under the reference's current format rules
(driver_io.f90:88-529) and the derived quantities follow driver_io.f90:531-664."""
import math
import os

import numpy as np

from oracle import deck
from helpers import ROOT

CFG = os.path.join(ROOT, "configs")


def test_all_shipped_decks_parse():
    n = 0
    for f in sorted(os.listdir(CFG)):
        if f.endswith("-input.dat") or f.endswith(".in"):
            d = deck.read_deck(os.path.join(CFG, f))
            assert 0 <= d["model"] <= 6 and len(d["tD"]) >= 1 and len(d["j0z"]) == max(d["j0s"]) + d["gl_nacc"] + 1
            n += 1
    assert n >= 10


def test_theis_deck_is_a_contour_run():
    d = deck.read_deck(os.path.join(CFG, "theis-input.dat"))
    assert not d["timeseries"] and (len(d["tD"]), len(d["rD"]), len(d["zD"])) == (1, 30, 20)
    assert abs(d["tD"][0] - 15.5) < 1e-13 and abs(d["rD"][0] - 0.075) < 1e-16


def test_hantush_deck_derived_quantities():
    d = deck.read_deck(os.path.join(CFG, "hantush-input.dat"))
    assert len(d["tD"]) == 100 and abs(d["tD"][0] - 0.1) < 1e-16 and abs(d["tD"][-1] - 1e8) < 1e-6
    assert abs(d["lD"] - 0.055) < 1e-17 and abs(d["dD"] - 0.045) < 1e-17 and abs(d["zD"][0] - 0.05) < 1e-17
    assert d["zLay"].tolist() == [1] and d["rD"][0] == 0.5 and set(d["sv"]) == {1}


def test_moench_deck():
    d = deck.read_deck(os.path.join(CFG, "cape-cod-moench.in"))
    assert d["model"] == 3 and len(d["moench_gamma"]) == 3 and len(d["zD"]) == 2 and not d["dimless"]
    g = 2.78e-4 * d["b"] * d["Sy"] / (d["kappa"] * d["Kr"])
    assert d["moench_gamma"][0] == g
    assert abs(d["Tc"] - 168.9 ** 2 * 1.305e-5 / 2.331e-1) < 1e-12
    assert abs(d["Hc"] - 42.8 / (4 * math.pi * 0.2331 * 168.9)) < 1e-15
