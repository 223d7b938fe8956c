"""Independent known-answer values for the oracle: the Laplace-Hankel solutions evaluated with
mpmath by DIFFERENT numerical methods than the reference's (Hankel integral by mpmath.quadosc
on the J0 oscillation, Laplace inversion by the fixed-Talbot method, 30 digits), from the
formulas of laplace_hankel_solutions.f90:64-93,133-202 and time.f90:49.  Writes
tests/golden/independent_mpmath.json.  Takes a few minutes.

These pin the MATHEMATICS (kernel formulas, Hankel inversion, Laplace inversion) of the oracle,
not the reference's rounding: with the decks' own quadrature orders the reference algorithm
itself is only ~1e-3 accurate (tanh-sinh k=7 on the first J0 interval), so the test refines the
orders and asserts convergence towards these values."""
import json
import os
import sys

import mpmath as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import load_deck  # noqa: E402

mp.mp.dps = 30


def kernel(pd, zD, L):
    kappa, lD, dD, bD = [mp.mpf(pd[k]) for k in ("kappa", "lD", "dD", "bD")]
    model = pd["model"]

    def udp(a, p, eta, z, lay):
        ff1, ff2, sh = mp.sinh(eta * dD), mp.sinh(eta * (1 - lD)), mp.sinh(eta)
        g2 = (ff1 * mp.cosh(eta * z) + ff2 * mp.cosh(eta * (1 - z))) / sh
        if lay == 1:
            u = (mp.exp(-eta * (1 - lD)) - (ff1 + mp.exp(-eta) * ff2) / sh) * mp.cosh(eta * z)
        elif lay == 2:
            u = 1 - g2
        else:
            u = mp.cosh(eta * (1 - dD - z)) - g2
        return u * 2 / (p + a * a) / bD

    def f(a, p):
        eta = mp.sqrt((p + a * a) / kappa)
        if model == 1:
            return udp(a, p, eta, zD, L)
        xi = eta * mp.mpf(pd["alphaD"]) / p                  # model 5 (Neuman 1974 when beta = 0)
        top = udp(a, p, eta, mp.mpf(1), 3)
        den = (1 + mp.mpf(pd["beta"]) * eta * xi) * mp.cosh(eta) + xi * mp.sinh(eta)
        return udp(a, p, eta, zD, L) - top * mp.cosh(eta * zD) / den
    return f


def drawdown(pd, rD, zD, L, tD):
    f = kernel(pd, mp.mpf(zD), L)
    rD = mp.mpf(rD)
    F = lambda p: mp.quadosc(lambda a: a * mp.besselj(0, a * rD) * f(a, p), [0, mp.inf], omega=rD) / p  # noqa: E731
    return mp.invertlaplace(F, mp.mpf(tD), method="talbot", degree=20)


out = []
for name, its in (("hantush-input.dat", (30, 60)), ("cape-cod-neuman74.in", (45,))):
    d, pd = load_deck(name)
    for it in its:
        v = drawdown(pd, d["rD"][0], d["zD"][0], int(d["zLay"][0]), d["tD"][it])
        out.append({"deck": name, "time_index": it, "tD": float(d["tD"][it]), "rD": float(d["rD"][0]),
                    "zD": float(d["zD"][0]), "s_D": float(v)})
        print(out[-1], flush=True)
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "independent_mpmath.json"), "w"), indent=1)
