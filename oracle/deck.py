"""TEST INFRASTRUCTURE ONLY -- oracle-side restatement of the reference's deck reader.

Follows driver_io.f90:88-666 (read_input): the 18-line positional deck, the time
file (input-explanation.txt:223-250) and the space file (:255-288), then the
non-dimensionalisation (driver_io.f90:531-567), the z-layer classification
(:572-586, via oracle.zlay), the J0 zeros (:628-647, via oracle.j0_zeros) and the
split index (:658-664, via oracle.split_index).  Used by tests/ and bench.py to feed
the SAME inputs to the oracle and to the CUDA library, and to check the product's
own C++ deck reader.  PARITY UNPINNED (see oracle_math.hpp).
"""
import math
import os

import numpy as np

from . import oracle


def _num(tok):
    return float(tok.replace("D", "E").replace("d", "e"))


def _logical(tok):
    t = tok.strip(".").lower()
    if t.startswith("t"):
        return True
    if t.startswith("f"):
        return False
    raise ValueError(f"bad logical {tok!r}")


def _toks(line):
    return line.replace(",", " ").split()


def linspace(lo, hi, num):
    """utility.f90:34-50"""
    if num == 1:
        return np.array([(lo + hi) / 2.0])
    dx = (hi - lo) / (num - 1)
    return np.array([lo + i * dx for i in range(num)])


def logspace(lo, hi, num):
    """utility.f90:52-57   10.0_DP**linspace(real(lo),real(hi),num)  (glibc pow)"""
    return np.array([math.pow(10.0, x) for x in linspace(float(lo), float(hi), num)])


def read_deck(path):
    base = os.path.dirname(os.path.abspath(path))
    L = open(path).read().split("\n")
    d = {}
    t = _toks(L[0])
    d["quiet"], d["model"] = int(t[0]), int(t[1])
    d["dimless"], d["timeseries"], d["piezometer"] = map(_logical, t[2:5])
    d["Q"] = _num(_toks(L[1])[0])
    t = _toks(L[2]); d["l"], d["d"] = _num(t[0]), _num(t[1])
    t = _toks(L[3]); d["rw"], d["rc"] = _num(t[0]), _num(t[1])
    d["gammaSkin"] = _num(_toks(L[4])[0])
    t = _toks(L[5])
    d["time_type"] = int(t[0])
    if d["time_type"] > -1:
        npar = 2
    else:
        npar = -2 * int(math.fmod(d["time_type"], 100)) + 1   # driver_io.f90:124
    d["time_par"] = [_num(x) for x in t[1:1 + npar]]
    d["b"] = _num(_toks(L[6])[0])
    t = _toks(L[7]); d["Kr"], d["kappa"] = _num(t[0]), _num(t[1])
    t = _toks(L[8]); d["Ss"], d["Sy"] = _num(t[0]), _num(t[1])
    t = _toks(L[9])
    d["beta"], d["MoenchM"] = _num(t[0]), int(t[1])
    d["MoenchAlpha"] = [_num(x) for x in t[2:2 + d["MoenchM"]]]
    t = _toks(L[10])
    d["ac"], d["ak"], d["psia"], d["psik"], d["usL"] = [_num(x) for x in t[:5]]
    d["MNtype"], d["order"] = int(t[5]), int(t[6])
    t = _toks(L[11]); d["M"], d["alpha"], d["tol"] = int(t[0]), _num(t[1]), _num(t[2])
    if d["tol"] < np.finfo(float).eps:
        d["tol"] = float(np.finfo(float).eps)
    t = _toks(L[12]); d["ts_k"], d["ts_R"] = int(t[0]), int(t[1])
    t = _toks(L[13]); d["j0s"] = (int(t[0]), int(t[1])); d["gl_nacc"], d["gl_ord"] = int(t[2]), int(t[3])
    t = _toks(L[14]); timefile, tval = t[0], _num(t[1])
    t = _toks(L[15]); spacefile, rval = t[0], _num(t[1])
    t = _toks(L[16])
    d["zTop"], d["zBot"], d["zOrd"] = _num(t[0]), _num(t[1]), int(t[2])
    d["rwobs"], d["sF"] = _num(t[3]), _num(t[4])
    d["outfile"] = _toks(L[17])[0]

    if d["timeseries"]:
        r = np.array([rval])
        if d["piezometer"]:
            d["zOrd"] = 1
        z = linspace(d["zBot"], d["zTop"], d["zOrd"])
        T = open(os.path.join(base, timefile)).read().split("\n")
        t0 = _toks(T[0]); compute, numfile = _logical(t0[0]), int(t0[1])
        t1 = _toks(T[1]); minlog, maxlog, numcomp = int(t1[0]), int(t1[1]), int(t1[2])
        if compute:
            tt = logspace(minlog, maxlog, numcomp)
        else:
            tt = np.array([_num(_toks(T[2 + i])[0]) for i in range(numfile)])
    else:
        tt = np.array([tval])
        S = open(os.path.join(base, spacefile)).read().split("\n")
        t0 = _toks(S[0]); compute, nrf, nzf = _logical(t0[0]), int(t0[1]), int(t0[2])
        t1 = _toks(S[1]); minR, maxR, nrc = _num(t1[0]), _num(t1[1]), int(t1[2])
        t2 = _toks(S[2]); minZ, maxZ, nzc = _num(t2[0]), _num(t2[1]), int(t2[2])
        if compute:
            r = linspace(minR, maxR, nrc)
            z = linspace(minZ, maxZ, nzc)
        else:
            r = np.array([_num(x) for x in _toks(S[3])[:nrf]])
            z = np.array([_num(x) for x in _toks(S[4])[:nzf]])
    d["t"], d["r"], d["z"] = tt, r, z
    return derive(d)


def derive(d):
    """driver_io.f90:531-567, 572-586, 628-664."""
    PI = 4.0 * math.atan(1.0)
    Lc = d["b"]
    Tc = Lc ** 2 / (d["Kr"] / d["Ss"])
    d["Lc"], d["Tc"] = Lc, Tc
    d["Hc"] = d["Q"] / (4 * PI * d["Kr"] * d["b"])
    sigma = d["Sy"] / (d["Ss"] * d["b"])
    d["alphaD"] = d["kappa"] / sigma
    d["lD"] = d["l"] / Lc
    d["dD"] = d["d"] / Lc
    d["bD"] = d["lD"] - d["dD"]
    d["rDw"] = d["rw"] / Lc
    d["rDwobs"] = d["rwobs"] / Lc
    d["moench_gamma"] = [a * Lc * d["Sy"] / (d["kappa"] * d["Kr"]) for a in d["MoenchAlpha"]]
    d["zD"] = np.asarray(d["z"]) / Lc
    d["rD"] = np.asarray(d["r"]) / Lc
    d["tD"] = np.asarray(d["t"]) / Tc
    d["zLay"] = oracle.zlay(d["zD"], d["lD"], d["dD"])
    terms = max(d["j0s"]) + d["gl_nacc"] + 1
    d["j0z"] = oracle.j0_zeros(terms)
    d["sv"] = oracle.split_index(d["tD"], d["j0s"])
    return d


def params_dict(d):
    """The fields of unc_params / orc::params."""
    keys = ("model", "M", "alpha", "tol", "time_type", "time_par", "ts_k", "ts_R", "gl_nacc",
            "gl_ord", "j0z", "kappa", "alphaD", "beta", "moench_gamma", "lD", "dD", "bD", "rDw",
            "l", "d", "Ss", "rDwobs", "sF")
    out = {k: d[k] for k in keys}
    out["tee_mult"] = 2.0     # driver.f90:54
    # model 6 (laplace_hankel_solutions.f90:404-442 reads f%ak, f%b, f%psia, f%psik, f%Sy, f%Ss)
    out.update(mn_type=d["MNtype"], mn_ak=d["ak"], mn_psia=d["psia"], mn_psik=d["psik"],
               mn_b=d["b"], mn_Sy=d["Sy"])
    return out
