"""Shared helpers for the parity tests (test infrastructure; the oracle is the checker)."""
import os

import numpy as np

from oracle import oracle, deck

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# Parity bar (BASELINE.json north_star): relative 1e-9 on dimensionless drawdown and on its
# log-time derivative.  The reference's own double-precision result carries rounding
# noise that exceeds 1e-9 at ill-conditioned points (de Hoog amplifies
# ulp-level differences; e.g. malama-partpen early times differ by >100% between the
# double and long-double builds of the same algorithm).  A point therefore passes if
#   |gpu - oracle| <= RTOL*|oracle| + NOISE_K * spread
# where spread = how far the ORACLE's own result moves when every libm result is
# perturbed by <= 2 ulp (oracle.noise_envelope).  Well-conditioned points (spread
# below RTOL*|oracle|/NOISE_K) are thus held to 1e-9; the report also counts them.
RTOL = 1e-9
NOISE_K = 12.0


def load_deck(name):
    d = deck.read_deck(os.path.join(ROOT, "configs", name))
    return d, deck.params_dict(d)


def stale_scale(d):
    """Reference-compatible tanh-sinh abscissa scale: arg of the first (t,r) (driver.f90:121-126)."""
    return d["j0z"][d["sv"][0] - 1] / d["rD"][0]


def check_parity(got_s, got_ds, ref_s, ref_ds, sp_s, sp_ds, what=""):
    got_s, got_ds = np.asarray(got_s), np.asarray(got_ds)
    for name, g, r, sp in (("s", got_s, ref_s, sp_s), ("ds", got_ds, ref_ds, sp_ds)):
        nan_same = np.isnan(g) == np.isnan(r)
        assert nan_same.all(), f"{what} {name}: NaN pattern differs at {np.argwhere(~nan_same)[:5]}"
        ok = np.isnan(r) | (np.abs(g - r) <= RTOL * np.abs(r) + NOISE_K * sp) | (g == r)
        if not ok.all():
            i = tuple(np.argwhere(~ok)[0])
            raise AssertionError(f"{what} {name}: {int((~ok).sum())} of {ok.size} points fail; first {i}: "
                                 f"gpu={g[i]!r} oracle={r[i]!r} spread={sp[i]:.3e}")
    well = (NOISE_K * sp_s <= RTOL * np.abs(ref_s))
    return float(well.mean())


def oracle_with_noise(po, fn_args, points=False, nsamples=5, carry=False, **kw):
    """Oracle result plus its own rounding-noise envelope per point: the larger of
    (a) the spread under <=2-ulp libm jitter (nsamples draws) and (b) the distance to the
    same algorithm run in x87 long double.  The noise is heavy-tailed (Wynn's 1/denom, the
    q-d divisions), hence both estimates and the factor NOISE_K.  carry=True: the reference's
    stale-infint semantics (driver.f90:205-214), columns in (t outer, r inner) order."""
    if points:
        f = lambda **k2: oracle.eval_points(po, *fn_args, **kw, **k2)  # noqa: E731
    else:
        f = lambda **k2: oracle.eval_grid(po, *fn_args, carry=carry, **kw, **k2)  # noqa: E731
    s0, d0, sp_s, sp_d = oracle.noise_envelope(f, nsamples=nsamples)
    sl, dl = f(long_double=True)[:2]
    with np.errstate(invalid="ignore"):
        sp_s = np.fmax(sp_s, np.abs(s0 - sl))
        sp_d = np.fmax(sp_d, np.abs(d0 - dl))
    # a handful of draws under-samples a heavy-tailed noise: the LEVEL of the noise varies
    # smoothly with (t,r,z), so take the running maximum over +-2 neighbours on every axis
    if points:          # scattered points have no neighbours
        return s0, d0, sp_s, sp_d
    return s0, d0, _running_max(sp_s), _running_max(sp_d)


def _running_max(a, half=2):
    a = np.nan_to_num(np.asarray(a, float), nan=0.0, posinf=0.0)
    out = a.copy()
    for ax in range(a.ndim):
        n = a.shape[ax]
        if n == 1:
            continue
        acc = out.copy()
        for sh in range(1, half + 1):
            for sgn in (1, -1):
                rolled = np.roll(out, sgn * sh, axis=ax)
                idx = [slice(None)] * a.ndim
                idx[ax] = slice(0, sh) if sgn == 1 else slice(n - sh, n)
                rolled[tuple(idx)] = 0.0
                acc = np.maximum(acc, rolled)
        out = acc
    return out
