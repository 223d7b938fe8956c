"""Model identities on the GPU at sizes the oracle could not check point by point
(SURVEY.md 8(c) pin 4): no oracle in the loop, only closed forms and relations between the
reference's own solution types (laplace_hankel_solutions.f90:122-202), evaluated through the
C ABI on contour grids that run through the 128-z grid kernels."""
import numpy as np
import pytest
import scipy.special as sp

import unconfined_b200 as ub
from helpers import load_deck

pytestmark = pytest.mark.gpu


def contour(nr=256, nz=128, r_lo=0.5, r_hi=40.0):
    _, pd = load_deck("theis-contours-input.dat")
    tD = np.array([1.0, 30.0, 1e3, 1e5])
    sv = ub.split_index(tD, (1, 1))
    rD = np.geomspace(r_lo, r_hi, nr)
    zD = np.linspace(0.0, 1.0, nz)
    return pd, tD, sv, rD, zD


def test_theis_on_the_grid_kernels_against_the_exponential_integral():
    """s_D = E1(rD^2/4tD) (Theis 1935); the quadrature as coded has a ~1e-3 floor,
    the same for every z of a column."""
    pd, tD, sv, rD, zD = contour()
    lay = ub.zlay(zD, pd["lD"], pd["dD"])
    s, ds = ub.eval_grid(ub.Params(dict(pd, model=0)), tD, sv, rD, zD, lay)      # (nt, nr, nz)
    assert s.shape == (4, 256, 128)
    assert np.array_equal(s, np.repeat(s[:, :, :1], 128, axis=2))               # no z dependence, bitwise
    u = rD[None, :] ** 2 / (4 * tD[:, None])
    exact = sp.exp1(u)
    crit = np.abs(s[:, :, 0] - exact) / (3e-3 * exact + 1e-5)
    rel = (np.abs(s[:, :, 0] - exact) / exact)[exact > 1e-3]
    print("Theis vs E1: worst |err|/(3e-3 E1 + 1e-5)", crit.max(), "median rel", np.median(rel))
    assert crit.max() < 1.0 and np.median(rel) < 5e-4
    # d s/d ln t = exp(-u); the derivative inversion (p F(p)) is the less accurate of the two
    dex = np.exp(-u)
    errd = np.abs(ds[:, :, 0] - dex)
    print("Theis derivative vs exp(-u): max abs", errd.max(), "median rel", np.median((errd / dex)[dex > 1e-3]))
    assert errd.max() < 0.1 and np.median((errd / dex)[dex > 1e-3]) < 5e-3


def test_fully_penetrating_hantush_is_theis_on_the_grid_kernels():
    """lD = 1, dD = 0: the three-layer Hantush expressions collapse to Theis for every z
    (the idea of hantush-fullpen-test.in); away from the small radii where the layer
    expressions overflow (the reference's NaN flow, tested elsewhere) the two agree to rounding."""
    pd, tD, sv, rD, zD = contour()
    lay = ub.zlay(zD, 1.0, 0.0)
    th, dth = ub.eval_grid(ub.Params(dict(pd, model=0)), tD, sv, rD, zD, lay)
    ha, dha, fl = ub.eval_grid(ub.Params(dict(pd, model=1, lD=1.0, dD=0.0, bD=1.0)), tD, sv, rD, zD, lay,
                               want_flags=True)
    assert not fl.any()
    scale = np.abs(th).max(axis=(1, 2), keepdims=True)
    err = np.abs(ha - th) / scale
    errd = np.abs(dha - dth) / np.abs(dth).max(axis=(1, 2), keepdims=True)
    print("Hantush(full) - Theis: max |diff|/max|s| per time", err.max(axis=(1, 2)), "derivative", errd.max(axis=(1, 2)))
    assert err.max() < 1e-9 and errd.max() < 1e-9


def test_neuman_limits_on_the_grid_kernels():
    """Model 5 with beta = 0 is Neuman 1974; Moench (model 3) with one very large gamma (instantaneous
    drainage) tends to it; and both lie between the two Theis curves of elastic and
    specific-yield storage at late time (drawdown below the confined Hantush value)."""
    d, pd = load_deck("cape-cod-neuman74.in")
    tD = np.geomspace(1e1, 1e6, 6)
    sv = ub.split_index(tD, (2, 2))
    rD = np.geomspace(0.3, 3.0, 64)
    zD = np.linspace(0.0, 1.0, 128)
    lay = ub.zlay(zD, pd["lD"], pd["dD"])
    neu, _ = ub.eval_grid(ub.Params(dict(pd, model=5, beta=0.0)), tD, sv, rD, zD, lay)
    moe, _ = ub.eval_grid(ub.Params(dict(pd, model=3, moench_gamma=np.array([1e9]))), tD, sv, rD, zD, lay)
    han, _ = ub.eval_grid(ub.Params(dict(pd, model=1)), tD, sv, rD, zD, lay)
    assert np.isfinite(neu).all() and np.isfinite(moe).all()
    scale = np.abs(neu).max()
    print("Moench(gamma=1e9) - Neuman: max/scale", np.abs(moe - neu).max() / scale)
    assert np.abs(moe - neu).max() < 1e-5 * scale
    # the water table can only add water: drawdown never exceeds the confined (no-flow top) value
    assert (neu <= han + 1e-6 * scale).all()
