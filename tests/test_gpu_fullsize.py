"""Size-independent properties at BASELINE.json's full size: the 2^20-point C5a contour
grid (Malama partial penetration) that bench.py times.  The oracle needs ~50 ms of CPU
per point, so here it only checks a random sample; the rest are invariants."""
import os

import numpy as np
import pytest

import bench
import unconfined_b200 as ub
from oracle import oracle
from helpers import check_parity, oracle_with_noise

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c5a():
    os.environ.pop("UNC_FORCE_KERNEL", None)
    d, t, r, z = bench.c5a_grid(0)
    p, tD, sv, rD, zD, lay = bench.derive(d, t, r, z, ub)
    prm = ub.Params(p)
    s, ds, fl = ub.eval_grid(prm, tD, sv, rD, zD, lay, want_flags=True)
    return dict(p=p, prm=prm, tD=tD, sv=sv, rD=rD, zD=zD, lay=lay, s=s, ds=ds, fl=fl)


def test_full_grid_shape_determinism_and_finiteness(c5a):
    g = c5a
    assert g["s"].shape == (8, 1024, 128) and g["s"].size == 2 ** 20
    s2, ds2 = ub.eval_grid(g["prm"], g["tD"], g["sv"], g["rD"], g["zD"], g["lay"])
    assert np.array_equal(s2, g["s"], equal_nan=True) and np.array_equal(ds2, g["ds"], equal_nan=True)
    clean = g["fl"] == 0
    assert clean.mean() > 0.97
    # non-finite results are data, as in the reference (NaN totlap entries, invlap.f90:71-74)
    assert np.isfinite(g["s"][clean]).mean() > 0.97


def test_full_grid_sharding_invariance(c5a):
    g = c5a
    half = 512
    sa, da = ub.eval_grid(g["prm"], g["tD"], g["sv"], g["rD"][:half], g["zD"], g["lay"])
    sb, db = ub.eval_grid(g["prm"], g["tD"], g["sv"], g["rD"][half:], g["zD"], g["lay"])
    assert np.array_equal(np.concatenate([sa, sb], axis=1), g["s"], equal_nan=True)
    assert np.array_equal(np.concatenate([da, db], axis=1), g["ds"], equal_nan=True)


def test_full_grid_physics_monotone_in_time_and_radius(c5a):
    g = c5a
    # away from the overflow-affected smallest radii, from the water table (where the
    # reference's double-precision formulas are rounding noise at early time, see
    # malama-partpen in DESIGN.md) and from the earliest times
    s = g["s"][3:, 64:, :100]
    assert np.isfinite(s).all()
    scale = np.abs(s).max()
    assert (np.diff(s, axis=0) > -1e-6 * scale).all()      # drawdown grows with time (step pumping)
    late = g["s"][-1, 64:, 64]
    assert (np.diff(late) < 1e-8 * scale).all()             # and decays with distance


def test_full_grid_random_sample_against_oracle(c5a):
    g = c5a
    rng = np.random.default_rng(7)
    n = 256
    it, ir, iz = rng.integers(0, 8, n), rng.integers(0, 1024, n), rng.integers(0, 128, n)
    args = (g["tD"][it], g["sv"][it], g["rD"][ir], g["zD"][iz], g["lay"][iz])
    po = oracle.Params(g["p"])
    so, do, sps, spd = oracle_with_noise(po, args, points=True, nsamples=2)
    _, _, fo = oracle.eval_points(po, *args)
    assert np.array_equal(fo, g["fl"][it, ir, iz])
    keep = fo == 0
    well = check_parity(g["s"][it, ir, iz][keep], g["ds"][it, ir, iz][keep], so[keep], do[keep],
                        sps[keep], spd[keep], what="C5a sample")
    print('well-conditioned fraction of the sample:', well)
    assert well > 0.3     # a good share of the sample is well-conditioned and held to 1e-9 outright


def test_grid_kernel_equals_point_kernel_on_sample(c5a):
    g = c5a
    rng = np.random.default_rng(11)
    n = 4096
    it, ir, iz = rng.integers(0, 8, n), rng.integers(64, 1024, n), rng.integers(0, 128, n)
    sp_, dp_ = ub.eval_points(g["prm"], g["tD"][it], g["sv"][it], g["rD"][ir], g["zD"][iz], g["lay"][iz])
    ref = g["s"][it, ir, iz]
    rel = np.abs(sp_ - ref) / np.maximum(np.abs(ref), 1e-300)
    # two different summation orders of the same arithmetic: agree far below the parity bar
    # except where the result itself is rounding noise (tiny early-time drawdowns)
    assert np.median(rel) < 1e-12
    assert (rel < 1e-6).mean() > 0.98


def test_small_radius_columns_overflow_flow_matches_oracle(c5a):
    """The few smallest radii of the C5a grid are where the reference's cosh/sinh overflow:
    Wynn truncation, sentinel and stale flags of the 128-z persistent kernel (which stops
    evaluating a p once every z's series is settled) must match the oracle exactly, and the
    finite results must agree."""
    g = c5a
    ir = np.arange(0, 12)
    it = np.array([1, 6])
    po = oracle.Params(g["p"])
    so, do, fo = oracle.eval_grid(po, g["tD"][it], g["sv"][it], g["rD"][ir], g["zD"], g["lay"], carry=False)
    sg, dg, fg = g["s"][it][:, ir], g["ds"][it][:, ir], g["fl"][it][:, ir]
    assert np.array_equal(fo, fg)
    assert np.array_equal(np.isnan(so), np.isnan(sg))
    # the overflow regime is really exercised: at these radii a/sqrt(kappa) exceeds 710
    assert g["p"]["j0z"][-1] / g["rD"][0] / np.sqrt(g["p"]["kappa"]) > 1000.0
    args = (g["tD"][it], g["sv"][it], g["rD"][ir], g["zD"], g["lay"])
    so2, do2, sps, spd = oracle_with_noise(po, args, nsamples=3)
    keep = fo == 0
    mask = lambda a: np.where(keep, a, 0.0)   # noqa: E731  (flagged points: documented deviation)
    check_parity(mask(sg), mask(dg), mask(so2), mask(do2), sps, spd, what="C5a small radii")


def test_grid4_equals_grid2_kernel_bitwise_semantics(c5a):
    """The persistent 128-z kernel and the 64-z kernel implement the same arithmetic except
    for the z-recurrence of the exponentials: results agree far below the parity bar."""
    g = c5a
    os.environ["UNC_FORCE_KERNEL"] = "grid2"
    try:
        s2, d2, f2 = ub.eval_grid(g["prm"], g["tD"][2:4], g["sv"][2:4], g["rD"][::16], g["zD"], g["lay"], want_flags=True)
    finally:
        os.environ.pop("UNC_FORCE_KERNEL", None)
    s4, f4 = g["s"][2:4][:, ::16], g["fl"][2:4][:, ::16]
    assert np.array_equal(f2, f4)
    ok = np.isfinite(s2) & np.isfinite(s4)
    rel = np.abs(s4[ok] - s2[ok]) / np.maximum(np.abs(s2[ok]), 1e-300)
    assert np.median(rel) < 1e-12 and (rel < 1e-7).mean() > 0.98


def test_grid8_equals_grid4_on_odd_column_counts_and_ragged_z(c5a):
    """lh_grid8_kernel (8 z-slots per lane, two Laplace parameters per warp) against
    lh_grid4_kernel on shapes that exercise its edges: odd numbers of columns, z-blocks that
    are not full, a second partial z-block."""
    g = c5a
    for nr, nt, zsel in ((5, 1, slice(0, 128)), (3, 3, slice(0, 100)), (2, 1, slice(0, 128, 1))):
        tD, sv = g["tD"][2:2 + nt], g["sv"][2:2 + nt]
        rD = g["rD"][100:100 + 37 * nr:37]
        zD, lay = g["zD"][zsel], g["lay"][zsel]
        s8, d8, f8 = ub.eval_grid(g["prm"], tD, sv, rD, zD, lay, want_flags=True)
        os.environ["UNC_FORCE_KERNEL"] = "grid4"
        try:
            s4, d4, f4 = ub.eval_grid(g["prm"], tD, sv, rD, zD, lay, want_flags=True)
        finally:
            os.environ.pop("UNC_FORCE_KERNEL", None)
        assert np.array_equal(f8, f4)
        assert np.array_equal(np.isnan(s8), np.isnan(s4))
        ok = np.isfinite(s4)
        rel = np.abs(s8[ok] - s4[ok]) / np.maximum(np.abs(s4[ok]), 1e-300)
        assert np.median(rel) < 1e-12 and (rel < 1e-7).mean() > 0.98, (nr, nt, np.median(rel))
    # 200 z = one full and one partial z-block, equally spaced; against the point kernel
    z = np.linspace(0.0, 1.0, 200)
    lay = ub.zlay(z, g["p"]["lD"], g["p"]["dD"])
    tD, sv, rD = g["tD"][4:5], g["sv"][4:5], g["rD"][300:303]
    s8, d8 = ub.eval_grid(g["prm"], tD, sv, rD, z, lay)
    os.environ["UNC_FORCE_KERNEL"] = "point"
    try:
        sp_, dp_ = ub.eval_grid(g["prm"], tD, sv, rD, z, lay)
    finally:
        os.environ.pop("UNC_FORCE_KERNEL", None)
    ok = np.isfinite(sp_)
    rel = np.abs(s8[ok] - sp_[ok]) / np.maximum(np.abs(sp_[ok]), 1e-300)
    assert np.median(rel) < 1e-12 and (rel < 1e-6).mean() > 0.98
