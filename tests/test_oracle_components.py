"""CPU tests of the oracle's components against known answers (SURVEY.md 8c pins 3-4).

The reference has no test-suite; the only numeric fixture it ships for this path is
besJ0zeros.dat.  Everything else is pinned by identities at the tolerance the
reference's own quadrature allows."""
import math
import os

import numpy as np
import pytest
import scipy.special as sp

from oracle import oracle, deck
from helpers import load_deck, ROOT


def base_params(**kw):
    d = dict(model=0, M=10, alpha=1e-8, tol=1e-9, ts_k=7, ts_R=5, gl_nacc=10, gl_ord=50,
             j0z=oracle.j0_zeros(12), kappa=0.1, alphaD=0.1, beta=0.0, lD=0.75, dD=0.25, bD=0.5,
             rDw=1e-3, l=7.5, d=2.5, Ss=1e-6, rDwobs=1e-3, sF=20.0, time_type=1, time_par=[0.0, 1.0])
    d.update(kw)
    return oracle.Params(d)


def test_j0_zeros_match_reference_fixture():
    ref = np.loadtxt(os.path.join(ROOT, "tests", "golden", "besJ0zeros_first101.dat"))
    assert ref[0] == 0.0
    z = oracle.j0_zeros(100)
    rel = np.abs(z - ref[1:]) / ref[1:]
    assert rel.max() < 5e-16          # file has 16 significant digits (2.4e-16)


def test_j0_zero_is_a_zero():
    z = oracle.j0_zeros(20)
    assert np.all(np.abs(sp.j0(z)) < 1e-15)


def test_cbesk_series_and_miller_vs_scipy_amos():
    # scipy.special.kv wraps the same Amos algorithm (zbesk)
    rng = np.random.default_rng(1)
    for mag in (1e-6, 1e-3, 0.04, 0.5, 1.9, 2.1, 5.0, 17.0, 40.0, 300.0):
        for ang in (0.0, 0.3, 0.78, 1.2, 1.5):
            z = mag * complex(math.cos(ang), math.sin(ang))
            k0, k1, ierr, nz = oracle.cbesk01(z)
            assert ierr == 0 and nz == 0
            assert abs(k0 - sp.kv(0, z)) <= 4e-15 * abs(sp.kv(0, z))
            assert abs(k1 - sp.kv(1, z)) <= 4e-15 * abs(sp.kv(1, z))


def test_cbesk_error_codes():
    assert oracle.cbesk01(0j)[2] == 1                      # z = 0 (cbessel.f90:1027)
    assert oracle.cbesk01(complex(1e300, 0))[2] == 4       # |z| too large (:1062-1075)


def test_gauss_lobatto_nodes_weights():
    for ord_ in (5, 10, 50):
        x, w = oracle.gauss_lobatto(ord_)
        n = ord_ - 1
        wend = 2.0 / (n * (n + 1))
        assert len(x) == ord_ - 2 and np.all(np.diff(x) < 0)          # descending, endpoints dropped
        assert abs(w.sum() + 2 * wend - 2.0) < 1e-14
        # exact for polynomials of degree <= 2*ord-3 (including the endpoint contributions)
        for deg in (2, 4, 2 * ord_ - 4):
            if deg % 2:
                continue
            quad = (w * x ** deg).sum() + 2 * wend
            assert abs(quad - 2.0 / (deg + 1)) < 2e-13


def test_tanh_sinh_weights_and_error_floor():
    for k in (5, 6, 7, 8):
        w, a = oracle.tanh_sinh(k, 2.0)
        assert len(w) == 2 ** k - 1 and abs(w.sum() - 2.0) < 1e-14
        assert np.all(np.diff(a) > 0) and a[0] > 0 and a[-1] < 2.0
    # the rule as coded truncates at |t|<2 and renormalises: k-independent floor
    w, a = oracle.tanh_sinh(7, 2.0)
    est = (2.0 / 2.0) * (w * a ** 2).sum()
    assert 1e-5 < abs(est - 8.0 / 3.0) / (8.0 / 3.0) < 1e-4


def test_extraptozero_exact_on_polynomials():
    x = np.array([0.5, 0.25, 0.125, 0.0625, 0.03125])
    for coef in ([3.0], [1.0, 2.0], [0.5, -1.0, 4.0], [2.0, 1.0, -3.0, 0.25, 1.5]):
        y = sum(c * x ** i for i, c in enumerate(coef)) * (1 + 0.5j)
        got = oracle.extrap(x, y)
        assert abs(got - coef[0] * (1 + 0.5j)) < 1e-12


def test_wynn_alternating_series_and_edge_semantics():
    n = 10
    terms = np.array([(-1.0) ** k / (k + 1) for k in range(n)], dtype=complex)
    acc, info = oracle.wynn(terms)
    assert info == 0 and abs(acc - math.log(2)) < 1e-7          # partial sum alone: 5e-2
    # truncation at the first non-finite term (integration.f90:140-160)
    t2 = terms.copy(); t2[6] = complex(np.nan, 0)
    acc2, info2 = oracle.wynn(t2)
    assert info2 == 1 and abs(acc2 - math.log(2)) < 1e-3
    acc6, _ = oracle.wynn(terms[:6])
    assert acc2 == acc6
    # fewer than 4 good terms -> sentinel, a default-real literal (integration.f90:147)
    t3 = terms.copy(); t3[3] = complex(np.inf, 0)
    acc3, info3 = oracle.wynn(t3)
    assert info3 == 2 and acc3 == complex(np.float32(-999999.9), 0)
    # |denom| <= epsilon(1.0) exits early with eps(m+1,j) (integration.f90:169-177)
    t4 = np.array([1.0, 0.5, 1e-17, 1e-18, 1e-19, 1e-20], dtype=complex)
    acc4, info4 = oracle.wynn(t4)
    assert info4 == 3 and abs(acc4 - 1.5) < 1e-15


def test_dehoog_known_transforms():
    prm = base_params(M=10)
    for t in (0.1, 1.0, 30.0, 1e4):
        tee = 2.0 * t
        p = oracle.pvalues(prm, tee)
        assert p.shape == (21,) and p[0].imag == 0 and abs(p[1].imag - math.pi / tee) < 1e-15
        assert abs(oracle.dehoog(prm, t, tee, 1 / p) - 1.0) < 5e-8
        if t <= 30:
            assert abs(oracle.dehoog(prm, t, tee, 1 / (p + 1)) - math.exp(-t)) < 5e-8
        r = 0.5     # Theis in Laplace space: 2 K0(r sqrt p)/p  ->  E1(r^2/4t)   
        f = 2 * sp.kv(0, r * np.sqrt(p)) / p
        exact = sp.exp1(r * r / (4 * t))
        assert abs(oracle.dehoog(prm, t, tee, f) - exact) < 2e-8 * max(exact, 1.0)


def test_dehoog_nan_and_zero_semantics():
    prm = base_params(M=10)
    p = oracle.pvalues(prm, 2.0)
    assert oracle.dehoog(prm, 1.0, 2.0, np.zeros(21, complex)) == 0.0      # invlap.f90:69,139
    assert oracle.dehoog(prm, 1.0, 2.0, np.full(21, complex(np.nan, np.nan))) == 0.0
    f = 1 / p
    g = f.copy(); g[5] = complex(np.nan, 1.0)                              # NaN entries zeroed (:71-74)
    h = f.copy(); h[5] = 0
    a, b = oracle.dehoog(prm, 1.0, 2.0, g), oracle.dehoog(prm, 1.0, 2.0, h)
    assert a == b or (math.isnan(a) and math.isnan(b))   # a zeroed entry poisons the q-d table: NaN


@pytest.mark.parametrize("tt,par", [(1, [0.5, 1.0]), (2, [0.5, 2.0]), (3, [0.3, 1.0]), (4, [1.0, 3.0]),
                                    (5, [1.0, 0.2]), (6, [2.0, 0.1]), (7, [1.0, 0.1]), (8, [1.0, 0.1]),
                                    (-2, [0.0, 1.0, 2.0, 1.0, 0.5]), (-102, [0.0, 1.0, 2.0, 1.0, 3.0])])
def test_lap_time_behaviours(tt, par):
    prm = base_params(time_type=tt, time_par=par)
    p = oracle.pvalues(prm, 4.0)
    got = oracle.lap_time(prm, p)
    e = np.exp
    if tt == 1: want = e(-par[0] * p) / p
    elif tt == 2: want = e(-par[0] * p) / p - e(-par[1] * p) / p
    elif tt == 3: want = e(-par[0] * p)
    elif tt == 4: want = 1 / (p - p * e(-par[0] * p)) * (1 - e(-par[1] * p)) / p
    elif tt == 5: want = e(-par[1] * p) / (p + p * e(-par[0] * p))
    elif tt == 6: want = e(-par[1] * p) * p / (p ** 2 + par[0] ** 2)
    elif tt == 7: want = np.zeros_like(p)        # identically zero as written (time.f90:72-74)
    elif tt == 8: want = e(-par[1] * p) * (1 - e(-par[0] * p / 2)) / ((1 + e(-par[0] * p / 2)) * p)
    elif tt == -2:
        ti, tf, Q = par[0:2], par[2], [0.0] + par[3:5]
        want = (sum((Q[i + 1] - Q[i]) * e(-ti[i] * p) for i in range(2)) - (Q[2] - Q[0]) * e(-tf * p)) / p
    else:
        ti, tf, y = par[0:2], par[2], par[3:5] + [0.0]
        W = [0.0] + [(y[i + 1] - y[i]) / (([ti[1], tf][i]) - ti[i]) for i in range(2)]
        want = (sum((W[i + 1] - W[i]) * e(-ti[i] * p) for i in range(2)) - (W[2] - W[0]) * e(-tf * p)) / p ** 2
    assert np.allclose(got, want, rtol=1e-13, atol=1e-300)


def test_soln_theis_and_hantush_fullpen_identity():
    # Hantush with lD=1, dD=0 is Theis in Laplace-Hankel space (idea of hantush-fullpen-test.in)
    th = base_params(model=0)
    ha = base_params(model=1, lD=1.0, dD=0.0, bD=1.0)
    z = np.array([0.0, 0.3, 0.9]); lay = oracle.zlay(z, 1.0, 0.0)
    for a in (0.01, 1.0, 30.0):
        f0 = oracle.soln(th, a, 0.5, 1.0, z, lay)
        f1 = oracle.soln(ha, a, 0.5, 1.0, z, lay)
        assert np.allclose(f0, f1, rtol=1e-11)
        p = oracle.pvalues(th, 2.0)
        want = a * sp.j0(a * 0.5) * (2 / (p + a * a)) / p
        assert np.allclose(f0[:, 0], want, rtol=1e-13)


def test_theis_deck_vs_exponential_integral():
    # loose: the quadrature as coded has a ~3e-5 floor
    d, pd = load_deck("hantush-input.dat")
    pd = dict(pd, model=0)
    s, ds, fl = oracle.eval_grid(oracle.Params(pd), d["tD"][20:60:10], d["sv"][20:60:10], d["rD"],
                                 d["zD"], d["zLay"], carry=False)
    exact = sp.exp1(d["rD"][0] ** 2 / (4 * d["tD"][20:60:10]))
    assert np.all(np.abs(s[:, 0, 0] - exact) / exact < 2e-3)   # up to 1.7e-3 at tD=1e4


def test_zlay_and_split_index_edges():
    lay = oracle.zlay([0.0, 0.2, 0.25, 0.5, 0.75, 0.9, 1.0], lD=0.75, dD=0.25)
    # zD<=0 or zD<1-lD -> 1 ; zD>=1 or zD<1-dD -> 2 ; else 3  (driver_io.f90:575-586)
    assert lay.tolist() == [1, 1, 2, 2, 3, 3, 2]
    sv = oracle.split_index(np.array([0.1, 10.0, 1e8]), (1, 1))
    assert sv.tolist() == [1, 1, 1]
    sv = oracle.split_index(np.array([1e-1, 1e2, 1e8]), (2, 6))
    assert sv[0] >= sv[1] >= sv[2] >= 2 and sv[0] <= 6


def test_oracle_reproduces_golden_fixtures():
    gdir = os.path.join(ROOT, "tests", "golden")
    for name in ("hantush-input.dat", "cape-cod-moench.in", "theis-input.dat"):
        g = np.load(os.path.join(gdir, "oracle_" + name.replace(".", "_") + ".npz"))
        d, pd = load_deck(name)
        assert np.array_equal(g["tD"], d["tD"]) and np.array_equal(g["sv"], d["sv"])
        s, ds, fl = oracle.eval_grid(oracle.Params(pd), d["tD"], d["sv"], d["rD"], d["zD"], d["zLay"],
                                     ts_scale=float(g["ts_scale"]), carry=False)
        assert np.array_equal(s, g["s"], equal_nan=True) and np.array_equal(ds, g["ds"], equal_nan=True)


def test_survey_smoke_values():
    # BASELINE.md section 4 (numpy probe of the survey, 6 digits)
    d, pd = load_deck("hantush-input.dat")
    s, ds, _ = oracle.eval_grid(oracle.Params(pd), d["tD"], d["sv"], d["rD"], d["zD"], d["zLay"], carry=False)
    for i, want in ((0, 1.23490e-10), (22, 1.93747), (44, 6.53545), (77, 13.4414)):
        assert abs(s[i, 0, 0] - want) / want < 2e-5


def test_stale_infint_carry_matches_reference_order():
    # driver.f90:209-211: when every GL area is 0/NaN, infint keeps the previous (t,r) value.
    d, pd = load_deck("hantush-contours-input.dat")
    po = oracle.Params(pd)
    rD = np.array([d["rD"][5], 1e-4, d["rD"][6]])      # middle radius: everything overflows
    sc = d["j0z"][d["sv"][0] - 1] / rD[0]
    s1, _, f1 = oracle.eval_grid(po, d["tD"], d["sv"], rD, d["zD"][:3], d["zLay"][:3], ts_scale=sc, carry=True)
    s0, _, f0 = oracle.eval_grid(po, d["tD"], d["sv"], rD, d["zD"][:3], d["zLay"][:3], ts_scale=sc, carry=False)
    assert np.array_equal(f0, f1)
    assert np.array_equal(s0[0, 0], s1[0, 0]) and np.array_equal(s0[0, 2], s1[0, 2])


def test_oracle_converges_to_independent_mpmath_values():
    """Known answers from a DIFFERENT numerical route (tools/make_independent_truth.py: mpmath
    quadosc for the Hankel integral, fixed-Talbot Laplace inversion, 30 digits) for Hantush
    (model 1) and Neuman 1974 (model 5, beta = 0).  With the decks' own orders the reference
    algorithm is ~1e-3..1e-4 accurate (tanh-sinh k=7 / k=6 on the first J0 interval);
    refining the tanh-sinh rule must bring the oracle to the algorithm's 3e-5 floor of them."""
    import json
    truth = json.load(open(os.path.join(ROOT, "tests", "golden", "independent_mpmath.json")))
    assert len(truth) >= 3
    for t in truth:
        d, pd = load_deck(t["deck"])
        it = t["time_index"]
        assert abs(d["tD"][it] - t["tD"]) < 1e-12 * t["tD"]
        err = {}
        for tag, kw in (("deck", {}), ("fine", dict(ts_k=9, ts_R=7))):
            q = dict(pd, **kw)
            s, _, _ = oracle.eval_grid(oracle.Params(q), d["tD"][it:it + 1], d["sv"][it:it + 1], d["rD"],
                                       d["zD"][:1], d["zLay"][:1], carry=False)
            err[tag] = abs(s.ravel()[0] - t["s_D"]) / t["s_D"]
        assert err["deck"] < 2e-3, (t["deck"], err)
        assert err["fine"] < 5e-5 and err["fine"] < err["deck"], (t["deck"], err)


def test_cbknu_underflow_branch_equals_scipy_amos():
    """Re z > alim = 664.87: cbknu keeps the values scaled by exp(z) and ckscl/cuchk decide which
    members underflow (cbessel.f90:5215,5458-5476,5499-5611,5895-5927).  scipy.special.kv wraps the
    same Amos routines in double precision: the restatement must agree with it to rounding and
    underflow to exact zeros where it does."""
    from scipy.special import kv
    rng = np.random.default_rng(5)
    zs = np.concatenate([rng.uniform(665, 697.5, 200) + 1j * rng.uniform(-400, 400, 200),
                         rng.uniform(698.5, 2000, 50) + 1j * rng.uniform(-400, 400, 50)])
    nzero = 0
    for z in zs:
        k0, k1, ierr, nz = oracle.cbesk01(complex(z))
        assert ierr in (0, 3)
        for got, want in ((k0, kv(0, z)), (k1, kv(1, z))):
            if want == 0:
                assert got == 0
                nzero += 1
            else:
                assert abs(got - want) <= 1e-12 * abs(want), (z, got, want)   # (bitwise on most hosts)
    assert nzero >= 100
