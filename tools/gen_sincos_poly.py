"""Generate the sin/cos kernel polynomials used by sincos_q() in csrc/fast.cuh.

On |r| <= pi/4 (after Cody-Waite reduction by multiples of pi/2):
    sin r = r + r^3 * g(r^2),      g(t) = (sin(sqrt t)/sqrt t - 1)/t
    cos r = 1 - r^2/2 + r^4 * h(r^2),  h(t) = (cos(sqrt t) - 1 + t/2)/t^2
g and h are interpolated at Chebyshev nodes of t in [0,(pi/4 + margin)^2] in 50-digit
arithmetic (near-minimax), degree 5, converted to monomials in t.
Prints the coefficient lines to paste into fast.cuh and the measured max error.
"""
import mpmath as mp
import numpy as np

mp.mp.dps = 50
DEG = 5
TMAX = (mp.pi / 4 * mp.mpf("1.002")) ** 2


def g(t):
    if t == 0:
        return -mp.mpf(1) / 6
    r = mp.sqrt(t)
    return (mp.sin(r) / r - 1) / t


def h(t):
    if t == 0:
        return mp.mpf(1) / 24
    r = mp.sqrt(t)
    return (mp.cos(r) - 1 + t / 2) / (t * t)


def fit(f, n):
    # interpolate at Chebyshev nodes on [0, TMAX]; solve Vandermonde in mp
    xs = [TMAX / 2 * (1 + mp.cos(mp.pi * (2 * k + 1) / (2 * (n + 1)))) for k in range(n + 1)]
    A = mp.matrix(n + 1, n + 1)
    b = mp.matrix(n + 1, 1)
    for i, x in enumerate(xs):
        for j in range(n + 1):
            A[i, j] = x ** j
        b[i] = f(x)
    c = mp.lu_solve(A, b)
    return [float(c[j]) for j in range(n + 1)]


gs, hs = fit(g, DEG), fit(h, DEG)
rs = np.linspace(-float(mp.pi / 4), float(mp.pi / 4), 20001)
t = rs * rs
gp = np.zeros_like(t); hp = np.zeros_like(t)
for c in gs[::-1]:
    gp = gp * t + c
for c in hs[::-1]:
    hp = hp * t + c
sn = rs + rs * t * gp
cs = (1 - 0.5 * t) + t * t * hp
es = max(abs(float(mp.sin(mp.mpf(float(r))) - mp.mpf(float(v)))) for r, v in zip(rs[::20], sn[::20]))
ec = max(abs(float(mp.cos(mp.mpf(float(r))) - mp.mpf(float(v)))) for r, v in zip(rs[::20], cs[::20]))
print("max abs err sin %.3e cos %.3e" % (es, ec))
print("SIN g:", ", ".join("%.20e" % c for c in gs))
print("COS h:", ", ".join("%.20e" % c for c in hs))
