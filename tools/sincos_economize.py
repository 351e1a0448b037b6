"""Chebyshev economisation (exact rational arithmetic) of the sin(pi r) / cos(pi r) Taylor series on |r| <= 1/4:
the coefficients of tbk_math.cuh sincospi_lean (6 for the sine, 7 for the cosine) come from this script."""
from fractions import Fraction as F
from math import factorial
import numpy as np
PI = F("3.14159265358979323846264338327950288419716939937510582097494459230781640628620899")
H = F(1, 16)  # u = r^2 in [0, 1/16]

def cheb_T(n):  # coefficients of T_n(x) in monomials, ascending
    a, b = [F(1)], [F(0), F(1)]
    if n == 0: return a
    for _ in range(n - 1):
        c = [F(0)] + [2 * t for t in b]
        for i, t in enumerate(a): c[i] -= t
        a, b = b, c
    return b

def poly_mul(p, q):
    r = [F(0)] * (len(p) + len(q) - 1)
    for i, a in enumerate(p):
        for j, b in enumerate(q): r[i + j] += a * b
    return r

def to_x(pu):  # p(u), u = H (x + 1) / 2  -> polynomial in x
    res = [F(0)]
    base = [F(1)]
    lin = [H / 2, H / 2]
    for c in pu:
        res = [ (res[i] if i < len(res) else 0) + c * (base[i] if i < len(base) else 0) for i in range(max(len(res), len(base)))]
        base = poly_mul(base, lin)
    return res

def mono_to_cheb(px):  # exact: peel off highest degree
    px = list(px); n = len(px) - 1
    c = [F(0)] * (n + 1)
    for k in range(n, -1, -1):
        t = cheb_T(k)
        ck = px[k] / t[k]
        c[k] = ck
        for i in range(k + 1): px[i] -= ck * t[i]
    return c

def cheb_to_mono_u(c):  # sum c_k T_k(x), x = 2u/H - 1 -> polynomial in u
    px = [F(0)] * len(c)
    for k, ck in enumerate(c):
        for i, t in enumerate(cheb_T(k)): px[i] += ck * t
    # x = 2u/H - 1
    res = [F(0)]; base = [F(1)]; lin = [F(-1), 2 / H]
    for cx in px:
        res = [ (res[i] if i < len(res) else 0) + cx * (base[i] if i < len(base) else 0) for i in range(max(len(res), len(base)))]
        base = poly_mul(base, lin)
    return res

def economize(pu, deg):
    c = mono_to_cheb(to_x(pu))
    dropped = sum(abs(x) for x in c[deg + 1:])
    return cheb_to_mono_u(c[:deg + 1]), float(dropped)

K = 14
sin_c = [F((-1) ** k) * PI ** (2 * k + 1) / factorial(2 * k + 1) for k in range(K)]   # sin(pi r)/r = sum sin_c[k] u^k
cos_c = [F((-1) ** k) * PI ** (2 * k) / factorial(2 * k) for k in range(K)]           # cos(pi r)   = sum cos_c[k] u^k
ps_full = sin_c[1:]   # (sin(pi r)/r - pi)/u
pc_full = cos_c[1:]   # (cos(pi r) - 1)/u
for name, full, degs in (("sin", ps_full, (7, 6, 5)), ("cos", pc_full, (8, 7, 6, 5))):
    for d in degs:
        pe, dropped = economize(full, d)
        # error bound on the function: sin: r*u*dropped <= (1/4)(1/16) dropped ; cos: u*dropped <= dropped/16
        bound = dropped * (float(H) / 4 if name == "sin" else float(H))
        print(name, "coeffs", d + 1, "abs error bound %.2e" % bound)
        print("   ", ", ".join(repr(float(x)) for x in reversed(pe)))
