"""TEST INFRASTRUCTURE ONLY -- numpy emulation of the warp-level schedule of ``tbmodels_b200/csrc/eig_tridiag_reg.cu``
(register-resident Hermitian -> tridiagonal reduction, one warp per matrix, N <= 48).

Nothing like it exists in the reference (it calls LAPACK through scipy, src/tbmodels/_tb_model.py:1149).  This file
restates the device ALGORITHM lane by lane -- index reversal, full rows per lane, the "second slot" rows 32 .. N-1
that exist only as conj(XC) plus the corner block C, block-granular column loops over zero-padded v / w -- so that
its algebra is pinned against LAPACK on the CPU (tests/test_oracle_golden.py) before the kernel runs on a GPU.
Arrays indexed [lane] stand for per-lane registers; V, W, XC, C stand for the kernel's shared-memory arrays.
"""
from __future__ import annotations

import numpy as np

from .blocked_hetrd import reflector

LANES = 32


def _tri(i):
    return i * (i + 1) // 2


def _trs(i):
    return i * (i - 1) // 2


def pack_lower(A: np.ndarray) -> np.ndarray:
    """Packed Hermitian layout of tbk_math.cuh: real plane tri(i)+j (j <= i), imaginary plane ntri + trs(i)+j (j < i)."""
    n = A.shape[0]
    out = np.zeros(n * n)
    ntri = _tri(n)
    for i in range(n):
        for j in range(i + 1):
            out[_tri(i) + j] = A[i, j].real
            if j < i:
                out[ntri + _trs(i) + j] = A[i, j].imag
    return out


def reg_tridiagonalise(S: np.ndarray, N: int):
    """(d [N], e [N-1]) from the packed matrix ``S`` exactly as tridiag_reg_kernel<NREG, XMAX> schedules it."""
    nreg = min(32, -(-N // 4) * 4)
    X = max(N - 32, 0)
    xmax = 0 if X == 0 else (4 if X <= 4 else 8 if X <= 8 else 16)
    nv = max(nreg + xmax, 32)
    ntri = _tri(N)
    Sre, Sim = S, S[ntri:]
    lanes = np.arange(LANES)

    # ---- register fill: row `lane` of B[c][b] = A[N-1-c][N-1-b], columns b < nreg ----
    a = np.zeros((LANES, nreg), dtype=complex)
    for c in range(LANES):
        if c >= N:
            continue
        I = N - 1 - c
        for b in range(min(nreg, N)):
            J = N - 1 - b
            if b >= c:
                a[c, b] = Sre[_tri(I) + J] + (1j * Sim[_trs(I) + J] if b > c else 0.0)
            else:
                a[c, b] = Sre[_tri(J) + I] - 1j * Sim[_trs(J) + I]
    XC = np.zeros((max(xmax, 1), LANES), dtype=complex)
    C = np.zeros((max(xmax, 1), max(xmax, 1)), dtype=complex)
    for s in range(X):
        J = X - 1 - s
        for c in range(LANES):
            I = N - 1 - c
            XC[s, c] = Sre[_tri(I) + J] + 1j * Sim[_trs(I) + J]
    for r in range(X):
        for s in range(X):
            Ir, Js = X - 1 - r, X - 1 - s
            if Js <= Ir:
                C[r, s] = Sre[_tri(Ir) + Js] + (1j * Sim[_trs(Ir) + Js] if Js < Ir else 0.0)
            else:
                C[r, s] = Sre[_tri(Js) + Ir] - 1j * Sim[_trs(Js) + Ir]

    V = np.zeros(nv, dtype=complex)
    W = np.zeros(nv, dtype=complex)
    d1 = np.zeros(LANES)
    e1 = np.zeros(LANES)
    d2 = np.zeros(LANES)
    e2 = np.zeros(LANES)

    def blocks(m):  # columns the block-granular loops touch
        return min(nreg, -(-m // 4) * 4)

    p = N - 1
    while p >= 32:  # ---- corner steps ----
        xc = p - 32
        x1 = XC[xc].copy()
        x2 = np.zeros(LANES, dtype=complex)
        x2[:xc] = C[:xc, xc]
        d2[xc] = C[xc, xc].real
        if xc >= 1:
            alpha = x2[xc - 1]
            xn = np.sum(np.abs(x1) ** 2) + np.sum(np.abs(x2[: xc - 1]) ** 2)
        else:
            alpha = x1[31]
            xn = np.sum(np.abs(x1[:31]) ** 2)
        beta, tau, scale = reflector(complex(alpha), float(xn))
        if xc >= 1:
            e2[xc - 1] = beta
        else:
            e1[31] = beta
        if tau == 0:
            p -= 1
            continue
        v1 = x1 * scale
        v2 = np.zeros(LANES, dtype=complex)
        if xc >= 1:
            v2[: xc - 1] = x2[: xc - 1] * scale
            v2[xc - 1] = 1.0
        else:
            v1[31] = 1.0
        V[:32] = v1
        V[32 : 32 + xmax] = v2[:xmax]
        q1 = a[:, : blocks(32)] @ V[: blocks(32)]
        for s in range(xc):
            q1 = q1 + XC[s] * V[32 + s]
        q2 = np.zeros(LANES, dtype=complex)
        for r in range(xc):
            q2[r] = np.sum(np.conj(XC[r]) * v1)
        for r in range(xc):
            q2[r] += np.sum(C[r, :xc] * V[32 : 32 + xc])
        p1, p2 = tau * q1, tau * q2
        dot = np.sum(np.conj(p1) * v1) + np.sum(np.conj(p2) * v2)
        coef = -0.5 * tau * dot
        w1, w2 = p1 + coef * v1, p2 + coef * v2
        W[:32] = w1
        W[32 : 32 + xmax] = w2[:xmax]
        nb = blocks(32)
        a[:, :nb] -= np.outer(v1, np.conj(W[:nb])) + np.outer(w1, np.conj(V[:nb]))
        for s in range(xc):
            XC[s] -= v1 * np.conj(W[32 + s]) + w1 * np.conj(V[32 + s])
        for r in range(xc):
            for s in range(xc):
                C[r, s] -= V[32 + r] * np.conj(W[32 + s]) + W[32 + r] * np.conj(V[32 + s])
        p -= 1

    while p >= 1:  # ---- lean steps ----
        x = a[:, p].copy()
        d1[p] = x[p].real
        alpha = x[p - 1]
        xn = np.sum(np.abs(x[: p - 1]) ** 2)
        beta, tau, scale = reflector(complex(alpha), float(xn))
        e1[p - 1] = beta
        if tau == 0:
            p -= 1
            continue
        v = np.zeros(LANES, dtype=complex)
        v[: p - 1] = x[: p - 1] * scale
        v[p - 1] = 1.0
        V[:32] = v
        nb = blocks(p)
        q = a[:, :nb] @ V[:nb]
        pv = tau * q
        dot = np.sum(np.conj(pv) * v)
        coef = -0.5 * tau * dot
        w = np.zeros(LANES, dtype=complex)
        w[:p] = pv[:p] + coef * v[:p]
        W[:32] = w
        a[:, :nb] -= np.outer(v, np.conj(W[:nb])) + np.outer(w, np.conj(V[:nb]))
        p -= 1
    d1[0] = a[0, 0].real

    # ---- store, undoing the reversal ----
    d = np.zeros(N)
    e = np.zeros(max(N - 1, 0))
    for lane in range(LANES):
        if lane < N:
            d[N - 1 - lane] = d1[lane]
        if lane <= N - 2:
            e[N - 2 - lane] = e1[lane]
        if 32 + lane < N:
            d[N - 1 - 32 - lane] = d2[lane]
        if 32 + lane <= N - 2:
            e[N - 2 - 32 - lane] = e2[lane]
    return d, e
