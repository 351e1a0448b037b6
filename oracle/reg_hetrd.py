"""TEST INFRASTRUCTURE ONLY -- numpy emulation of the warp-level schedule of ``tbmodels_b200/csrc/eig_tridiag_reg.cu``
(register-resident Hermitian -> tridiagonal reduction, one warp per matrix, N <= 48).

Nothing like it exists in the reference (it calls LAPACK through scipy, src/tbmodels/_tb_model.py:1149).  This file
restates the device ALGORITHM lane by lane -- index reversal, full rows per lane, the "second slot" rows 32 .. N-1
that exist only as conj(XC) plus the corner block C, block-granular column loops over zero-padded v / w, and the
single fused reduction per step (product with the UNSCALED column, v^H A v from three sums) -- so that
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

    def scalars(alpha, s1, s2, s3, adiag):
        """(beta, tau, scale, coef): reflector + the real coefficient of v in w (one fused reduction: s1, s2, s3)."""
        beta, tau, scale = reflector(complex(alpha), float(s1))
        g = abs(scale) ** 2 * s2 + 2.0 * (np.conj(scale) * s3).real + adiag
        return beta, tau, scale, -0.5 * abs(tau) ** 2 * g

    p = N - 1
    while p >= 32:  # ---- corner steps ----
        xc = p - 32
        x1 = XC[xc].copy()
        x2 = np.zeros(LANES, dtype=complex)
        a2 = np.zeros(LANES, dtype=complex)
        x2[:xc] = C[:xc, xc]
        d2[xc] = C[xc, xc].real
        if xc >= 1:
            a1 = XC[xc - 1].copy()
            a2[:xc] = C[:xc, xc - 1]
            a2[xc - 1] = a2[xc - 1].real
            alpha, adiag = x2[xc - 1], a2[xc - 1].real
            x2[xc - 1] = 0.0
        else:
            a1 = a[:, 31].copy()
            alpha, adiag = x1[31], a1[31].real
            x1[31] = 0.0
            a1[31] = a1[31].real
        V[:32] = x1
        V[32 : 32 + xmax] = x2[:xmax]
        y1 = a[:, : blocks(32)] @ V[: blocks(32)]
        for s in range(xc):
            y1 = y1 + XC[s] * V[32 + s]
        y2 = np.zeros(LANES, dtype=complex)
        for r in range(xc):
            y2[r] = np.sum(np.conj(XC[r]) * x1) + np.sum(C[r, :xc] * V[32 : 32 + xc])
        s1 = np.sum(np.abs(x1) ** 2) + np.sum(np.abs(x2) ** 2)
        s2 = np.sum((np.conj(x1) * y1).real) + np.sum((np.conj(x2) * y2).real)
        s3 = np.sum(np.conj(x1) * a1) + np.sum(np.conj(x2) * a2)
        beta, tau, scale, coef = scalars(alpha, s1, s2, s3, adiag)
        if xc >= 1:
            e2[xc - 1] = beta
        else:
            e1[31] = beta
        if tau == 0:
            p -= 1
            continue
        v1, v2 = x1 * scale, x2 * scale
        if xc >= 1:
            v2[xc - 1] = 1.0
        else:
            v1[31] = 1.0
        w1 = tau * (scale * y1 + a1) + coef * v1
        w2 = np.zeros(LANES, dtype=complex)
        w2[:xc] = tau * (scale * y2[:xc] + a2[:xc]) + coef * v2[:xc]
        V[:32] = v1
        W[:32] = w1
        V[32 : 32 + xmax] = v2[:xmax]
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
        ac = a[:, p - 1].copy()
        d1[p] = x[p].real
        alpha, adiag = x[p - 1], ac[p - 1].real
        x[p - 1 :] = 0.0
        ac[p - 1] = ac[p - 1].real
        V[:32] = x
        nb = blocks(p - 1)
        y = a[:, :nb] @ V[:nb]
        s1 = np.sum(np.abs(x) ** 2)
        s2 = np.sum((np.conj(x) * y).real)
        s3 = np.sum(np.conj(x) * ac)
        beta, tau, scale, coef = scalars(alpha, s1, s2, s3, adiag)
        e1[p - 1] = beta
        if tau == 0:
            p -= 1
            continue
        v = x * scale
        v[p - 1] = 1.0
        w = np.zeros(LANES, dtype=complex)
        w[:p] = tau * (scale * y[:p] + ac[:p]) + coef * v[:p]
        V[:32] = v
        W[:32] = w
        nb = blocks(p)
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
