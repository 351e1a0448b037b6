"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the BLOCKED Hermitian -> tridiagonal reduction that
``tbmodels_b200/csrc/eig_tridiag_panel.cu`` runs for N >= 120 (panels of NB columns, the trailing matrix updated once per
panel), and of the STAGED form of the shared-memory kernels (``eig_tridiag.cu``: stop after some steps, continue on the
trailing block as an independent smaller problem).

Neither exists in the reference (it calls LAPACK through scipy, src/tbmodels/_tb_model.py:1149); these functions restate
the device ALGORITHMS so that their algebra -- the column update with the panel's V / W, the corrected product
``p = A v - V (W^H v) - W (V^H v)``, the deferred rank-2NB update, the last (partial) panel, the stage hand-over -- is
pinned against LAPACK on the CPU (tests/test_oracle_golden.py).  The kernels were written from this prototype.
"""
from __future__ import annotations

import numpy as np


def reflector(alpha: complex, xnorm2: float):
    """(beta, tau, scale) with the contract of tbk_math.cuh ``householder_gen`` (LAPACK zlarfg restated)."""
    ar, ai = alpha.real, alpha.imag
    if xnorm2 == 0.0 and ai == 0.0:
        return ar, 0j, 0j
    nrm = np.sqrt(ar * ar + ai * ai + xnorm2)
    beta = -nrm if ar >= 0 else nrm
    return beta, complex((beta - ar) / beta, -ai / beta), 1.0 / (alpha - beta)


def blocked_tridiagonalise(A: np.ndarray, nb: int = 8):
    """Diagonal d [N] and sub-diagonal e [N-1] of the tridiagonal matrix unitarily similar to Hermitian ``A`` (only the
    lower triangle is referenced), computed panel by panel like ``tridiag_panel_kernel``."""
    A = np.array(A, dtype=complex)
    N = A.shape[0]
    d = np.zeros(N)
    e = np.zeros(max(N - 1, 0))
    k0 = 0
    while k0 < N:
        w_cols = min(nb, N - k0)
        V = np.zeros((N, w_cols), dtype=complex)
        W = np.zeros((N, w_cols), dtype=complex)
        for j in range(w_cols):
            c = k0 + j
            a = A[c:, c].copy()
            a[0] = a[0].real
            for p in range(j):  # column c brought up to date with the panel's previous reflectors
                a -= V[c:, p] * np.conj(W[c, p]) + W[c:, p] * np.conj(V[c, p])
            d[c] = a[0].real
            if c == N - 1:
                break
            beta, tau, scale = reflector(a[1], float(np.sum(np.abs(a[2:]) ** 2)))
            e[c] = beta
            v = np.zeros(N, dtype=complex)
            v[c + 1] = 1.0
            v[c + 2:] = a[2:] * scale
            V[:, j] = v
            T = A[c + 1:, c + 1:]  # the STORED trailing block (not updated inside the panel), lower triangle only
            L = np.tril(T, -1)
            Tfull = L + L.conj().T + np.diag(np.diag(T).real)
            p_vec = np.zeros(N, dtype=complex)
            p_vec[c + 1:] = Tfull @ v[c + 1:]
            p_vec -= V[:, :j] @ (W[:, :j].conj().T @ v) + W[:, :j] @ (V[:, :j].conj().T @ v)
            w = tau * p_vec
            w += -0.5 * tau * np.vdot(w, v) * v
            W[:, j] = w
        r = k0 + w_cols
        if r < N:  # deferred rank-2 NB update of the trailing block (the tensor-core her2k of the kernel)
            A[r:, r:] -= np.tril(V[r:] @ W[r:].conj().T + W[r:] @ V[r:].conj().T)
        k0 += w_cols
    return d, e


def unblocked_steps(A: np.ndarray, nsteps: int):
    """``nsteps`` unblocked Householder steps on the lower triangle of ``A``: (d[:nsteps], e[:nsteps], trailing block) --
    one STAGE of the shared-memory kernels; the trailing block is an independent Hermitian problem."""
    A = np.array(A, dtype=complex)
    N = A.shape[0]
    nst = min(nsteps, N - 1)
    d, e = np.zeros(nst), np.zeros(nst)
    for j in range(nst):
        d[j] = A[j, j].real
        x = A[j + 1:, j]
        beta, tau, scale = reflector(x[0], float(np.sum(np.abs(x[1:]) ** 2)))
        e[j] = beta
        v = np.concatenate([[1.0], x[1:] * scale])
        T = A[j + 1:, j + 1:]
        L = np.tril(T, -1)
        Tfull = L + L.conj().T + np.diag(np.diag(T).real)
        p_vec = tau * (Tfull @ v)
        w = p_vec - 0.5 * tau * np.vdot(p_vec, v) * v
        A[j + 1:, j + 1:] = Tfull - np.outer(v, w.conj()) - np.outer(w, v.conj())
    return d, e, A[nst:, nst:]


def staged_tridiagonalise(A: np.ndarray, ratio: float = 0.67):
    """The staged schedule of ``launch_tridiag`` (sizes N -> ratio N -> ... until <= 16) on top of :func:`unblocked_steps`."""
    cur = np.array(A, dtype=complex)
    ds, es = [], []
    while True:
        n = cur.shape[0]
        nxt = int(n * ratio + 0.5) if n > 16 else 0
        if nxt and nxt < 12:
            nxt = 12
        if not nxt or nxt >= n:
            d, e, rest = unblocked_steps(cur, n - 1)
            ds.append(d)
            es.append(e)
            ds.append(np.array([rest[0, 0].real]) if n >= 1 else np.zeros(0))
            break
        d, e, cur = unblocked_steps(cur, n - nxt)
        ds.append(d)
        es.append(e)
    return np.concatenate(ds), np.concatenate(es)
