"""numpy walk-through of the two-stage reduction implemented in tbmodels_b200/csrc/eig_band.cu (development aid).

Stage 1: Hermitian -> band (half bandwidth B = 8): per panel of 8 columns a Householder QR of the block below the band,
         Y = A22 V, the compact-WY coefficients from G = V^H V and M = V^H Y, Z = Y T - 1/2 V (T^H M T),
         A22 -= V Z^H + Z V^H.
Stage 2: band -> tridiagonal by bulge chasing with length-8 reflectors on a band array with 16 diagonals of room.

Same steps, same index conventions (band[c][d] = A[c + d, c]) as the kernels; checks the eigenvalues against numpy.
"""
import sys

import numpy as np

B = 8
BW = 16


def householder_gen(alpha, xnorm2):
    """tbk_math.cuh householder_gen: H = I - tau v v^H, v = [1; scale * x], H^H [alpha; x] = [beta; 0], beta real."""
    if xnorm2 == 0.0 and alpha.imag == 0.0:
        return alpha.real, 0.0, 0.0
    nrm = np.sqrt(alpha.real ** 2 + alpha.imag ** 2 + xnorm2)
    beta = -nrm if alpha.real >= 0 else nrm
    tau = (beta - alpha) / beta
    scale = 1.0 / (alpha - beta)
    return beta, tau, scale


def stage1(A):
    N = A.shape[0]
    A = A.copy()
    band = np.zeros((N, BW), dtype=complex)
    c0 = 0
    while N - c0 - B >= 2:
        r0 = c0 + B
        m = N - r0
        for k in range(B):  # diagonal block of the band
            for d in range(B - k):
                band[c0 + k, d] = A[c0 + k + d, c0 + k]
        E = A[r0:, c0:c0 + B].copy()  # m x 8
        R = np.zeros((B, B), dtype=complex)
        tau = np.zeros(B, dtype=complex)
        for k in range(B):
            if k >= m:
                break
            s = np.array([np.vdot(E[k + 1:, k], E[k + 1:, c]) for c in range(k, B)])
            beta, tau[k], scale = householder_gen(E[k, k], s[0].real)
            R[k, k] = beta
            for c in range(k + 1, B):
                z = E[k, c] + np.conj(scale) * s[c - k]
                R[k, c] = E[k, c] - np.conj(tau[k]) * z
                E[k + 1:, c] -= np.conj(tau[k]) * (scale * E[k + 1:, k]) * z
            E[k + 1:, k] *= scale
        V = E
        for k in range(min(B, m)):
            V[k, k:] = 0.0
            V[k, k] = 1.0
        for k in range(B):
            for i in range(k + 1):
                band[c0 + k, B + i - k] = R[i, k]
        A22 = A[r0:, r0:]
        Y = A22 @ V
        G = V.conj().T @ V
        M = V.conj().T @ Y
        T = np.zeros((B, B), dtype=complex)
        for k in range(B):
            for i in range(k):
                T[i, k] = -tau[k] * sum(T[i, l] * G[l, k] for l in range(i, k))
            T[k, k] = tau[k]
        C2 = -0.5 * T.conj().T @ (M @ T)
        Z = Y @ T + V @ C2
        A[r0:, r0:] = A22 - V @ Z.conj().T - Z @ V.conj().T
        c0 += B
    for c in range(c0, N):  # the remaining block is inside the band already
        for d in range(BW):
            if c + d < N and d <= B:
                band[c, d] = A[c + d, c]
    return band


def band_to_full(band):
    N = band.shape[0]
    A = np.zeros((N, N), dtype=complex)
    for c in range(N):
        for d in range(BW):
            if c + d < N:
                A[c + d, c] = band[c, d]
                A[c, c + d] = np.conj(band[c, d])
    for c in range(N):
        A[c, c] = band[c, 0].real
    return A


def stage2(band):
    """16 lanes per matrix in the kernel: lane r holds row r of the 16 x 8 panel P = [D; Bk] (columns R0 .. R0 + 7)."""
    N = band.shape[0]
    band = band.copy()
    d = np.zeros(N)
    e = np.zeros(N)

    def load(R0):
        P = np.zeros((16, B), dtype=complex)
        for r in range(16):
            for c in range(B):
                dd = r - c
                if 0 <= dd < BW and R0 + c < N and R0 + r < N:
                    P[r, c] = band[R0 + c, dd]
        return P

    def store(R0, P):
        for r in range(16):
            for c in range(B):
                dd = r - c
                if 0 <= dd < BW and R0 + c < N and R0 + r < N:
                    band[R0 + c, dd] = P[r, c]

    for j in range(N - 1):
        x = np.array([band[j, 1 + i] if j + 1 + i < N else 0.0 for i in range(B)], dtype=complex)
        beta, tau, scale = householder_gen(x[0], float(np.sum(np.abs(x[1:]) ** 2)))
        d[j] = band[j, 0].real
        e[j] = beta
        v = scale * x
        v[0] = 1.0
        R0 = j + 1
        while R0 < N and tau != 0.0:
            P = load(R0)
            D = P[:8]
            Bk = P[8:]
            # p = D v using the lower triangle only
            Dl = np.tril(D)
            p = Dl @ v + (np.tril(D, -1).conj().T) @ v
            p[np.arange(8)] += 0.0
            # diagonal is real by construction
            tp = tau * p
            dot = np.vdot(tp, v)  # (tau p)^H v
            w = tp - 0.5 * tau * dot * v
            for r in range(8):
                for c in range(r + 1):
                    D[r, c] -= v[r] * np.conj(w[c]) + w[r] * np.conj(v[c])
            y = Bk @ v
            Bk -= tau * np.outer(y, v.conj())
            if R0 + 8 < N:
                x2 = Bk[:, 0].copy()
                beta2, tau2, scale2 = householder_gen(x2[0], float(np.sum(np.abs(x2[1:]) ** 2)))
                v2 = scale2 * x2
                v2[0] = 1.0
                z = v2.conj() @ Bk
                Bk -= np.conj(tau2) * np.outer(v2, z)
                Bk[0, 0] = beta2
                Bk[1:, 0] = 0.0
            else:
                tau2, v2 = 0.0, v
            store(R0, P)
            R0 += 8
            v, tau = v2, tau2
    d[N - 1] = band[N - 1, 0].real
    return d, e


def tridiag_eigs(d, e):
    N = len(d)
    T = np.diag(d) + np.diag(e[:N - 1], 1) + np.diag(e[:N - 1], -1)
    return np.linalg.eigvalsh(T)


def main():
    rng = np.random.default_rng(0)
    for N in [int(a) for a in sys.argv[1:]] or [9, 10, 16, 17, 18, 25, 40, 57, 64, 100, 161]:
        X = rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))
        A = X + X.conj().T
        ref = np.linalg.eigvalsh(A)
        band = stage1(A)
        e1 = np.linalg.eigvalsh(band_to_full(band))
        d, e = stage2(band)
        e2 = tridiag_eigs(d, e)
        print(N, "stage1", np.abs(e1 - ref).max(), "stage2", np.abs(e2 - ref).max(), "max |d| beyond 8 after stage 1:",
              np.abs(band[:, 9:]).max())


if __name__ == "__main__":
    main()
