"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the TWO-STAGE Hermitian -> tridiagonal reduction that
``tbmodels_b200/csrc/eig_band.cu`` runs for N >= 128.

Stage 1 (``band_reduce_kernel``): Hermitian -> band of half bandwidth B = 8.  Per panel of 8 columns a Householder QR of
         the block below the band (one reduction per column: the norm and the dot products with the later columns
         together), Y = A22 V, the compact-WY factor T from G = V^H V and the reflector scalars (zlarft recurrence),
         M = V^H Y, Z = Y T - 1/2 V (T^H M T), A22 -= V Z^H + Z V^H.
Stage 2 (``band_chase_pipe_kernel``): band -> tridiagonal by bulge chasing with length-8 reflectors on a band array with 16
         diagonals of room, ``band[c][d] = A[c + d, c]``; ``stage2`` runs the sweeps one after the other,
         ``stage2_pipelined`` in the kernel's order: four sweeps in flight two steps apart, start times from the
         data-independent recurrence ``sweep_start_times``.

Neither exists in the reference (it calls LAPACK through scipy, src/tbmodels/_tb_model.py:1149); these functions restate
the device ALGORITHMS -- same steps, same index conventions -- so that their algebra, the room the bulges need and the
dependency analysis behind the pipelined schedule are pinned on the CPU (tests/test_oracle_golden.py).  The kernels were
written from this prototype.
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
    """Band array [N][16] of the band matrix unitarily similar to Hermitian ``A`` (lower triangle referenced)."""
    N = A.shape[0]
    A = np.array(A, dtype=complex)
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
            s = np.array([np.vdot(E[k + 1:, k], E[k + 1:, c]) for c in range(k, B)])  # ONE reduction per column
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
        if m < B:
            V[:, m:] = 0.0
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


def _load(band, R0):
    """The 16 x 8 panel [diagonal block; block below] of columns R0 .. R0 + 7 (lower triangle of the diagonal block)."""
    N = band.shape[0]
    P = np.zeros((16, B), dtype=complex)
    for r in range(16):
        for c in range(B):
            dd = r - c
            if 0 <= dd < BW and R0 + c < N and R0 + r < N:
                P[r, c] = band[R0 + c, dd]
    return P


def _store(band, R0, P):
    N = band.shape[0]
    for r in range(16):
        for c in range(B):
            dd = r - c
            if 0 <= dd < BW and R0 + c < N and R0 + r < N:
                band[R0 + c, dd] = P[r, c]


def sweep_steps(N, s):
    """Steps of sweep s: blocks R0 = s + 1, s + 9, ... < N."""
    return (N + 6 - s) >> 3


def first_reflector(band, s):
    """(beta, tau, v) of the reflector that annihilates column s below the sub-diagonal."""
    N = band.shape[0]
    x = np.array([band[s, 1 + i] if s + 1 + i < N else 0.0 for i in range(B)], dtype=complex)
    beta, tau, scale = householder_gen(x[0], float(np.sum(np.abs(x[1:]) ** 2)))
    v = scale * x
    v[0] = 1.0
    return beta, tau, v


def chase_step(band, R0, v, tau):
    """One step of a sweep on the panel at R0 with the reflector (v, tau) of rows R0 .. R0 + 7: two-sided update of the
    diagonal block, right update of the block below, the next reflector (annihilates column 0 of the block below) and
    its left update.  Returns the next (v, tau)."""
    P = _load(band, R0)
    D = P[:8]
    Bk = P[8:]
    p = np.tril(D) @ v + np.tril(D, -1).conj().T @ v  # Hermitian product from the lower triangle
    tp = tau * p
    dot = np.vdot(tp, v)  # (tau p)^H v
    w = tp - 0.5 * tau * dot * v
    for r in range(8):
        for c in range(r + 1):
            D[r, c] -= v[r] * np.conj(w[c]) + w[r] * np.conj(v[c])
        D[r, r] = D[r, r].real
    y = Bk @ v
    Bk -= tau * np.outer(y, v.conj())
    x2 = Bk[:, 0].copy()
    beta2, tau2, scale2 = householder_gen(x2[0], float(np.sum(np.abs(x2[1:]) ** 2)))
    v2 = scale2 * x2
    v2[0] = 1.0
    z = v2.conj() @ Bk
    Bk -= np.conj(tau2) * np.outer(v2, z)
    Bk[0, 0] = beta2
    Bk[1:, 0] = 0.0
    _store(band, R0, P)
    return v2, tau2


def stage2(band):
    """d [N], e [N]: the sweeps one after the other."""
    N = band.shape[0]
    band = band.copy()
    d = np.zeros(N)
    e = np.zeros(N)
    for s in range(N - 1):
        d[s] = band[s, 0].real
        e[s], tau, v = first_reflector(band, s)
        for k in range(sweep_steps(N, s)):
            v, tau = chase_step(band, s + 1 + B * k, v, tau)
    d[N - 1] = band[N - 1, 0].real
    return d, e


def sweep_start_times(N):
    """Start time (in steps of the lockstep loop) of every sweep: four 8-lane groups, group g runs sweeps g, g + 4, ...;
    sweep s starts when its group is free and sweep s - 1 is two steps ahead:
    start(s) = max(start(s - 1) + 2, start(s - 4) + steps(s - 4))."""
    st = []
    for s in range(N - 1):
        if s < 4:
            st.append(2 * s)
        else:
            st.append(max(st[s - 1] + 2, st[s - 4] + sweep_steps(N, s - 4)))
    return st


def stage2_pipelined(band):
    """The same sweeps in the order of ``band_chase_pipe_kernel``: at time t group g performs step t - start(s) of its
    current sweep s; within one time step the groups touch disjoint columns, so their order does not matter."""
    N = band.shape[0]
    band = band.copy()
    d = np.zeros(N)
    e = np.zeros(N)
    st = sweep_start_times(N)
    state = {}  # sweep -> (v, tau)
    cur = [g if g < N - 1 else None for g in range(4)]
    t = 0
    while any(s is not None for s in cur):
        touched = []
        for g in (3, 1, 0, 2):  # any order
            s = cur[g]
            if s is None or t < st[s]:
                continue
            k = t - st[s]
            if k == 0:
                d[s] = band[s, 0].real
                beta, tau, v = first_reflector(band, s)
                e[s] = beta
                state[s] = (v, tau)
            R0 = s + 1 + B * k
            touched.append((R0, R0 + B - 1))
            state[s] = chase_step(band, R0, *state[s])
            if k + 1 == sweep_steps(N, s):
                del state[s]
                cur[g] = s + 4 if s + 4 < N - 1 else None
        touched.sort()
        for (a0, a1), (b0, b1) in zip(touched, touched[1:]):
            assert a1 < b0, "two sweeps touched the same band columns in one time step"
        t += 1
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
        dp, ep = stage2_pipelined(band)
        e2 = tridiag_eigs(d, e)
        print(N, "stage1", np.abs(e1 - ref).max(), "stage2", np.abs(e2 - ref).max(), "pipelined == sequential:",
              np.array_equal(d, dp) and np.array_equal(e, ep))


if __name__ == "__main__":
    main()
