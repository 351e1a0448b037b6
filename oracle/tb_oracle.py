"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy/scipy) of the TBmodels k-space hot path.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks every function here against
(a) outputs of the unmodified reference run in the build container (``tests/golden/*.npz``, made by
``oracle/make_golden.py``), (b) the reference's own known-answer file
``tests/samples/cli_eigenvals/silicon_eigenvals.hdf5`` (values committed in
``tests/golden/silicon_cli_eigenvals.npz``) and (c) the reference's regression goldens
``tests/regression_data/test_hamilton|test_eigenval`` (subset committed in
``tests/golden/ref_regression.npz``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module.  The product package ``tbmodels_b200`` must never import it.

The functions operate on the *packed* form of a model (what ``tbmodels_b200.pack_model`` produces):

* ``R``   int   [n_R, dim]      -- the keys of ``Model.hop`` (half set: first non-zero component > 0, or 0)
* ``hop`` c128  [n_R, N, N]     -- the dense values of ``Model.hop`` (R = 0 entry is HALF the on-site block,
                                   reference src/tbmodels/_tb_model.py:218,268)
* ``pos`` f64   [N, dim]        -- ``Model.pos``
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as la


def hamilton(R, hop, pos, k, convention=2):
    """Restates ``Model.hamilton`` (reference src/tbmodels/_tb_model.py:1076-1132) operation by operation.

    Same loop order over the stored R vectors, same ``exp(2j*pi*dot(k, R))`` phase (argument rounded in
    f64 before the exponential, :1118), same ``H += H^dagger`` (:1123) and the orbital-position phases for
    ``convention == 1`` only (:1124-1128).
    """
    if convention not in [1, 2]:  # :1097-1102
        raise ValueError(
            "Invalid value '{}' for 'convention': must be either '1' or '2'".format(convention)
        )
    R = np.asarray(R)
    hop = np.asarray(hop)
    pos = np.asarray(pos)
    size = pos.shape[0]
    k_array = np.array(k, ndmin=1)  # :1103
    if k_array.ndim == 1:
        single_point = True
        k_array = k_array.reshape((1, -1))
    else:
        single_point = False
    H = np.zeros((k_array.shape[0], size, size), dtype=complex)  # :1109
    tmp_array = np.empty_like(H)
    for r_vec, mat in zip(R, hop):  # :1111-1122
        np.multiply(
            np.exp(2j * np.pi * np.dot(k_array, r_vec)).reshape((-1, 1, 1)),
            mat[np.newaxis, :, :],
            out=tmp_array,
        )
        H += tmp_array
    H += H.conjugate().transpose((0, 2, 1))  # :1123
    if convention == 1:  # :1124-1128
        pos_exponential = np.array(
            [[np.exp(2j * np.pi * np.dot(k_array, p)) for p in pos]]
        ).transpose((2, 0, 1))
        H = pos_exponential.conjugate().transpose((0, 2, 1)) * H * pos_exponential
    if single_point:  # :1130-1132
        return H[0]
    return H


def eigenval(R, hop, pos, k):
    """Restates ``Model.eigenval`` (reference src/tbmodels/_tb_model.py:1134-1150).

    ``hamilton`` with the default convention 2, then one ``scipy.linalg.eigvalsh`` (LAPACK, lower
    triangle, ascending) per k-point in a Python loop; returns a list of arrays for a k-list and a single
    array for a single k-point.
    """
    hamiltonians = hamilton(R, hop, pos, k)
    if hamiltonians.ndim == 3:
        return [la.eigvalsh(ham) for ham in hamiltonians]
    return la.eigvalsh(hamiltonians)


def eigenval_array(R, hop, pos, k, chunk=4096):
    """``eigenval`` for a k-list, chunked so the two [n_k, N, N] temporaries of the reference stay small.

    Returns one ``[n_k, N]`` array.  Chunking does not change any number (every k-point is independent,
    reference :1109-1132); it only bounds memory, which the reference cannot do at the benchmark sizes.
    """
    k = np.asarray(k, dtype=float)
    assert k.ndim == 2
    size = np.asarray(pos).shape[0]
    out = np.empty((k.shape[0], size))
    for start in range(0, k.shape[0], chunk):
        ev = eigenval(R, hop, pos, k[start : start + chunk])
        if len(ev):
            out[start : start + chunk] = np.asarray(ev)
    return out


def eigh(R, hop, pos, k):
    """Eigenvalues AND eigenvectors, the step after the path: ``scipy.linalg.eigh`` (LAPACK zheevr,
    JOBZ='V', UPLO='L') on every convention-2 ``hamilton(k)[i]`` exactly where ``eigenval`` (reference :1147-1149) calls
    ``eigvalsh``.  Returns ``(w [n_k, N], v [n_k, N, N])`` (or the squeezed pair for a single k-point)."""
    hamiltonians = hamilton(R, hop, pos, k)
    if hamiltonians.ndim == 3:
        pairs = [la.eigh(ham) for ham in hamiltonians]
        size = hamiltonians.shape[1]
        return (np.array([p[0] for p in pairs]).reshape(len(pairs), size),
                np.array([p[1] for p in pairs]).reshape(len(pairs), size, size))
    return la.eigh(hamiltonians)


def construct_kdotp(R, hop, pos, k, order):
    """Restates ``Model.construct_kdotp`` (reference src/tbmodels/_tb_model.py:942-982) on the packed arrays: the
    ``taylor_coefficients`` dict ``{power tuple: matrix}`` of the k.p expansion of the convention-2 Hamiltonian at ``k``,
    every operation in the reference's order (``pos`` is not used: convention 2)."""
    import itertools

    from scipy.special import factorial

    if order < 0:
        raise ValueError("The order for the k.p model must be positive.")
    R = np.asarray(R)
    hop = np.asarray(hop)
    size = hop.shape[1] if hop.ndim == 3 and hop.shape[0] else np.asarray(pos).shape[0]
    dim = R.shape[1] if R.ndim == 2 and R.shape[0] else np.asarray(pos).shape[1]
    taylor_coefficients = dict()
    for k_powers in itertools.product(range(order + 1), repeat=dim):
        curr_order = sum(k_powers)
        if curr_order > order:
            continue
        taylor_coefficients[k_powers] = ((2j * np.pi) ** curr_order / np.prod(factorial(k_powers, exact=True))) * sum(
            (
                np.prod(np.array(Rv) ** np.array(k_powers)) * np.exp(2j * np.pi * np.dot(k, Rv)) * mat
                + np.prod((-np.array(Rv)) ** np.array(k_powers)) * np.exp(-2j * np.pi * np.dot(k, Rv)) * mat.T.conj()
                for Rv, mat in zip(R, hop)
            ),
            np.zeros((size, size), dtype=complex),
        )
    return taylor_coefficients


def kdotp_hamilton(taylor_coefficients, k):
    """Restates ``KdotpModel.hamilton`` (reference src/tbmodels/kdotp.py:51-82): sum over the Taylor terms of
    ``prod(k**powers) * C`` in dict order."""
    k_array = np.array(k, ndmin=1)
    if k_array.ndim == 1:
        single_point = True
        k_array = k_array.reshape((1, -1))
    else:
        single_point = False
    ham = sum(
        np.prod(k_array**k_powers, axis=-1).reshape(-1, 1, 1) * np.asarray(mat, dtype=complex)[np.newaxis, :, :]
        for k_powers, mat in taylor_coefficients.items()
    )
    if single_point:
        return ham[0]
    return ham


def kdotp_eigenval(taylor_coefficients, k):
    """Restates ``KdotpModel.eigenval`` (reference src/tbmodels/kdotp.py:84-100)."""
    hamiltonians = kdotp_hamilton(taylor_coefficients, k)
    if hamiltonians.ndim == 3:
        return [la.eigvalsh(ham) for ham in hamiltonians]
    return la.eigvalsh(hamiltonians)


def hamilton_longdouble(R, hop, pos, k, convention=2):
    """Same formula evaluated with 80-bit phases (argument reduced exactly); used only to show which of two
    f64 implementations is closer to the exact answer when they disagree at the 1e-15 level."""
    R = np.asarray(R)
    hop = np.asarray(hop)
    pos = np.asarray(pos, dtype=np.longdouble)
    k_array = np.array(k, ndmin=2, dtype=np.longdouble)
    size = pos.shape[0]
    H = np.zeros((k_array.shape[0], size, size), dtype=np.clongdouble)
    pi_l = np.longdouble("3.14159265358979323846264338327950288")
    for r_vec, mat in zip(R, hop):
        x = k_array @ r_vec.astype(np.longdouble)
        x = x - np.rint(x)
        ph = np.cos(2 * pi_l * x) + 1j * np.sin(2 * pi_l * x)
        H += ph.reshape((-1, 1, 1)) * mat[None].astype(np.clongdouble)
    H = H + H.conjugate().transpose((0, 2, 1))
    if convention == 1:
        x = k_array @ pos.T  # [nk, N]
        x = x - np.rint(x)
        pe = np.cos(2 * pi_l * x) + 1j * np.sin(2 * pi_l * x)
        H = pe.conjugate()[:, :, None] * H * pe[:, None, :]
    return H


def hamilton_mesh_factorised(R, hop, pos, dims, shift=None):
    """Convention-2 H(k) on the regular mesh ``k_d = (i_d + shift_d) / dims[d]`` ('ij' order), evaluated the way
    ``tbk_eigenval_mesh`` does it (tbmodels_b200/csrc/hk_mesh.cu): stored R sorted into classes by their last component,
    per mesh line the class sums ``G_c = sum_{r in c} e^{2 pi i kappa.R_r} T_r`` over the leading coordinates, then
    ``H = sum_c e^{2 pi i k_z z_c} G_c + h.c.`` along the line.  A restatement of the device algorithm's ALGEBRA (to be
    compared with :func:`hamilton` on the explicit mesh points), not of reference code: the reference has no mesh path.
    """
    R = np.asarray(R)
    hop = np.asarray(hop)
    dims = [int(n) for n in dims]
    dim = len(dims)
    size = np.asarray(pos).shape[0]
    shift = [0.0] * dim if shift is None else [float(x) for x in shift]
    axes = [(np.arange(n) + shift[d]) / n for d, n in enumerate(dims)]
    kz = axes[-1]
    lines = np.stack(np.meshgrid(*axes[:-1], indexing="ij"), axis=-1).reshape(-1, dim - 1) if dim > 1 else np.zeros((1, 0))
    classes = sorted(set(int(z) for z in R[:, -1])) if len(R) else []
    H = np.zeros((lines.shape[0], dims[-1], size, size), dtype=complex)
    for z in classes:
        sel = R[:, -1] == z
        phases = np.exp(2j * np.pi * (lines @ R[sel, :-1].T))              # [n_lines, n_r]
        G = np.einsum("lr,rij->lij", phases, hop[sel])                     # [n_lines, N, N]
        H += np.exp(2j * np.pi * kz * z)[None, :, None, None] * G[:, None, :, :]
    H = H.reshape(-1, size, size)
    return H + H.conjugate().transpose((0, 2, 1))
