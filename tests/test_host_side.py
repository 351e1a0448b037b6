"""CPU-only tests: host logic, the scalar device code run on the host, the C-ABI surface, workloads."""
import ctypes as C
import os
import pickle
import re

import numpy as np
import pytest
import scipy.linalg as la

from conftest import ROOT, load_golden, packed_from

dp = C.POINTER(C.c_double)


def _lib():
    from tbmodels_b200 import _capi

    return _capi.load()


def hp_pack(H):
    n = H.shape[0]
    out = np.zeros(n * n)
    nre = n * (n + 1) // 2
    for i in range(n):
        for j in range(i + 1):
            out[i * (i + 1) // 2 + j] = H[i, j].real
            if j < i:
                out[nre + i * (i - 1) // 2 + j] = H[i, j].imag
    return out


def test_library_exports_every_declared_symbol():
    from tbmodels_b200 import _capi

    lib = _lib()
    with open(os.path.join(ROOT, "include", "tbk.h")) as fh:
        header = fh.read()
    declared = set(re.findall(r"\b(tbk_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_capi.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.tbk_version() >= 100


def test_compute_entry_points_fail_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import tbmodels_b200 as tbk

    with pytest.raises(tbk.TbkError, match="no CPU fallback"):
        tbk.KModel.from_packed(packed_from(load_golden("haldane.npz"))).eigenval([0.0, 0.0])


@pytest.mark.parametrize("n", [1, 2, 3, 4, 7, 8, 13, 36, 64, 97])
def test_host_hetrd_and_ql_match_lapack(n):
    """The exact scalar code of the kernels (tbk_math.cuh), executed on the CPU."""
    lib = _lib()
    rng = np.random.default_rng(n)
    for trial in range(6):
        A = rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n))
        H = A + A.conj().T
        if trial == 3:
            H = np.diag(rng.normal(size=n)).astype(complex)
        if trial == 4 and n > 2:  # two n/2-fold degenerate levels
            Q, _ = np.linalg.qr(A)
            H = (Q * np.repeat([1.0, 2.0], [n // 2, n - n // 2])) @ Q.conj().T
            H = (H + H.conj().T) / 2
        if trial == 5:
            H = H * 1e-3 + np.diag(np.arange(n) * 10.0)  # graded
        ref = la.eigvalsh(H)
        hp = hp_pack(H)
        d = np.zeros(n)
        e = np.zeros(n)
        assert lib.tbk_host_hetrd(n, hp.ctypes.data_as(dp), d.ctypes.data_as(dp), e.ctypes.data_as(dp)) == 0
        if n > 1:
            assert np.abs(la.eigvalsh_tridiagonal(d, e[: n - 1]) - ref).max() <= 1e-13 * max(1, np.abs(ref).max())
        db = d.copy()
        assert lib.tbk_host_tridiag_bisect(n, db.ctypes.data_as(dp), e.ctypes.data_as(dp)) == 0
        assert np.all(np.diff(db) >= 0)
        assert np.abs(db - ref).max() <= 1e-13 * max(1, np.abs(ref).max())
        assert lib.tbk_host_tridiag_ql(n, d.ctypes.data_as(dp), e.ctypes.data_as(dp)) == 0
        assert np.all(np.diff(d) >= 0)
        assert np.abs(d - ref).max() <= 1e-13 * max(1, np.abs(ref).max())


def test_hermitian_split_weights_reproduce_reference_sum():
    """W rows (T + T^H, i(T - T^H)) with [cos | sin] coefficients == sum_R e^{2 pi i k.R} T_R + h.c. (:1117-1123)."""
    from oracle import tb_oracle as orc

    lib = _lib()
    d = load_golden("silicon.npz")
    p = packed_from(d)
    n, nR = p.size, p.n_R
    W = np.zeros((2 * nR, n * n))
    assert lib.tbk_host_pack_weights(n, nR, p.hop.view(np.float64).ctypes.data_as(dp), W.ctypes.data_as(dp)) == 0
    for k in d["k"][:8]:
        x = p.R @ k
        Q = np.empty(2 * nR)
        Q[0::2] = np.cos(2 * np.pi * x)
        Q[1::2] = np.sin(2 * np.pi * x)
        want = hp_pack(orc.hamilton(p.R, p.hop, p.pos, k, 2))
        assert np.abs(Q @ W - want).max() <= 1e-12


def test_pack_model_dense_sparse_and_zero_matrices():
    import scipy.sparse as sp

    import tbmodels_b200 as tbk

    class M:
        pass

    rng = np.random.default_rng(0)
    m = M()
    m.size, m.dim = 3, 2
    m.pos = rng.random((3, 2))
    a = rng.normal(size=(3, 3)) + 1j * rng.normal(size=(3, 3))
    m.hop = {(0, 0): a, (1, -1): sp.csr_matrix(a * 2), (0, 1): np.zeros((3, 3), complex)}
    p = tbk.pack_model(m)
    assert p.n_R == 2 and p.R.tolist() == [[0, 0], [1, -1]]
    assert np.array_equal(p.hop[1], a * 2)
    q = tbk.pack_arrays(p.R, p.hop, p.pos)
    assert q.digest() == p.digest()
    m.hop[(1, -1)] = sp.csr_matrix(a * 3)
    assert tbk.pack_model(m).digest() != p.digest()
    assert set(tbk.hop_dict(p)) == {(0, 0), (1, -1)}
    with pytest.raises(ValueError):
        bad = M()
        bad.size, bad.dim, bad.pos, bad.hop = 3, 2, m.pos, {(0, 0, 0): a}
        tbk.pack_model(bad)


def test_k_normalisation_matches_reference_rules():
    from tbmodels_b200._evaluator import _check_convention, _normalise_k

    k, single = _normalise_k(0.2, 1)
    assert single and k.shape == (1, 1)
    k, single = _normalise_k((0.1, 0.2, 0.7), 3)
    assert single and k.shape == (1, 3)
    k, single = _normalise_k([[1, 2, 3]], 3)
    assert not single and k.dtype == np.float64
    k, single = _normalise_k(np.zeros((0, 3)), 3)
    assert not single and k.shape == (0, 3)
    with pytest.raises(ValueError):
        _normalise_k((0.1, 0.2), 3)
    for bad in ("a", "1", None, 0, 3):
        with pytest.raises(ValueError):
            _check_convention(bad)
    _check_convention(1)
    _check_convention(2)


def test_kmodel_pickles_without_device_state():
    import tbmodels_b200 as tbk

    m = tbk.KModel.from_packed(packed_from(load_golden("haldane.npz")))
    m._cache = ("digest", object())
    m2 = pickle.loads(pickle.dumps(m))
    assert m2._cache is None and set(m2.hop) == set(m.hop)


def test_workload_generators_golden():
    from oracle import workloads as wl

    d = load_golden("haldane.npz")
    h = wl.haldane()
    assert np.array_equal(h.R, d["R"]) and np.array_equal(h.hop, d["hop"]) and np.array_equal(h.pos, d["pos"])
    s = load_golden("simple_models.npz")
    t_values = load_golden("ref_regression.npz")["t_values"]
    for dim in (2, 3, 4):
        for ti, (t1, t2) in enumerate(t_values):
            p = wl.simple_model(t1, t2, dim=dim)
            assert np.array_equal(p.R, s[f"d{dim}_t{ti}_R"]) and np.array_equal(p.hop, s[f"d{dim}_t{ti}_hop"])
    sc = load_golden("supercell.npz")
    si = packed_from(load_golden("silicon.npz"))
    for tag, size in (("s222", (2, 2, 2)), ("s444", (4, 4, 4))):
        p = wl.supercell(si, size)
        assert np.array_equal(p.R, sc[f"{tag}_R"]) and np.allclose(p.pos, sc[f"{tag}_pos"], atol=0, rtol=0)
        assert np.allclose([np.abs(p.hop).sum(), np.abs(p.hop).max()], sc[f"{tag}_hop_absum"], rtol=1e-14)
    assert wl.supercell(si, (4, 4, 4)).size == 512
    c3 = wl.synthetic(36, 250)
    assert (c3.size, c3.n_R) == (36, 251)
    # hermitian-consistent by construction: R = 0 block is (half of) a Hermitian matrix
    assert np.abs(c3.hop[0] - c3.hop[0].conj().T).max() == 0
    g = wl.kgrid(4, 3)
    assert g.shape == (64, 3) and g[1].tolist() == [0, 0, 0.25]
    v = wl.shortest_half_vectors(13)
    assert len({tuple(x) for x in v}) == 13 and all(next(c for c in x if c != 0) > 0 for x in v)


def test_kdotp_mirror_rejects_non_hermitian_and_packs_in_dict_order():
    import tbmodels_b200 as tbk

    with pytest.raises(ValueError, match="not hermitian"):
        tbk.KdotpModel({(0, 0): [[0, 1], [2, 0]]})
    m = tbk.KdotpModel({(0, 0): np.eye(2), (1, 0): [[0, 1j], [-1j, 0]], (0, 2): [[0, 1], [1, 0]]})
    powers, coeff = tbk.pack_kdotp(m.taylor_coefficients)
    assert powers.tolist() == [[0, 0], [1, 0], [0, 2]] and coeff.shape == (3, 2, 2)
    assert pickle.loads(pickle.dumps(m))._cache is None


def test_sincospi_lean_accuracy_in_ulps():
    """The device phase function (tbk_math.cuh sincospi_lean, compiled for the host): <= 1.5 ulp against an 80-bit
    reference after the exact argument reduction, exact at multiples of 1/2, large arguments keep their quadrant."""
    import ctypes as C

    lib = _lib()
    ld = np.longdouble
    if np.finfo(ld).nmant < 63:
        pytest.skip("no extended-precision long double on this platform")
    pi = ld("3.14159265358979323846264338327950288")

    def sc(t):
        s, c = C.c_double(), C.c_double()
        assert lib.tbk_host_sincospi(float(t), C.byref(s), C.byref(c)) == 0
        return s.value, c.value

    rng = np.random.default_rng(0)
    ts = np.concatenate([rng.uniform(-4, 4, 20000), rng.uniform(-1e6, 1e6, 4000), np.linspace(-0.25, 0.25, 2001),
                         np.linspace(0.24, 0.26, 1001)])
    worst = 0.0
    for t in ts:
        s, c = sc(t)
        n = np.rint(2 * t)
        r = ld(t) - ld(n) / 2
        sv, cv = np.sin(pi * r), np.cos(pi * r)
        rs, rc = [(sv, cv), (cv, -sv), (-sv, -cv), (-cv, sv)][int(n) % 4]
        for got, ref in ((s, rs), (c, rc)):
            if ref != 0:
                worst = max(worst, float(abs(ld(got) - ref) / np.spacing(abs(np.float64(ref)))))
    assert worst <= 1.5, worst
    assert sc(0.5) == (1.0, 0.0) and sc(1.0)[1] == -1.0 and sc(-0.5)[0] == -1.0 and sc(0.0) == (0.0, 1.0)
    assert sc(2.0 ** 40 + 0.5) == (1.0, 0.0) and abs(sc(0.25)[0] - np.sqrt(0.5)) < 2e-16


def test_kdotp_power_tuples_follow_the_reference_order():
    """Evaluator.construct_kdotp returns its coefficients under the power tuples of the reference's dict
    (src/tbmodels/_tb_model.py:963-966: itertools.product(range(order + 1), repeat=dim) filtered by sum <= order)."""
    from tbmodels_b200._evaluator import kdotp_powers

    d = load_golden("construct_kdotp.npz")
    for name in d["names"]:
        order = int(d[f"{name}_order"])
        dim = d[f"{name}_k"].shape[1]
        assert np.array_equal(kdotp_powers(dim, order), d[f"{name}_powers"]), name
    assert kdotp_powers(3, 0).tolist() == [[0, 0, 0]]
    assert kdotp_powers(1, 3).ravel().tolist() == [0, 1, 2, 3]


def test_supercell_duck_type_host_side():
    """SupercellKModel: sizes, positions (reference :1670-1678) and argument errors (:1656-1662) without touching a GPU;
    pickling drops the device handle."""
    import pickle

    import tbmodels_b200 as tbk
    from oracle import workloads as wl

    base = wl.synthetic(3, 4, seed=5)
    sup = tbk.SupercellKModel(base, (2, 1, 3))
    big = wl.supercell(base, (2, 1, 3))
    assert sup.size == big.size == 18 and sup.dim == 3
    assert np.array_equal(sup.pos, big.pos)
    clone = pickle.loads(pickle.dumps(sup))
    assert clone.size == 18 and clone._ev is None and np.array_equal(clone.pos, sup.pos)
    with pytest.raises(ValueError):
        tbk.SupercellKModel(base, (2, 2))
    with pytest.raises(ValueError):
        tbk.SupercellKModel(base, (2, 0, 1))
    with pytest.raises(ValueError):
        sup.hamilton((0, 0, 0), convention=3)  # before any device work


def test_new_entry_points_fail_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import tbmodels_b200 as tbk
    from oracle import workloads as wl

    with pytest.raises(tbk.TbkError) as exc:
        tbk.Evaluator.from_supercell(wl.synthetic(3, 4, seed=5), (2, 1, 1))
    assert "no CPU fallback" in str(exc.value)
