"""The oracle (oracle/tb_oracle.py) pinned against the reference: golden vectors produced by the unmodified
reference, the reference's own known-answer fixtures, and -- when /root/reference is present -- the reference
itself, live."""
import numpy as np
import pytest

from conftest import assert_eig_close, assert_h_close, load_golden, packed_from
from oracle import tb_oracle as orc
from oracle.ref_shim import import_reference, reference_available

T_COUNT, K_COUNT = 6, 4


def _check_block(d, prefix, with_h=True):
    p = packed_from(d, prefix)
    k = d[prefix + "k"]
    if with_h:
        # the oracle restates the reference operation by operation -> bit-exact on the same machine,
        # allclose across BLAS builds
        np.testing.assert_allclose(orc.hamilton(p.R, p.hop, p.pos, k, 1), d[prefix + "H1"], rtol=0, atol=1e-13)
        np.testing.assert_allclose(orc.hamilton(p.R, p.hop, p.pos, k, 2), d[prefix + "H2"], rtol=0, atol=1e-13)
    assert_eig_close(np.array(orc.eigenval(p.R, p.hop, p.pos, k)), d[prefix + "eig"], prefix)


def test_silicon_golden():
    d = load_golden("silicon.npz")
    _check_block(d, "")
    p = packed_from(d)
    assert (p.size, p.dim, p.n_R) == (8, 3, 95)
    assert_eig_close(orc.eigenval_array(p.R, p.hop, p.pos, d["k_grid"]), d["eig_grid"], "20^3 sub-grid")


def test_reference_cli_known_answer():
    """reference tests/test_cli_eigenvals.py:47-50 compares at atol 1e-10 with this very file's content."""
    d = load_golden("silicon_cli_eigenvals.npz")
    p = packed_from(d)
    got = orc.eigenval_array(p.R, p.hop, p.pos, d["k"])
    assert np.abs(got - d["eig"]).max() <= 1e-10
    assert abs(d["eig"][0, 0] - (-5.821847625730381)) < 1e-12  # value quoted in SURVEY.md 8 c3


def test_reference_regression_goldens():
    """96 + 48 goldens of tests/regression_data/test_hamilton|test_eigenval (np.allclose in the reference)."""
    from oracle import workloads as wl

    d = load_golden("ref_regression.npz")
    n = 0
    for ti, (t1, t2) in enumerate(d["t_values"]):
        p = wl.simple_model(t1, t2)
        for ki, kpt in enumerate(d["kpt"]):
            for conv in (1, 2):
                want = d[f"H{conv}_t{ti}_k{ki}"]
                got = orc.hamilton(p.R, p.hop, p.pos, kpt, conv)
                assert np.allclose(got, want) and np.abs(got - want).max() < 1e-12
                n += 1
            want = d[f"E_t{ti}_k{ki}"]
            got = orc.eigenval(p.R, p.hop, p.pos, kpt)
            assert np.allclose(got, want) and np.abs(got - want).max() < 1e-12
            n += 1
    assert n == T_COUNT * K_COUNT * 3


@pytest.mark.parametrize("dim", [2, 3, 4])
def test_simple_models(dim):
    d = load_golden("simple_models.npz")
    for ti in range(T_COUNT):
        _check_block(d, f"d{dim}_t{ti}_")


def test_haldane():
    _check_block(load_golden("haldane.npz"), "")


def test_edge_cases():
    d = load_golden("edge_cases.npz")
    for tag in ("empty_", "n1_", "d1_", "shift_"):
        _check_block(d, tag)
    p = packed_from(d, "d1_")
    np.testing.assert_allclose(orc.hamilton(p.R, p.hop, p.pos, 0.2, 1), d["d1_scalar_H1"], atol=1e-14)
    np.testing.assert_allclose(orc.eigenval(p.R, p.hop, p.pos, 0.2), d["d1_scalar_eig"], atol=1e-14)
    np.testing.assert_allclose(orc.hamilton(p.R, p.hop, p.pos, [[1], [2]], 1), d["d1_int_H1"], atol=1e-14)
    assert not packed_from(d, "empty_").n_R and np.all(d["empty_H2"] == 0)


def test_synthetic_golden():
    from oracle import workloads as wl

    d = load_golden("synthetic.npz")
    for tag in ("c3", "n3", "n5", "n7", "n12", "n17", "n33"):
        n_orb, n_half = (int(x) for x in d[f"{tag}_shape"])
        p = wl.synthetic(n_orb, n_half, seed=1234)
        k = d[f"{tag}_k"]
        nh = d[f"{tag}_H1"].shape[0]
        np.testing.assert_allclose(orc.hamilton(p.R, p.hop, p.pos, k[:nh], 1), d[f"{tag}_H1"], rtol=0, atol=1e-12)
        np.testing.assert_allclose(orc.hamilton(p.R, p.hop, p.pos, k[:nh], 2), d[f"{tag}_H2"], rtol=0, atol=1e-12)
        assert_eig_close(np.array(orc.eigenval(p.R, p.hop, p.pos, k)), d[f"{tag}_eig"], tag)


def test_c5_benchmarked_model_golden():
    """Oracle vs the reference-written golden of the benchmarked C5 model (N = 128, 1001 stored R), 8 of the 64 points."""
    from oracle import workloads as wl

    d = load_golden("c5_full.npz")
    n_orb, n_half, seed = (int(x) for x in d["shape"])
    p = wl.synthetic(n_orb, n_half, seed=seed)
    assert p.n_R == 1001
    assert_eig_close(orc.eigenval_array(p.R, p.hop, p.pos, d["k"][:8]), d["eig"][:8], "c5 full")
    np.testing.assert_allclose(orc.hamilton(p.R, p.hop, p.pos, d["k"][:2], 1), d["H1"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(orc.hamilton(p.R, p.hop, p.pos, d["k"][:2], 2), d["H2"], rtol=0, atol=1e-12)


def test_invalid_convention():
    d = load_golden("haldane.npz")
    p = packed_from(d)
    for bad in ("a", "1", None, 3):
        with pytest.raises(ValueError):
            orc.hamilton(p.R, p.hop, p.pos, (0, 0), convention=bad)


@pytest.mark.skipif(not reference_available(), reason="reference tree only exists in the build container")
def test_oracle_equals_live_reference():
    """Bit-for-bit: same numpy calls in the same order on the same machine."""
    import warnings

    from tbmodels_b200 import pack_model

    warnings.simplefilter("ignore")
    tb = import_reference()
    rng = np.random.default_rng(5)
    hop = {}
    for R in [(0, 0, 0), (1, 0, 0), (0, 1, -1), (1, -2, 3)]:
        hop[R] = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
    m = tb.Model(hop=hop, pos=rng.random((4, 3)), contains_cc=False)
    p = pack_model(m)
    k = rng.uniform(-2, 2, size=(17, 3))
    for conv in (1, 2):
        assert np.array_equal(m.hamilton(k, convention=conv), orc.hamilton(p.R, p.hop, p.pos, k, conv))
    assert np.array_equal(np.array(m.eigenval(k)), np.array(orc.eigenval(p.R, p.hop, p.pos, k)))
    assert np.array_equal(m.hamilton(k[0]), orc.hamilton(p.R, p.hop, p.pos, k[0]))


def test_kdotp_golden():
    """KdotpModel.hamilton / eigenval restatement against the reference (kdotp.py:51-100)."""
    d = load_golden("kdotp.npz")
    for tag in ("toy", "si0", "si1", "si2", "si3"):
        tc = {tuple(int(x) for x in p): c for p, c in zip(d[f"{tag}_powers"], d[f"{tag}_coeff"])}
        k = d[f"{tag}_k"]
        np.testing.assert_allclose(orc.kdotp_hamilton(tc, k), d[f"{tag}_H"], rtol=0, atol=1e-13)
        assert_eig_close(np.array(orc.kdotp_eigenval(tc, k)), d[f"{tag}_eig"], tag)
    # reference tests/test_kdotp.py:27-37 known answers
    tc = {tuple(int(x) for x in p): c for p, c in zip(d["toy_powers"], d["toy_coeff"])}
    assert np.allclose(orc.kdotp_hamilton(tc, (0, 0.5)), [[1, 0.25], [0.25, 1]])
    assert np.allclose(orc.kdotp_eigenval(tc, [(0, 0), (0, 0), (0, 0.5)]), [[1, 1], [1, 1], [0.75, 1.25]])


KDOTP_CASES = ["haldane", "silicon", "simple3d", "syn12", "syn2d", "syn36", "syn5"]


def test_construct_kdotp_golden():
    """Model.construct_kdotp restatement (reference :942-982) against the reference's output: same operations in the
    same order, so the coefficients are bit-equal; the constructed model's eigenvalues within the parity bound."""
    d = load_golden("construct_kdotp.npz")
    assert sorted(d["names"]) == KDOTP_CASES
    for name in KDOTP_CASES:
        order = int(d[f"{name}_order"])
        for i, k in enumerate(d[f"{name}_k"]):
            tc = orc.construct_kdotp(d[f"{name}_R"], d[f"{name}_hop"], d[f"{name}_pos"], k, order)
            assert np.array_equal(np.array(list(tc), dtype=np.int32), d[f"{name}_powers"])
            assert np.array_equal(np.stack(list(tc.values())), d[f"{name}_coeff{i}"]), (name, i)
            assert_eig_close(np.array(orc.kdotp_eigenval(tc, d[f"{name}_dk{i}"])), d[f"{name}_eig{i}"], name)
    with pytest.raises(ValueError):
        orc.construct_kdotp(d["syn5_R"], d["syn5_hop"], d["syn5_pos"], (0, 0, 0), -1)  # reference :961-962


@pytest.mark.skipif(not reference_available(), reason="reference tree only exists in the build container")
def test_construct_kdotp_equals_live_reference():
    import warnings

    from tbmodels_b200 import pack_model

    warnings.simplefilter("ignore")
    tb = import_reference()
    rng = np.random.default_rng(6)
    hop = {R: rng.normal(size=(3, 3)) + 1j * rng.normal(size=(3, 3)) for R in [(0, 0), (1, 0), (0, 2), (2, -3)]}
    m = tb.Model(hop=hop, pos=rng.random((3, 2)), contains_cc=False)
    p = pack_model(m)
    k = (0.37, -1.2)
    want = m.construct_kdotp(k, 3).taylor_coefficients
    got = orc.construct_kdotp(p.R, p.hop, p.pos, k, 3)
    assert list(want) == list(got)
    for key in want:
        assert np.array_equal(want[key], got[key])
    # the Taylor series reproduces H(k + dk) to the truncation order (what the reference tests check, tests/test_kdotp.py)
    dk = np.array([1e-4, -2e-4])  # fourth-order remainder ~ (2 pi |R| |dk|)^4 / 24 ~ 1e-11
    h = orc.kdotp_hamilton(got, dk)
    assert np.abs(h - m.hamilton(np.array(k) + dk)).max() < 1e-9


WANNIER_TAGS = ["hr_only_w90", "hr_only_w90v2", "hr_only_si", "hr_wsvec_si", "hr_wsvec_bi", "all_si", "all_bi", "all_bi_nearest"]


def test_reference_wannier_goldens():
    """tests/regression_data/test_wannier/* (reference tests/test_wannier.py:18-46, :182-214): [hamilton(k) for k in KPT]."""
    d = load_golden("ref_wannier.npz")
    for tag in WANNIER_TAGS:
        p = packed_from(d, tag + "_")
        got = orc.hamilton(p.R, p.hop, p.pos, d["kpt"], 2)
        want = d[f"{tag}_H2"]
        assert np.allclose(got, want) and np.abs(got - want).max() < 1e-10, tag


def test_reference_simple_model_goldens():
    """tests/regression_data/test_simple_model/* (reference tests/test_simple_model.py:12-20)."""
    from oracle import workloads as wl

    d = load_golden("ref_simple_model.npz")
    r = load_golden("ref_regression.npz")
    for ti, (t1, t2) in enumerate(r["t_values"]):
        p = wl.simple_model(t1, t2)
        for ki, kpt in enumerate(r["kpt"]):
            assert np.abs(orc.hamilton(p.R, p.hop, p.pos, kpt) - d[f"hamilton_t{ti}_k{ki}"]).max() < 1e-12
            assert np.abs(orc.eigenval(p.R, p.hop, p.pos, kpt) - d[f"eigenval_t{ti}_k{ki}"]).max() < 1e-12


def test_mesh_factorisation_algebra_matches_the_fourier_sum():
    """The identity behind tbk_eigenval_mesh (class sums over the leading mesh coordinates, then one phase per class
    along the last one), restated in numpy, against the pinned restatement of Model.hamilton on the explicit points."""
    from oracle import tb_oracle as orc
    from oracle import workloads as wl

    for p, dims, shift in (
        (wl.synthetic(5, 40, seed=1), (3, 4, 6), None),
        (wl.synthetic(3, 12, dim=2, seed=2), (5, 7), (0.5, 0.25)),
        (wl.synthetic(4, 9, dim=4, seed=3), (2, 2, 3, 4), None),
        (wl.synthetic(2, 3, dim=1, seed=4), (9,), (0.125,)),
    ):
        k = wl.kgrid_points(dims, shift)
        want = orc.hamilton(p.R, p.hop, p.pos, k, 2)
        got = orc.hamilton_mesh_factorised(p.R, p.hop, p.pos, dims, shift)
        assert got.shape == want.shape
        assert np.abs(got - want).max() <= 1e-12 * np.abs(p.hop).max() * p.n_R


@pytest.mark.parametrize("n", [1, 2, 3, 7, 8, 9, 16, 17, 40, 129])
def test_blocked_and_staged_reductions_match_lapack(n):
    """The numpy restatements of the device tridiagonalisation schedules (blocked panels for N >= 120, staged relaunch on
    the trailing block below) give LAPACK's spectrum -- incl. sizes that are not multiples of the panel width."""
    import scipy.linalg as la

    from oracle import blocked_hetrd as bh

    rng = np.random.default_rng(n)
    M = rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n))
    H = M + M.conj().T
    ref = la.eigvalsh(H)
    for d, e in (bh.blocked_tridiagonalise(H, nb=8), bh.blocked_tridiagonalise(H, nb=3), bh.staged_tridiagonalise(H)):
        got = la.eigvalsh_tridiagonal(d, e) if n > 1 else d
        assert np.abs(got - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("n", [2, 9, 10, 11, 16, 17, 18, 25, 33, 40, 57, 64, 90])
def test_twostage_reduction_matches_lapack(n):
    """The numpy restatement of the two-stage reduction of eig_band.cu (band of half bandwidth 8 by panel QR + compact-WY
    update, then bulge chasing): the band matrix and the tridiagonal matrix have LAPACK's spectrum, the first stage
    leaves the 8 diagonals of bulge room empty, and the PIPELINED order of the second stage (four sweeps in flight two
    steps apart, the kernel's schedule) touches disjoint band columns within a time step and reproduces the sequential
    sweeps bit for bit."""
    import scipy.linalg as la

    from oracle import twostage_hetrd as ts

    rng = np.random.default_rng(n)
    M = rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n))
    H = M + M.conj().T
    ref = la.eigvalsh(H)
    scale = max(1.0, np.abs(ref).max())
    band = ts.stage1(H)
    assert not band[:, ts.B + 1:].any()
    assert np.abs(la.eigvalsh(ts.band_to_full(band)) - ref).max() <= 1e-12 * scale
    d, e = ts.stage2(band)
    assert np.abs(la.eigvalsh_tridiagonal(d, e[:n - 1]) - ref).max() <= 1e-12 * scale
    dp, ep = ts.stage2_pipelined(band)
    assert np.array_equal(d, dp) and np.array_equal(e, ep)
    st = ts.sweep_start_times(n)
    for s in range(1, n - 1):
        assert st[s] >= st[s - 1] + 2  # sweep s takes step k after sweep s - 1 has finished step k + 1
        if s >= 4:
            assert st[s] >= st[s - 4] + ts.sweep_steps(n, s - 4)  # its 8-lane group is free
    # structured input: a diagonal matrix needs no reflector
    Dg = np.diag(np.arange(float(n))).astype(complex)
    d0, e0 = ts.stage2(ts.stage1(Dg))
    assert np.array_equal(np.sort(d0), np.arange(float(n))) and not e0.any()


@pytest.mark.parametrize("n", [2, 3, 5, 12, 21, 24, 25, 31, 32, 33, 34, 36, 37, 40, 41, 47, 48])
def test_register_kernel_schedule_matches_lapack(n):
    """The lane-by-lane emulation of eig_tridiag_reg.cu (index reversal, rows beyond lane 31 kept as conj(XC) + corner
    block, block-granular loops over zero-padded v / w) reproduces LAPACK's lower-storage zhetrd: same diagonal, same
    |sub-diagonal|, hence the same spectrum."""
    import scipy.linalg as la
    from scipy.linalg import lapack

    from oracle import reg_hetrd as rh

    rng = np.random.default_rng(100 + n)
    M = rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n))
    H = M + M.conj().T
    d, e = rh.reg_tridiagonalise(rh.pack_lower(H), n)
    _, d_ref, e_ref, _, info = lapack.zhetrd(H, lower=1)
    assert info == 0
    scale = np.abs(H).max() * n
    assert np.abs(d - d_ref).max() <= 1e-14 * scale
    assert np.abs(np.abs(e) - np.abs(e_ref)).max() <= 1e-14 * scale
    assert np.abs(la.eigvalsh_tridiagonal(d, e) - la.eigvalsh(H)).max() <= 1e-13 * scale
    # structured input: a diagonal matrix needs no reflector at all (tau == 0 on every step)
    Dg = np.diag(np.arange(float(n))).astype(complex)
    d0, e0 = rh.reg_tridiagonalise(rh.pack_lower(Dg), n)
    assert np.array_equal(np.sort(d0), np.arange(float(n))) and not e0.any()
