"""Parity of the CUDA path (through the C ABI) with the oracle and the reference's golden vectors.

Bounds (BASELINE.json north_star): H(k): max|dH| <= 1e-11 * max|H_R|; eigenvalues: <= 1e-10 * spectral radius.
"""
import os

import numpy as np
import pytest

from conftest import assert_eig_close, assert_h_close, h_scale, load_golden, packed_from

pytestmark = pytest.mark.gpu
ROOT_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

T_COUNT = 6


@pytest.fixture(scope="module")
def tbk():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import tbmodels_b200

    return tbmodels_b200


def _oracle():
    from oracle import tb_oracle

    return tb_oracle


def _check(tbk, packed, k, H1=None, H2=None, eig=None, what=""):
    ev = tbk.Evaluator(packed)
    try:
        if H1 is not None:
            assert_h_close(ev.hamilton(k[: len(H1)], convention=1), H1, packed, what + " H conv1")
        if H2 is not None:
            got = ev.hamilton(k[: len(H2)], convention=2)
            assert_h_close(got, H2, packed, what + " H conv2")
            # structural facts of the reference output: bit-wise Hermitian, real diagonal
            assert np.array_equal(got, got.conj().transpose(0, 2, 1)), what + ": conv-2 H not bit-wise Hermitian"
            assert np.all(np.einsum("kii->ki", got).imag == 0.0)
        if eig is not None:
            got = ev.eigenval_array(k)
            assert_eig_close(got, eig, what + " eig")
            assert np.all(np.diff(got, axis=1) >= 0), what + ": eigenvalues not ascending"
    finally:
        ev.close()


def _block(tbk, d, prefix):
    p = packed_from(d, prefix)
    _check(tbk, p, d[prefix + "k"], d[prefix + "H1"], d[prefix + "H2"], d[prefix + "eig"], prefix or "model")


# ----------------------------------------------------------------------------------- golden vectors
def test_silicon_c1(tbk):
    d = load_golden("silicon.npz")
    _block(tbk, d, "")
    p = packed_from(d)
    ev = tbk.Evaluator(p)
    assert ev.path == "fused-small"
    assert_eig_close(ev.eigenval_array(d["k_grid"]), d["eig_grid"], "silicon 20^3 sub-grid")


def test_silicon_c1_full_grid_vs_oracle(tbk):
    """Config C1 at full size: 20x20x20 mesh, both conventions and eigenvalues, against the oracle."""
    from oracle import workloads as wl

    orc = _oracle()
    p = packed_from(load_golden("silicon.npz"))
    k = wl.kgrid(20, 3)
    _check(
        tbk,
        p,
        k,
        orc.hamilton(p.R, p.hop, p.pos, k, 1),
        orc.hamilton(p.R, p.hop, p.pos, k, 2),
        orc.eigenval_array(p.R, p.hop, p.pos, k),
        "C1 full grid",
    )


def test_reference_cli_known_answer(tbk):
    """reference tests/test_cli_eigenvals.py:47-50: atol 1e-10 against silicon_eigenvals.hdf5."""
    d = load_golden("silicon_cli_eigenvals.npz")
    got = tbk.Evaluator(packed_from(d)).eigenval_array(d["k"])
    assert np.abs(got - d["eig"]).max() <= 1e-10


def test_reference_regression_goldens(tbk):
    """The reference's own goldens for test_simple_hamilton / test_simple_eigenval (np.allclose there)."""
    from oracle import workloads as wl

    d = load_golden("ref_regression.npz")
    for ti, (t1, t2) in enumerate(d["t_values"]):
        m = tbk.KModel.from_packed(wl.simple_model(t1, t2))
        for ki, kpt in enumerate(d["kpt"]):
            for conv in (1, 2):
                got = m.hamilton(tuple(kpt), convention=conv)
                assert got.shape == (2, 2)
                assert np.allclose(got, d[f"H{conv}_t{ti}_k{ki}"]) and np.abs(got - d[f"H{conv}_t{ti}_k{ki}"]).max() < 1e-11
            got = m.eigenval(tuple(kpt))
            assert got.shape == (2,)
            assert np.allclose(got, d[f"E_t{ti}_k{ki}"]) and np.abs(got - d[f"E_t{ti}_k{ki}"]).max() < 1e-10


@pytest.mark.parametrize("dim", [2, 3, 4])
def test_simple_models(tbk, dim):
    d = load_golden("simple_models.npz")
    for ti in range(T_COUNT):
        _block(tbk, d, f"d{dim}_t{ti}_")


def test_haldane_c2_golden(tbk):
    _block(tbk, load_golden("haldane.npz"), "")


def test_edge_cases(tbk):
    d = load_golden("edge_cases.npz")
    for tag in ("empty_", "n1_", "d1_", "shift_"):
        _block(tbk, d, tag)
    m = tbk.KModel.from_packed(packed_from(d, "d1_"))
    # scalar and integer k (reference tests/test_convention.py:23-33)
    got = m.hamilton(0.2, convention=1)
    assert got.shape == (2, 2) and np.abs(got - d["d1_scalar_H1"]).max() < 1e-11
    got = m.eigenval(0.2)
    assert got.shape == (2,) and np.abs(got - d["d1_scalar_eig"]).max() < 1e-10
    got = m.hamilton([[1], [2]], convention=1)
    assert np.abs(got - d["d1_int_H1"]).max() < 1e-11
    # empty batch
    ev = tbk.Evaluator(packed_from(d, "d1_"))
    assert ev.hamilton(np.zeros((0, 1))).shape == (0, 2, 2)
    assert ev.eigenval(np.zeros((0, 1))) == []


@pytest.mark.parametrize("tag", ["c3", "c5s", "n3", "n5", "n7", "n12", "n17", "n33", "n50", "n70"])
def test_synthetic_golden(tbk, tag):
    from oracle import workloads as wl

    d = load_golden("synthetic.npz")
    n_orb, n_half = (int(x) for x in d[f"{tag}_shape"])
    p = wl.synthetic(n_orb, n_half, seed=1234)
    _check(tbk, p, d[f"{tag}_k"], d[f"{tag}_H1"], d[f"{tag}_H2"], d[f"{tag}_eig"], tag)


@pytest.mark.parametrize("tag,size", [("s222", (2, 2, 2)), ("s444", (4, 4, 4))])
def test_supercell_c4(tbk, tag, size):
    """Config C4 (N = 512) and its N = 64 little brother, against reference eigenvalues."""
    from oracle import workloads as wl

    d = load_golden("supercell.npz")
    p = wl.supercell(packed_from(load_golden("silicon.npz")), size)
    assert np.array_equal(p.R, d[f"{tag}_R"])
    H1 = d.get(f"{tag}_H1")
    _check(tbk, p, d[f"{tag}_k"], H1, None, d[f"{tag}_eig"], tag)


# ----------------------------------------------------------------------------------- general path on small models
def test_general_path_on_small_models(tbk, monkeypatch):
    """TBK_FORCE_GEMM routes N <= 8 models through the DMMA GEMM + cooperative eigensolver as well."""
    monkeypatch.setenv("TBK_FORCE_GEMM", "1")
    d = load_golden("silicon.npz")
    p = packed_from(d)
    ev = tbk.Evaluator(p)
    assert ev.path == "gemm+tridiag-ql"
    ev.close()
    _block(tbk, d, "")
    _block(tbk, load_golden("haldane.npz"), "")
    e = load_golden("edge_cases.npz")
    for tag in ("empty_", "n1_", "d1_"):
        _block(tbk, e, tag)
    s = load_golden("simple_models.npz")
    for dim in (2, 3, 4):
        _block(tbk, s, f"d{dim}_t0_")


# ----------------------------------------------------------------------------------- API contract
def test_batch_invariance_bit_exact(tbk):
    """reference tests/test_hamilton.py:21-32, test_eigenval.py:17-20 (assert_allclose rtol 1e-7, atol 0):
    a k-point's result must not depend on where it sits in the batch."""
    from oracle import workloads as wl

    rng = np.random.default_rng(3)
    for p in (packed_from(load_golden("silicon.npz")), wl.synthetic(36, 40), wl.synthetic(12, 10), wl.haldane()):
        m = tbk.KModel.from_packed(p)
        k = rng.uniform(-1, 1, size=(301, p.dim))
        for conv in (1, 2):
            batched = m.hamilton(k, convention=conv)
            for i in (0, 1, 127, 128, 129, 300):
                assert np.array_equal(batched[i], m.hamilton(k[i], convention=conv))
            assert np.array_equal(batched[130:], m.hamilton(k[130:], convention=conv))
        eb = m.eigenval(k)
        assert isinstance(eb, list) and len(eb) == 301
        for i in (0, 1, 127, 128, 129, 300):
            assert np.array_equal(eb[i], m.eigenval(k[i]))
        assert np.array_equal(np.array(eb[130:]), np.array(m.eigenval(k[130:])))


def test_invalid_convention_raises_before_device(tbk):
    m = tbk.KModel.from_packed(packed_from(load_golden("haldane.npz")))
    for bad in ("a", "1", None, 0, 3):
        with pytest.raises(ValueError):
            m.hamilton((0, 0), convention=bad)
    with pytest.raises(ValueError):
        m.hamilton((0, 0, 0))  # wrong number of components


def test_device_pointer_api_matches_host_api(tbk):
    import torch

    from oracle import workloads as wl

    rng = np.random.default_rng(11)
    for p in (wl.haldane(), packed_from(load_golden("silicon.npz")), wl.synthetic(36, 30)):
        ev = tbk.Evaluator(p)
        k = rng.random((1000, p.dim))
        kd = torch.from_numpy(k).cuda()
        e_dev = ev.eigenval_device(kd)
        h_dev = ev.hamilton_device(kd, convention=1)
        ev.check()
        assert np.array_equal(e_dev.cpu().numpy(), ev.eigenval_array(k))
        assert np.array_equal(h_dev.cpu().numpy(), ev.hamilton(k, convention=1))
        assert ev.launch_count > 0


def test_mutation_invalidates_device_copy(tbk):
    """Model.hop is public and mutated in place by the reference (add_hop :1215); the cache must notice."""
    orc = _oracle()
    p = packed_from(load_golden("haldane.npz"))
    m = tbk.KModel.from_packed(p)
    k = np.array([[0.1, 0.7], [0.3, -0.2]])
    before = np.array(m.eigenval(k))
    m.hop[(1, 0)][0, 0] += 0.25
    after = np.array(m.eigenval(k))
    assert np.abs(after - before).max() > 1e-3
    p2 = tbk.pack_model(m)
    assert_eig_close(after, np.array(orc.eigenval(p2.R, p2.hop, p2.pos, k)), "mutated")


class Model:  # stand-in with the attributes the hot path reads (the reference class is not on the GPU box)
    def __init__(self, packed):
        import tbmodels_b200

        self.hop = tbmodels_b200.hop_dict(packed)
        self.pos = packed.pos.copy()
        self.size = packed.size
        self.dim = packed.dim

    def hamilton(self, k, convention=2):
        raise AssertionError("numpy path must not run")

    def eigenval(self, k):
        raise AssertionError("numpy path must not run")


def test_patch_model_class(tbk):
    """install() swaps hamilton/eigenval on a Model class; the instance stays picklable."""
    import pickle

    orc = _oracle()
    p = packed_from(load_golden("silicon.npz"))
    tbk.install(Model)
    try:
        m = Model(p)
        k = np.array([[0.1, 0.2, 0.7], [0.0, 0.0, 0.0]])
        assert_eig_close(np.array(m.eigenval(k)), np.array(orc.eigenval(p.R, p.hop, p.pos, k)), "patched")
        assert_h_close(m.hamilton(k[0], convention=1), orc.hamilton(p.R, p.hop, p.pos, k[0], 1), p, "patched")
        m2 = pickle.loads(pickle.dumps(m))
        assert np.array_equal(np.array(m2.eigenval(k)), np.array(m.eigenval(k)))
    finally:
        tbk.uninstall(Model)
    with pytest.raises(AssertionError):
        Model(p).eigenval([0, 0, 0])


# ----------------------------------------------------------------------------------- full-size properties
def test_haldane_c2_large_properties(tbk):
    """C2 at 2e7 k-points (device resident): size-independent properties + oracle on a subsample."""
    import torch

    from oracle import workloads as wl

    orc = _oracle()
    p = wl.haldane()
    ev = tbk.Evaluator(p)
    n = 20_000_000
    g = torch.Generator(device="cuda").manual_seed(0)
    k = torch.rand((n, 2), dtype=torch.float64, device="cuda", generator=g)
    e = ev.eigenval_device(k)
    ev.check()
    assert bool((e[:, 0] <= e[:, 1]).all())
    # trace of H(k) = sum of eigenvalues: on-site (M, -M) cancels, t2 terms give 2 t2 cos(phi) sum cos = 0 at phi = pi/2
    assert float((e.sum(dim=1)).abs().max()) < 1e-12
    # lattice periodicity: eigenvalues at k and k + integer vector agree
    shift = torch.tensor([3.0, -5.0], dtype=torch.float64, device="cuda")
    e2 = ev.eigenval_device(k[:1_000_000] + shift)
    assert float((e2 - e[:1_000_000]).abs().max()) < 1e-10 * 3.5
    idx = torch.randint(0, n, (20000,), device="cuda", generator=g)
    ks = k[idx].cpu().numpy()
    assert_eig_close(e[idx].cpu().numpy(), orc.eigenval_array(p.R, p.hop, p.pos, ks), "C2 subsample")


def test_c3_properties_and_subsample(tbk):
    """C3 model (N = 36, 251 stored R) on a 64^3 slab of the 256^3 mesh: trace identity, ordering, oracle subsample."""
    import torch

    from oracle import workloads as wl

    orc = _oracle()
    p = wl.synthetic(36, 250, seed=1234)
    ev = tbk.Evaluator(p)
    axes = [torch.arange(64, dtype=torch.float64, device="cuda") / 256.0] * 3
    k = torch.stack(torch.meshgrid(*axes, indexing="ij"), dim=-1).reshape(-1, 3).contiguous()
    e = ev.eigenval_device(k)
    ev.check()
    assert bool((e[:, 1:] >= e[:, :-1]).all())
    h = ev.hamilton_device(k[:4096], convention=2)
    tr = torch.einsum("kii->k", h).real
    assert float((tr - e[:4096].sum(dim=1)).abs().max()) < 1e-10 * float(e.abs().max()) * 36
    idx = torch.arange(0, k.shape[0], 64**3 // 200, device="cuda")
    assert_eig_close(
        e[idx].cpu().numpy(), orc.eigenval_array(p.R, p.hop, p.pos, k[idx].cpu().numpy()), "C3 subsample"
    )


def test_c5_small_sweep(tbk):
    """C5 shape (N = 128) with 200 stored R, 512 k-points against the oracle."""
    from oracle import workloads as wl

    orc = _oracle()
    p = wl.synthetic(128, 200, seed=1234)
    k = np.random.default_rng(2).random((512, 3))
    got = tbk.Evaluator(p).eigenval_array(k)
    assert_eig_close(got[:64], orc.eigenval_array(p.R, p.hop, p.pos, k[:64]), "C5")
    assert np.all(np.diff(got, axis=1) >= 0)


def test_c5_benchmarked_model_full_size(tbk):
    """The model bench.py times as C5 -- synthetic N = 128 with 1001 stored R (126 K-stages of the GEMM, a 270 MB tiled
    weight tensor that no longer fits L2) -- at 64 k-points: ``hamilton`` in both conventions and ``eigenval`` against
    the oracle on all 64 points and against the golden written by the unmodified reference (oracle/make_golden_c5.py)."""
    from oracle import workloads as wl

    orc = _oracle()
    d = load_golden("c5_full.npz")
    n_orb, n_half, seed = (int(x) for x in d["shape"])
    p = wl.synthetic(n_orb, n_half, seed=seed)
    assert p.n_R == 1001 and p.size == 128
    k = d["k"]
    ev = tbk.Evaluator(p)
    assert ev.path == "gemm+tridiag-ql"
    eig = ev.eigenval_array(k)
    assert_eig_close(eig, d["eig"], "C5 vs reference golden")
    assert_eig_close(eig, orc.eigenval_array(p.R, p.hop, p.pos, k), "C5 vs oracle")
    for conv, key in ((1, "H1"), (2, "H2")):
        h = ev.hamilton(k, convention=conv)
        assert_h_close(h[:2], d[key], p, f"C5 hamilton conv {conv} vs reference golden")
        assert_h_close(h, orc.hamilton(p.R, p.hop, p.pos, k, conv), p, f"C5 hamilton conv {conv} vs oracle")
    # the mesh entry point on the same model: a small k-grid whose lines are long enough to be factorised
    dims = (2, 2, 48)
    km = wl.kgrid_points(dims)
    assert_eig_close(ev.eigenval_mesh(dims), orc.eigenval_array(p.R, p.hop, p.pos, km), "C5 mesh vs oracle")
    ev.close()


def test_host_pipeline_chunking(tbk, monkeypatch):
    """Small host chunks force many pipeline iterations; pinned and pageable buffers give identical results."""
    from oracle import workloads as wl

    p = wl.haldane()
    k = np.random.default_rng(4).random((300_001, 2))
    ev = tbk.Evaluator(p)
    ref = ev.eigenval_array(k)
    monkeypatch.setenv("TBK_HOST_CHUNK_MB", "1")
    ev2 = tbk.Evaluator(p)
    kp = tbk.pinned_empty(k.shape)
    kp[:] = k
    out = tbk.pinned_empty((k.shape[0], 2))
    got = ev2.eigenval_array(kp, out=out)
    assert got is out and np.array_equal(got, ref)
    h_ref = ev.hamilton(k[:70_000], convention=1)
    assert np.array_equal(ev2.hamilton(k[:70_000], convention=1), h_ref)


@pytest.mark.parametrize("g", ["1", "32", "64"])
def test_tridiag_variants_agree(tbk, monkeypatch, g):
    """Every thread-group size of the packed kernel and the tensor-core variant give reference eigenvalues."""
    from oracle import workloads as wl

    monkeypatch.setenv("TBK_TRIDIAG_G", g)
    d = load_golden("synthetic.npz")
    for tag in ("n33", "c3", "n50"):
        n_orb, n_half = (int(x) for x in d[f"{tag}_shape"])
        p = wl.synthetic(n_orb, n_half, seed=1234)
        _check(tbk, p, d[f"{tag}_k"], None, None, d[f"{tag}_eig"], f"{tag} G={g}")


@pytest.mark.parametrize("n_orb", [9, 10, 11, 20, 21, 32, 33, 36, 37, 40, 41, 48, 49, 64, 65, 96, 97, 112, 113, 119, 120, 121, 128, 129, 144,
                                   127, 136, 137, 159, 160, 161, 164, 165, 200, 223, 224, 225, 257, 288, 289, 300, 336, 337, 513, 600, 601, 641, 700, 808, 809])
def test_size_boundaries_vs_oracle(tbk, n_orb):
    """Every dispatch boundary of the eigensolver (thread-group sizes, smem / global, QL / bisection, one-stage /
    two-stage reduction from N = 128, its thread configurations, its last size 808)."""
    from oracle import workloads as wl

    orc = _oracle()
    p = wl.synthetic(n_orb, 4, seed=n_orb)
    nk = 37 if n_orb <= 128 else (5 if n_orb <= 300 else (2 if n_orb <= 601 else 1))
    k = np.random.default_rng(n_orb).uniform(-1, 1, size=(nk, 3))
    _check(tbk, p, k, None, orc.hamilton(p.R, p.hop, p.pos, k[:3], 2), orc.eigenval_array(p.R, p.hop, p.pos, k), f"N={n_orb}")


def test_trig_product_kernel_matches_general_small_kernel(tbk, monkeypatch):
    """Nearest-cell N <= 2 models run on the trigonometric-product kernel; the generic fused kernel (TBK_NO_BASIS=1)
    and the oracle must give the same H(k) and eigenvalues, for every (N, dim) instantiation and ragged batch sizes."""
    from oracle import workloads as wl

    orc = _oracle()
    rng = np.random.default_rng(7)
    models = [wl.haldane(), wl.simple_model(0.3, -0.2, dim=2), wl.simple_model(-0.4, 0.15, dim=3),
              wl.synthetic(1, 1, dim=1, seed=3), wl.synthetic(1, 3, dim=2, seed=4), wl.synthetic(1, 5, dim=3, seed=5),
              wl.synthetic(2, 1, dim=1, seed=6), wl.synthetic(2, 4, dim=2, seed=8), wl.synthetic(2, 13, dim=3, seed=9)]
    for p in models:
        assert np.abs(p.R).max() <= 1
        for nk in (1, 511, 1024, 2500):
            k = rng.uniform(-2.0, 2.0, size=(nk, p.dim))
            monkeypatch.delenv("TBK_NO_BASIS", raising=False)
            ev = tbk.Evaluator(p)
            e_new, h_new = ev.eigenval_array(k), ev.hamilton(k[:64], convention=2)
            ev.close()
            monkeypatch.setenv("TBK_NO_BASIS", "1")
            ev = tbk.Evaluator(p)
            e_old, h_old = ev.eigenval_array(k), ev.hamilton(k[:64], convention=2)
            ev.close()
            want = orc.eigenval_array(p.R, p.hop, p.pos, k)
            assert_eig_close(e_new, want, f"product kernel N={p.size} D={p.dim} nk={nk}")
            assert_eig_close(e_old, want, f"generic kernel N={p.size} D={p.dim} nk={nk}")
            assert_h_close(h_new, orc.hamilton(p.R, p.hop, p.pos, k[:64], 2), p, "product kernel H")
            assert np.abs(h_new - h_old).max() <= 1e-13 * h_scale(p)


@pytest.mark.parametrize("lpr", ["16", "32"])
@pytest.mark.parametrize("threads", ["128", "256", "512"])
def test_blocked_tridiag_every_shape(tbk, monkeypatch, threads, lpr):
    """The blocked (panel + tensor-core her2k) reduction forced onto small and ragged sizes: partial last panels,
    sizes that are not multiples of the 8 x 8 blocks, every thread-count instantiation."""
    from oracle import workloads as wl

    orc = _oracle()
    monkeypatch.setenv("TBK_TRIDIAG_PANEL_MIN", "2")
    monkeypatch.setenv("TBK_PANEL_T", threads)
    monkeypatch.setenv("TBK_PANEL_LPR", lpr)
    for n_orb in (9, 15, 16, 17, 24, 31, 33, 40, 63, 65, 100, 129, 200, 255, 256):
        p = wl.synthetic(n_orb, 3, seed=1000 + n_orb)
        k = np.random.default_rng(n_orb).uniform(-1, 1, size=(7 if n_orb <= 129 else 3, 3))
        _check(tbk, p, k, None, None, orc.eigenval_array(p.R, p.hop, p.pos, k), f"blocked N={n_orb} T={threads} LPR={lpr}")


@pytest.mark.parametrize("n_orb", [128, 129, 144, 160, 161, 164, 165, 200, 257, 300, 513, 600, 601])
def test_one_stage_kernels_still_agree(tbk, monkeypatch, n_orb):
    """The one-stage path the two-stage reduction replaced by default for N >= 128 (staged shared-memory kernels, the
    blocked kernel and its hand-over to the staged tail) stays reachable through TBK_TRIDIAG_TWOSTAGE=0."""
    from oracle import workloads as wl

    orc = _oracle()
    monkeypatch.setenv("TBK_TRIDIAG_TWOSTAGE", "0")
    p = wl.synthetic(n_orb, 4, seed=n_orb)
    k = np.random.default_rng(n_orb).uniform(-1, 1, size=(5 if n_orb <= 300 else 2, 3))
    _check(tbk, p, k, None, None, orc.eigenval_array(p.R, p.hop, p.pos, k), f"one-stage N={n_orb}")


@pytest.mark.parametrize("chase", ["1", "4"])
def test_twostage_bulge_chasing_variants(tbk, monkeypatch, chase):
    """The first form of the second stage (four matrices per warp, one sweep at a time) and the 16-warp build of the
    pipelined form stay reachable and agree with the oracle."""
    from oracle import workloads as wl

    orc = _oracle()
    monkeypatch.setenv("TBK_BAND_CHASE", chase)
    for n_orb in (128, 131, 200, 230):
        p = wl.synthetic(n_orb, 3, seed=3000 + n_orb)
        k = np.random.default_rng(n_orb).uniform(-1, 1, size=(5, 3))
        _check(tbk, p, k, None, None, orc.eigenval_array(p.R, p.hop, p.pos, k), f"two-stage chase={chase} N={n_orb}")


@pytest.mark.parametrize("threads", ["128", "256", "257", "512"])
def test_twostage_tridiag_every_shape(tbk, monkeypatch, threads):
    """The two-stage reduction (band of half bandwidth 8 on the tensor cores + bulge chasing, eig_band.cu) forced onto
    small and ragged sizes: sizes that are not multiples of 8, last panels with fewer than 8 rows, a single panel,
    every thread configuration of the first stage; odd batch sizes leave half a warp of the second stage idle."""
    from oracle import workloads as wl

    orc = _oracle()
    monkeypatch.setenv("TBK_TRIDIAG_TWOSTAGE", "12")
    monkeypatch.setenv("TBK_BAND_T", threads)
    for n_orb in (12, 13, 15, 16, 17, 18, 19, 23, 24, 25, 26, 31, 33, 40, 47, 63, 65, 100, 129, 200, 255, 256):
        p = wl.synthetic(n_orb, 3, seed=2000 + n_orb)
        k = np.random.default_rng(n_orb).uniform(-1, 1, size=(7 if n_orb <= 129 else 3, 3))
        _check(tbk, p, k, None, None, orc.eigenval_array(p.R, p.hop, p.pos, k), f"two-stage N={n_orb} T={threads}")


def test_twostage_degenerate_and_sparse_matrices(tbk, monkeypatch):
    """Structured spectra through the two-stage reduction: H = 0, diagonal, block diagonal (zero columns in the panel
    factorisation: reflectors with tau = 0), a supercell (block-sparse hopping matrices)."""
    from oracle import workloads as wl

    orc = _oracle()
    monkeypatch.setenv("TBK_TRIDIAG_TWOSTAGE", "12")
    rng = np.random.default_rng(11)
    n = 40
    R = np.zeros((1, 3), dtype=np.int32)
    pos = np.zeros((n, 3))
    k = rng.uniform(-1, 1, size=(3, 3))
    blocks = np.zeros((n, n), dtype=complex)
    for b0 in range(0, n, 10):
        a = rng.standard_normal((10, 10)) + 1j * rng.standard_normal((10, 10))
        blocks[b0:b0 + 10, b0:b0 + 10] = a + a.conj().T
    from tbmodels_b200._pack import PackedModel

    for name, h in (("zero", np.zeros((n, n), dtype=complex)), ("diagonal", np.diag(rng.standard_normal(n)).astype(complex)),
                    ("blocks", blocks)):
        p = PackedModel(R=R, hop=np.ascontiguousarray((0.5 * h)[None]), pos=pos)
        want = orc.eigenval_array(p.R, p.hop, p.pos, k)
        got = tbk.Evaluator(p).eigenval_array(k)
        assert_eig_close(got, want, f"two-stage {name}")
    sup = wl.supercell(wl.synthetic(8, 6, seed=3), (2, 2, 2))
    ks = rng.uniform(-1, 1, size=(3, 3))
    assert_eig_close(tbk.Evaluator(sup).eigenval_array(ks), orc.eigenval_array(sup.R, sup.hop, sup.pos, ks), "two-stage supercell")


def test_twostage_groups_and_chunks_bit_equal(tbk, monkeypatch):
    """The band arrays of several workspace chunks are collected and chased in one launch: neither the chunk size
    (TBK_WORKSPACE_MB) nor the group size (TBK_BAND_GROUP_MB) may change a bit; the k-mesh entry point takes the same route."""
    from oracle import workloads as wl

    orc = _oracle()
    p = wl.synthetic(230, 3, seed=5)
    k = np.random.default_rng(5).uniform(-1, 1, size=(75, 3))
    ref = tbk.Evaluator(p).eigenval_array(k)
    assert_eig_close(ref, orc.eigenval_array(p.R, p.hop, p.pos, k), "two-stage N=230")
    for ws, grp in (("8", "1024"), ("8", "1"), ("8", "2"), ("3", "3")):
        monkeypatch.setenv("TBK_WORKSPACE_MB", ws)
        monkeypatch.setenv("TBK_BAND_GROUP_MB", grp)
        ev = tbk.Evaluator(p)
        assert np.array_equal(ev.eigenval_array(k), ref), f"workspace {ws} MB, group {grp} MB"
        assert np.array_equal(ev.eigenval_array(k[:31]), ref[:31])
        mesh = ev.eigenval_mesh((3, 2, 5))
        ev.close()
        kk = np.stack(np.meshgrid(np.arange(3) / 3, np.arange(2) / 2, np.arange(5) / 5, indexing="ij"), axis=-1).reshape(-1, 3)
        assert_eig_close(np.asarray(mesh).reshape(-1, 230), orc.eigenval_array(p.R, p.hop, p.pos, kk), "two-stage mesh")
    monkeypatch.delenv("TBK_WORKSPACE_MB")
    monkeypatch.delenv("TBK_BAND_GROUP_MB")


@pytest.mark.parametrize("ratio", ["0", "50", "80"])
def test_staged_tridiag_ratios(tbk, monkeypatch, ratio):
    """Staged shared-memory reduction (trailing block relaunched at a smaller size): every stage boundary the default
    and two other ratios produce, against the oracle; "0" is the single-launch kernel."""
    from oracle import workloads as wl

    orc = _oracle()
    monkeypatch.setenv("TBK_TRIDIAG_STAGES", ratio)
    for n_orb in (25, 36, 47, 49, 64, 97, 119):
        p = wl.synthetic(n_orb, 3, seed=500 + n_orb)
        k = np.random.default_rng(n_orb).uniform(-1, 1, size=(11, 3))
        _check(tbk, p, k, None, None, orc.eigenval_array(p.R, p.hop, p.pos, k), f"staged N={n_orb} ratio={ratio}")


@pytest.mark.parametrize("mid", ["24", "0", "21"])
@pytest.mark.parametrize("stop", ["0", "2", "9", "12", "16"])
def test_register_tridiag_staged_tail(tbk, monkeypatch, stop, mid):
    """Register-resident reduction handing its last `stop` rows to the two-matrices-per-warp kernel (eig_tridiag_reg.cu):
    every hand-over size incl. none, sizes around the kernel's register / corner classes, odd batch sizes (the idle
    half-warp shadows the last matrix), a larger model whose shared-memory stages end in the register kernel; the
    staged and the single-launch reductions are the same sequence of reflectors, so they agree to rounding."""
    from oracle import workloads as wl

    orc = _oracle()
    monkeypatch.setenv("TBK_TRIDIAG_REG_STOP", stop)
    monkeypatch.setenv("TBK_TRIDIAG_REG_MID", mid)
    for n_orb in (21, 24, 25, 28, 29, 32, 33, 36, 40, 64):
        p = wl.synthetic(n_orb, 3, seed=600 + n_orb)
        k = np.random.default_rng(n_orb).uniform(-1, 1, size=(13, 3))
        _check(tbk, p, k, None, None, orc.eigenval_array(p.R, p.hop, p.pos, k), f"reg tail N={n_orb} stop={stop}")
    # bits do not depend on the batch (one matrix alone == the same matrix inside a batch), staged or not
    p = wl.synthetic(36, 5, seed=636)
    k = np.random.default_rng(1).random((7, 3))
    ev = tbk.Evaluator(p)
    full = ev.eigenval_array(k)
    for i in range(7):
        assert np.array_equal(ev.eigenval_array(k[i : i + 1])[0], full[i])
    ev.close()


@pytest.mark.parametrize("overlap", ["1", "0"])
def test_background_ql_overlap_is_bit_invariant(tbk, monkeypatch, overlap):
    """Chunk i's tridiagonal QL runs as a few persistent CTAs on a side stream behind chunk i + 1's build (tbk_api.cu
    eig_chunk, eig_ql.cu ql_background_kernel): same bits as the in-order path and as a single-chunk evaluation, for the
    explicit and the mesh entry points, device and host APIs, many small chunks (tiny workspace) and a ragged last one."""
    import torch

    from oracle import workloads as wl

    orc = _oracle()
    p = wl.synthetic(36, 12, seed=77)
    k = np.random.default_rng(7).random((20011, 3))
    want = tbk.Evaluator(p).eigenval_array(k)  # default workspace: one chunk, no overlap possible (and off by default)
    assert_eig_close(want[:300], orc.eigenval_array(p.R, p.hop, p.pos, k[:300]), "overlap reference")
    monkeypatch.setenv("TBK_WORKSPACE_MB", "16")  # ~1500 k-points per chunk -> 14 chunks
    monkeypatch.setenv("TBK_QL_OVERLAP", overlap)
    ev = tbk.Evaluator(p)
    for _ in range(2):  # second call: buffers and events are reused
        assert np.array_equal(ev.eigenval_array(k), want)
    got = ev.eigenval_device(torch.from_numpy(k).cuda())
    ev.check()
    assert np.array_equal(got.cpu().numpy(), want)
    dims = (7, 11, 64)
    mesh_one = tbk.Evaluator(p)  # env captured at create: still 16 MB, chunks of whole lines
    assert np.array_equal(ev.eigenval_mesh(dims), mesh_one.eigenval_mesh(dims))
    monkeypatch.delenv("TBK_WORKSPACE_MB")
    big = tbk.Evaluator(p).eigenval_mesh(dims)
    assert np.array_equal(ev.eigenval_mesh(dims), big)
    # N = 128: the background CTAs are the 16-thread shape; N = 200: bisection -> falls back to stream order
    for n_orb in (128, 200):
        q = wl.synthetic(n_orb, 3, seed=n_orb)
        kk = np.random.default_rng(n_orb).random((700, 3))
        ref = tbk.Evaluator(q).eigenval_array(kk)
        monkeypatch.setenv("TBK_WORKSPACE_MB", "32")
        assert np.array_equal(tbk.Evaluator(q).eigenval_array(kk), ref)
        monkeypatch.delenv("TBK_WORKSPACE_MB")


@pytest.mark.parametrize("n_orb", [121, 165, 300, 620])
def test_unblocked_large_kernels_still_agree(tbk, monkeypatch, n_orb):
    """The shared-memory / row-sweep kernels the blocked one replaced stay reachable (sizes 601..640, tuning hook)."""
    from oracle import workloads as wl

    orc = _oracle()
    monkeypatch.setenv("TBK_TRIDIAG_NOPANEL", "1")
    p = wl.synthetic(n_orb, 3, seed=n_orb)
    k = np.random.default_rng(n_orb).uniform(-1, 1, size=(3, 3))
    _check(tbk, p, k, None, None, orc.eigenval_array(p.R, p.hop, p.pos, k), f"unblocked N={n_orb}")


@pytest.mark.parametrize("tag", ["toy", "si0", "si1", "si2", "si3"])
def test_kdotp_models(tbk, tag):
    """k.p models (reference src/tbmodels/kdotp.py:51-100) through the same GEMM + eigensolver kernels."""
    d = load_golden("kdotp.npz")
    tc = {tuple(int(x) for x in p): c for p, c in zip(d[f"{tag}_powers"], d[f"{tag}_coeff"])}
    m = tbk.KdotpModel(tc)
    k = d[f"{tag}_k"]
    H = m.hamilton(k)
    scale = max(np.abs(d[f"{tag}_H"]).max(), 1.0)
    assert H.shape == d[f"{tag}_H"].shape and np.abs(H - d[f"{tag}_H"]).max() <= 1e-11 * scale
    assert np.array_equal(H, H.conj().transpose(0, 2, 1))
    e = m.eigenval(k)
    assert isinstance(e, list)
    assert_eig_close(np.array(e), d[f"{tag}_eig"], tag)
    # single point: squeezed result (kdotp.py:80-82, :97-100)
    assert m.hamilton(tuple(k[1])).shape == H.shape[1:]
    assert np.array_equal(m.eigenval(tuple(k[1])), e[1])
    with pytest.raises(ValueError):
        tbk.KdotpModel({(0, 0): [[0, 1], [2, 0]]})  # tests/test_kdotp.py:40-46


@pytest.mark.parametrize("shape", [(2, 4, 2), (1, 3, 1), (3, 5, 3), (8, 20, 3), (12, 10, 3), (31, 9, 3), (32, 9, 3), (33, 9, 3),
                                   (36, 60, 3), (64, 8, 3), (82, 5, 3), (83, 5, 3), (130, 4, 2), (260, 3, 3)])
def test_eigh_eigenvectors(tbk, shape):
    """Eigenvalues + eigenvectors (scipy.linalg.eigh where the reference calls eigvalsh): eigenvalues
    within the parity bound of the oracle's, residual ||H v - v w|| and orthonormality at rounding level, every size class
    of the kernel (one warp, several warps, shared-memory limit 82 / 83, global scratch), fused and GEMM H(k) builds."""
    import torch

    from oracle import workloads as wl

    n_orb, n_half, dim = shape
    packed = wl.synthetic(n_orb, n_half, seed=400 + n_orb, dim=dim)
    orc = _oracle()
    n_k = 9 if n_orb <= 100 else 3
    k = np.random.default_rng(n_orb).uniform(-1, 1, size=(n_k, dim))
    ev = tbk.Evaluator(packed)
    w, v = ev.eigh(k)
    assert w.shape == (n_k, n_orb) and v.shape == (n_k, n_orb, n_orb) and v.dtype == np.complex128
    H = orc.hamilton(packed.R, packed.hop, packed.pos, k)
    want_w, _ = orc.eigh(packed.R, packed.hop, packed.pos, k)
    assert_eig_close(w, want_w, f"eigh N={n_orb}")
    assert np.all(np.diff(w, axis=1) >= 0)
    assert_eig_close(w, ev.eigenval_array(k), f"eigh vs eigenval N={n_orb}")
    rho = float(np.abs(want_w).max())
    resid = np.abs(H @ v - v * w[:, None, :]).max()
    assert resid <= 1e-10 * rho, f"N={n_orb}: residual {resid:.3e}"
    gram = np.abs(v.conj().transpose(0, 2, 1) @ v - np.eye(n_orb)).max()
    assert gram <= 1e-12 * max(n_orb, 8), f"N={n_orb}: V^H V - I = {gram:.3e}"
    # non-degenerate spectrum: the spectral projectors are unique -> compare them with the oracle's (phase-free)
    _, want_v = orc.eigh(packed.R, packed.hop, packed.pos, k[:1])
    gaps = np.diff(want_w[0]).min() if n_orb > 1 else 1.0
    if gaps > 1e-6 * rho:
        proj = np.abs(np.einsum("ij,ik->jk", want_v[0].conj(), v[0]))  # |<u_j | v_k>| = delta_jk
        assert np.abs(proj - np.eye(n_orb)).max() <= 1e-8 / min(1.0, gaps / rho)
    # single point: squeezed pair; device-buffer entry point: same bits
    w0, v0 = ev.eigh(tuple(k[2]))
    assert w0.shape == (n_orb,) and v0.shape == (n_orb, n_orb)
    assert np.array_equal(w0, w[2]) and np.array_equal(v0, v[2])
    wd, vd = ev.eigh_device(torch.from_numpy(k).cuda())
    ev.check()
    assert np.array_equal(wd.cpu().numpy(), w) and np.array_equal(vd.cpu().numpy(), v)
    ev.close()


_EIGH_VARIANT_SIZES = (5, 12, 13, 36, 64, 82)


def _eigh_variant_case(n_orb):
    from oracle import workloads as wl

    return wl.synthetic(n_orb, 3, seed=700 + n_orb), np.random.default_rng(n_orb).random((6, 3))


@pytest.mark.parametrize("smem_max", ["82", "0"])
def test_eigh_storage_variants_agree(tbk, tmp_path, smem_max):
    """The eigenvector kernel with its matrices in shared memory (TBK_EIGH_SMEM_MAX=82, the first version) and on global
    scratch for every size (=0): same code, bit-identical to the default placement.  The limit is read once per process,
    so the variant runs in a fresh interpreter."""
    import subprocess
    import sys

    out = str(tmp_path / "variant.npz")
    code = ("import sys, numpy as np; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import tbmodels_b200 as tbk; from test_gpu_parity import _eigh_variant_case, _EIGH_VARIANT_SIZES\n"
            "res = {}\n"
            "for n in _EIGH_VARIANT_SIZES:\n"
            "    p, k = _eigh_variant_case(n); w, v = tbk.Evaluator(p).eigh(k); res['w%%d' %% n] = w; res['v%%d' %% n] = v\n"
            "np.savez(sys.argv[1], **res)\n" % (ROOT_DIR, os.path.join(ROOT_DIR, "tests")))
    env = dict(os.environ, TBK_EIGH_SMEM_MAX=smem_max)
    subprocess.run([sys.executable, "-c", code, out], check=True, env=env, timeout=600)
    with np.load(out) as got:
        for n_orb in _EIGH_VARIANT_SIZES:
            p, k = _eigh_variant_case(n_orb)
            w0, v0 = tbk.Evaluator(p).eigh(k)
            assert np.array_equal(got["w%d" % n_orb], w0) and np.array_equal(got["v%d" % n_orb], v0), f"N={n_orb}"


def test_eigh_degenerate_and_kdotp(tbk):
    """Degenerate spectra (empty model: H = 0; a model with two identical decoupled blocks) still give an orthonormal
    eigenbasis with zero residual; k.p handles take the same path; KModel exposes the method."""
    from oracle import workloads as wl

    orc = _oracle()
    empty = tbk.pack_arrays(np.zeros((0, 3), dtype=np.int32), np.zeros((0, 5, 5), dtype=complex), np.zeros((5, 3)))
    w, v = tbk.Evaluator(empty).eigh(np.zeros((2, 3)))
    assert np.array_equal(w, np.zeros((2, 5))) and np.abs(v.conj().transpose(0, 2, 1) @ v - np.eye(5)).max() <= 1e-14
    base = wl.synthetic(6, 5, seed=77)
    hop = np.zeros((base.n_R, 12, 12), dtype=complex)
    hop[:, :6, :6] = base.hop
    hop[:, 6:, 6:] = base.hop
    double = tbk.pack_arrays(base.R, hop, np.vstack([base.pos, base.pos]))
    k = np.random.default_rng(3).random((5, 3))
    m = tbk.KModel.from_packed(double)
    w, v = m.eigh(k)
    H = orc.hamilton(double.R, double.hop, double.pos, k)
    rho = float(np.abs(w).max())
    assert np.abs(w[:, 0::2] - w[:, 1::2]).max() <= 1e-12 * rho  # every level twice
    assert np.abs(H @ v - v * w[:, None, :]).max() <= 1e-10 * rho
    assert np.abs(v.conj().transpose(0, 2, 1) @ v - np.eye(12)).max() <= 1e-12 * 12
    m.evaluator().close()
    d = load_golden("kdotp.npz")
    tc = {tuple(int(x) for x in p): c for p, c in zip(d["si1_powers"], d["si1_coeff"])}
    kp = tbk.KdotpModel(tc)
    w, v = kp.evaluator().eigh(d["si1_k"])
    assert_eig_close(w, d["si1_eig"], "k.p eigh")
    assert np.abs(d["si1_H"] @ v - v * w[:, None, :]).max() <= 1e-10 * max(float(np.abs(w).max()), 1.0)


@pytest.mark.parametrize("case", [("silicon", (2, 2, 2)), ("silicon", (4, 4, 4)), ("silicon", (1, 3, 2)), ("syn5", (2, 1, 3)),
                                  ("syn2d", (3, 2)), ("syn1d", (5,)), ("syn12", (1, 1, 1))])
def test_supercell_device_pack(tbk, case, monkeypatch):
    """Model.supercell packed on the device (reference :1645-1724) against the host-packed supercell
    of the oracle's restatement (itself pinned to the reference in make_golden.py): same lattice vectors in the same
    order, H(k) both conventions and eigenvalues BIT-equal to the handle made from the dense supercell arrays (the
    gathered weights are the same numbers), and within the parity bounds of the oracle; block-sparse GEMM == dense GEMM."""
    from oracle import workloads as wl

    name, size = case
    base = {"silicon": lambda: packed_from(load_golden("silicon.npz")), "syn5": lambda: wl.synthetic(5, 9, seed=21),
            "syn2d": lambda: wl.synthetic(4, 6, seed=22, dim=2), "syn1d": lambda: wl.synthetic(3, 4, seed=23, dim=1),
            "syn12": lambda: wl.synthetic(12, 8, seed=24)}[name]()
    big = wl.supercell(base, size)  # dense host arrays, reference semantics
    orc = _oracle()
    m = tbk.KModel.from_packed(base).supercell(size)
    assert m.size == big.size and np.array_equal(m.pos, big.pos)
    ev = m.evaluator()
    assert ev.path == "gemm+tridiag-ql"
    assert np.array_equal(ev.lattice_vectors, big.R)
    monkeypatch.setenv("TBK_FORCE_GEMM", "1")
    ev_host = tbk.Evaluator(big)  # the same model through tbk_model_create on the dense arrays
    k = np.random.default_rng(7).uniform(-1, 1, size=(5 if big.size > 200 else 40, base.dim))
    for conv in (1, 2):
        got = ev.hamilton(k, convention=conv)
        assert np.array_equal(got, ev_host.hamilton(k, convention=conv)), f"conv {conv}: device pack != host pack"
        assert_h_close(got, orc.hamilton(big.R, big.hop, big.pos, k, conv), big, f"{name}{size} H conv{conv}")
    e = ev.eigenval_array(k)
    assert np.array_equal(e, ev_host.eigenval_array(k))
    assert_eig_close(e, orc.eigenval_array(big.R, big.hop, big.pos, k), f"{name}{size} eig")
    assert isinstance(m.eigenval(k), list) and m.hamilton(tuple(k[0])).shape == (big.size, big.size)
    # block-sparse stage skipping off: identical values (the skipped stages only ever add zeros)
    monkeypatch.setenv("TBK_GEMM_DENSE", "1")
    ev_dense = tbk.Evaluator.from_supercell(base, size)
    assert np.array_equal(ev_dense.eigenval_array(k), e)
    monkeypatch.delenv("TBK_GEMM_DENSE")
    for x in (ev, ev_host, ev_dense):
        x.close()
    with pytest.raises(ValueError):
        tbk.KModel.from_packed(base).supercell((2,) * (base.dim + 1))  # reference :1656-1662


def test_supercell_band_folding(tbk):
    """Reference tests/test_supercell.py:24-82: every eigenvalue of the base model at k + (shift / size) appears in the
    supercell spectrum at size * k (atol 1e-7 there; the parity bound here)."""
    from oracle import workloads as wl

    base = wl.synthetic(4, 10, seed=31)
    size = (2, 3, 1)
    sup = tbk.KModel.from_packed(base).supercell(size)
    kb = np.random.default_rng(1).random((6, 3))
    e_sup = np.array(sup.eigenval(kb * np.array(size)))
    ev = tbk.Evaluator(base)
    shifts = [np.array(o) / np.array(size) for o in np.ndindex(*size)]
    folded = np.sort(np.concatenate([ev.eigenval_array(kb + s) for s in shifts], axis=1), axis=1)
    assert_eig_close(e_sup, folded, "band folding")


def _kdotp_term_scales(d, name, powers):
    """Upper bound of |C_p| per Taylor term: (2 pi)^|p| / p! * sum_r |R_r^p| * 2 max|T_r| -- the parity bound scales with it."""
    from math import factorial, pi

    R, hop = d[f"{name}_R"].astype(float), d[f"{name}_hop"]
    tmax = np.abs(hop).reshape(hop.shape[0], -1).max(axis=1)
    out = []
    for p in powers:
        f = (2 * pi) ** int(p.sum()) / np.prod([factorial(int(x)) for x in p])
        out.append(f * float((np.abs(np.prod(R ** p, axis=1)) * 2 * tmax).sum()))
    return np.array(out)


@pytest.mark.parametrize("name", ["haldane", "silicon", "simple3d", "syn12", "syn2d", "syn36", "syn5"])
def test_construct_kdotp_matches_reference(tbk, name):
    """Model.construct_kdotp on the device (reference src/tbmodels/_tb_model.py:942-982) against the
    committed output of the unmodified reference: power tuples in the reference's order, coefficients within
    1e-11 of each term's scale, exactly Hermitian; fused small-N handles and GEMM-path handles both covered."""
    d = load_golden("construct_kdotp.npz")
    p = packed_from(d, name + "_")
    order = int(d[f"{name}_order"])
    m = tbk.KModel.from_packed(p)
    ev = m.evaluator()
    for i, k in enumerate(d[f"{name}_k"]):
        tc = ev.construct_kdotp(k, order)
        powers = np.array(list(tc), dtype=np.int32)
        assert np.array_equal(powers, d[f"{name}_powers"])
        got = np.stack(list(tc.values()))
        want = d[f"{name}_coeff{i}"]
        scales = _kdotp_term_scales(d, name, powers.astype(float))
        err = np.abs(got - want).reshape(len(powers), -1).max(axis=1)
        assert np.all(err <= 1e-11 * np.maximum(scales, 1e-300)), (name, i, float((err / np.maximum(scales, 1e-300)).max()))
        assert np.array_equal(got, got.conj().transpose(0, 2, 1)), "coefficients must be exactly Hermitian"
        kp = m.construct_kdotp(tuple(k), order)  # -> KdotpModel evaluated by the same kernels
        assert isinstance(kp, tbk.KdotpModel)
        assert_eig_close(np.array(kp.eigenval(d[f"{name}_dk{i}"])), d[f"{name}_eig{i}"], f"{name} k.p eigenvalues")
    # batched over expansion points (extension): same bits as point by point; device-buffer entry point too
    import torch

    powers_b, batch = ev.kdotp_coefficients(d[f"{name}_k"], order)
    for i, k in enumerate(d[f"{name}_k"]):
        assert np.array_equal(batch[i], np.stack(list(ev.construct_kdotp(k, order).values())))
    _, dev = ev.kdotp_coefficients_device(torch.from_numpy(d[f"{name}_k"]).cuda(), order)
    assert np.array_equal(dev.cpu().numpy(), batch)
    with pytest.raises(ValueError):
        ev.construct_kdotp(d[f"{name}_k"][0], -1)  # reference :961-962
    with pytest.raises(ValueError):
        ev.construct_kdotp(d[f"{name}_k"], order)  # a list of points is not a single expansion point
    m.evaluator().close()


def test_construct_kdotp_not_defined_for_kdotp_handles(tbk):
    ev = tbk.Evaluator.from_kdotp(np.array([[0, 0]], dtype=np.int32), np.eye(2, dtype=complex)[None])
    with pytest.raises(tbk.TbkError):
        ev.construct_kdotp((0.0, 0.0), 1)
    ev.close()


@pytest.mark.parametrize("tag", ["hr_only_w90", "hr_only_w90v2", "hr_only_si", "hr_wsvec_si", "hr_wsvec_bi", "all_si", "all_bi", "all_bi_nearest"])
def test_reference_wannier_goldens(tbk, tag):
    """The reference's own goldens for Wannier90-derived models (tests/test_wannier.py): N = 7, 8, 10."""
    d = load_golden("ref_wannier.npz")
    p = packed_from(d, tag + "_")
    m = tbk.KModel.from_packed(p)
    got = np.array([m.hamilton(tuple(k)) for k in d["kpt"]])  # per-k calls, exactly like the reference test
    want = d[f"{tag}_H2"]
    assert np.allclose(got, want)
    assert_h_close(got, want, p, tag)
    assert np.array_equal(got, m.hamilton(d["kpt"]))


def test_reference_simple_model_goldens(tbk):
    from oracle import workloads as wl

    d = load_golden("ref_simple_model.npz")
    r = load_golden("ref_regression.npz")
    for ti, (t1, t2) in enumerate(r["t_values"]):
        m = tbk.KModel.from_packed(wl.simple_model(t1, t2))
        for ki, kpt in enumerate(r["kpt"]):
            assert np.abs(m.hamilton(tuple(kpt)) - d[f"hamilton_t{ti}_k{ki}"]).max() < 1e-11
            assert np.abs(m.eigenval(tuple(kpt)) - d[f"eigenval_t{ti}_k{ki}"]).max() < 1e-10


def test_oversize_matrix_fallback(tbk):
    """N = 650 is past the row-sweep kernel's limit (640): exercises the thread-per-row fallback and bisection."""
    from oracle import workloads as wl

    orc = _oracle()
    p = wl.synthetic(650, 2, seed=3)
    k = np.array([[0.11, 0.52, -0.3], [0.0, 0.0, 0.0]])
    got = tbk.Evaluator(p).eigenval_array(k)
    assert_eig_close(got, orc.eigenval_array(p.R, p.hop, p.pos, k), "N=650")


def test_hamilton_above_the_48k_shared_memory_default(tbk):
    """N = 3100: the convention-1 expansion keeps N phase factors (49.6 KB) in dynamic shared memory, past the 48 KB a
    kernel gets without opting in (round-1 advisor finding); both conventions against the oracle."""
    from oracle import workloads as wl

    orc = _oracle()
    p = wl.synthetic(3100, 1, seed=31)
    k = np.array([[0.13, -0.41, 0.77]])
    ev = tbk.Evaluator(p)
    for conv in (1, 2):
        assert_h_close(ev.hamilton(k, convention=conv), orc.hamilton(p.R, p.hop, p.pos, k, conv), p, f"N=3100 conv{conv}")
    ev.close()


@pytest.mark.parametrize("verbosity", [[], ["-v"]])
@pytest.mark.parametrize("kpoints_file_name", ["kpoints.hdf5", "silicon_eigenvals.hdf5"])
def test_cli_eigenvals(tbk, tmp_path, capsys, kpoints_file_name, verbosity):
    """The reference's tests/test_cli_eigenvals.py, same fixtures and the same 1e-10 bound, through
    `python -m tbmodels_b200 eigenvals` (HDF5 in, GPU eigenvalues, HDF5 out)."""
    from conftest import GOLDEN
    from tbmodels_b200 import io
    from tbmodels_b200.__main__ import main

    samples_dir = os.path.join(GOLDEN, "cli_eigenvals")
    out = str(tmp_path / "eigenvals.hdf5")
    rc = main(["eigenvals", "-o", out, "-k", os.path.join(samples_dir, kpoints_file_name),
               "-i", os.path.join(samples_dir, "silicon_model.hdf5")] + verbosity)
    assert rc == 0
    printed = capsys.readouterr().out
    assert ("Calculating energy eigenvalues ..." in printed) == bool(verbosity)
    k, e = io.load_eigenvals(out)
    k_ref, e_ref = io.load_eigenvals(os.path.join(samples_dir, "silicon_eigenvals.hdf5"))
    assert np.array_equal(k, k_ref)
    np.testing.assert_allclose(e, e_ref, rtol=0, atol=1e-10)


def _mesh_points(dims, shift=None):
    axes = [(np.arange(n) + (0.0 if shift is None else shift[d])) / n for d, n in enumerate(dims)]
    return np.stack(np.meshgrid(*axes, indexing="ij"), axis=-1).reshape(-1, len(dims))


@pytest.mark.parametrize("case", ["c3-like", "ragged", "2d", "n50", "shifted", "fallback-small", "fallback-short-lines", "4d"])
def test_eigenval_mesh_matches_explicit_kpoints(tbk, case):
    """tbk_eigenval_mesh: factorised over the last mesh dimension where the model allows it, explicit
    device-generated k-points otherwise; either way the oracle's eigenvalues on the explicit mesh points."""
    from oracle import workloads as wl

    orc = _oracle()
    shift = None
    if case == "c3-like":
        p, dims, fact = wl.synthetic(36, 60, seed=11), (3, 5, 40), True
    elif case == "ragged":
        p, dims, fact = wl.synthetic(13, 30, seed=12), (2, 7, 67), True       # odd N: scalar stores, partial tiles
    elif case == "2d":
        p, dims, fact = wl.synthetic(20, 12, dim=2, seed=13), (9, 33), True
    elif case == "n50":
        p, dims, fact = wl.synthetic(50, 8, seed=14), (2, 2, 130), True       # two 64-point steps + remainder
    elif case == "shifted":
        p, dims, fact, shift = wl.synthetic(24, 20, seed=15), (4, 3, 32), True, (0.5, 0.25, 0.125)
    elif case == "fallback-small":
        p, dims, fact = wl.synthetic(6, 20, seed=16), (3, 4, 50), False       # fused small-N path: no factorisation
    elif case == "fallback-short-lines":
        p, dims, fact = wl.synthetic(24, 60, seed=17), (6, 6, 4), False       # 2 C > points per line
    else:
        p, dims, fact = wl.synthetic(12, 10, dim=4, seed=18), (2, 3, 2, 24), True
    k = _mesh_points(dims, shift)
    want = orc.eigenval_array(p.R, p.hop, p.pos, k)
    ev = tbk.Evaluator(p)
    try:
        assert ev.mesh_factorised(dims) == fact, case
        got = ev.eigenval_mesh(dims, shift)
        assert_eig_close(got, want, f"mesh {case}")
        assert np.all(np.diff(got, axis=1) >= 0)
        # a range of lines (what a rank of a sharded run evaluates) is the matching slice, bit for bit
        n_lines = int(np.prod(dims[:-1]))
        lo, cnt = n_lines // 3, max(1, n_lines // 2)
        cnt = min(cnt, n_lines - lo)
        part = ev.eigenval_mesh_device(dims, shift, first_line=lo, n_lines=cnt).cpu().numpy()
        assert np.array_equal(part, got[lo * dims[-1]:(lo + cnt) * dims[-1]])
        # and the ordinary entry point on the explicit mesh points agrees to rounding
        explicit = ev.eigenval_array(k)
        assert np.abs(explicit - got).max() <= 1e-12 * max(np.abs(want).max(), 1.0)
    finally:
        ev.close()


def test_eigenval_mesh_small_workspace_chunks_lines(tbk, monkeypatch):
    """Several workspace chunks of whole lines give the bits of one chunk (the line is the unit of the factorisation)."""
    from oracle import workloads as wl

    p = wl.synthetic(36, 40, seed=21)
    dims = (7, 9, 48)
    ev = tbk.Evaluator(p)
    ref = ev.eigenval_mesh(dims)
    ev.close()
    monkeypatch.setenv("TBK_WORKSPACE_MB", "8")  # ~ 600 matrices of 36 x 36 per chunk -> 12 lines per chunk
    ev2 = tbk.Evaluator(p)
    try:
        assert ev2.mesh_factorised(dims)
        assert np.array_equal(ev2.eigenval_mesh(dims), ref)
    finally:
        ev2.close()


def test_eigenval_mesh_host_pipeline(tbk, monkeypatch):
    """tbk_eigenval_mesh_host: groups of lines through two device buffers with overlapped D2H -- same bits as the
    device entry point for many small groups (1 MB host chunk), line ranges, a pinned result buffer, both mesh paths."""
    import torch

    from oracle import workloads as wl

    p = wl.synthetic(12, 10, seed=88)
    dims = (9, 10, 48)
    ev = tbk.Evaluator(p)
    want = ev.eigenval_mesh_device(dims).cpu().numpy()
    monkeypatch.setenv("TBK_HOST_CHUNK_MB", "1")
    ev_small = tbk.Evaluator(p)
    assert np.array_equal(ev_small.eigenval_mesh(dims), want)
    out = tbk.pinned_empty((7 * 48, 12))
    got = ev_small.eigenval_mesh(dims, first_line=5, n_lines=7, out=out)
    assert got is out and np.array_equal(out, want[5 * 48 : 12 * 48])
    monkeypatch.setenv("TBK_NO_MESH_FACTOR", "1")  # device-generated k-points through the ordinary path
    ev_plain = tbk.Evaluator(p)
    orc = _oracle()
    k = wl.kgrid_points(dims)
    assert_eig_close(ev_plain.eigenval_mesh(dims), orc.eigenval_array(p.R, p.hop, p.pos, k), "mesh host, unfactorised")
    with pytest.raises(ValueError):
        ev.eigenval_mesh(dims, out=np.empty((3, 12)))
    for e in (ev, ev_small, ev_plain):
        e.close()


def test_eigenval_mesh_argument_errors(tbk):
    from oracle import workloads as wl

    ev = tbk.Evaluator(wl.synthetic(12, 5, seed=3))
    try:
        with pytest.raises(ValueError):
            ev.eigenval_mesh((4, 4))
        with pytest.raises(ValueError):
            ev.eigenval_mesh((4, 0, 4))
        with pytest.raises(tbk.TbkError):
            ev.eigenval_mesh_device((4, 4, 4), first_line=10, n_lines=10)
    finally:
        ev.close()


@pytest.mark.parametrize("n_orb", [36, 64, 130, 200])
def test_degenerate_matrix_structures(tbk, n_orb):
    """Structures that make reflectors trivial or the spectrum highly degenerate, through every tridiagonalisation
    regime (staged shared-memory kernels, blocked kernel): diagonal H, decoupled blocks, rank-one hopping, zero model."""
    from tbmodels_b200 import pack_arrays

    orc = _oracle()
    rng = np.random.default_rng(n_orb)
    k = rng.uniform(-1, 1, size=(5, 3))
    pos = rng.random((n_orb, 3))
    R = np.array([[0, 0, 0], [1, 0, 0], [0, 1, -1]], dtype=np.int32)

    def model(h0, h1, h2):
        return pack_arrays(R, np.stack([0.5 * h0, h1, h2]), pos)

    z = np.zeros((n_orb, n_orb), dtype=complex)
    diag = np.diag(rng.normal(size=n_orb)).astype(complex)
    blocks = z.copy()
    h = n_orb // 2
    a = rng.normal(size=(h, h)) + 1j * rng.normal(size=(h, h))
    blocks[:h, :h] = a + a.conj().T
    blocks[h:, h:] = np.diag(np.full(n_orb - h, 2.0))
    u = rng.normal(size=n_orb) + 1j * rng.normal(size=n_orb)
    rank1 = np.outer(u, u.conj())
    hop_blocks = z.copy()
    hop_blocks[:h, :h] = rng.normal(size=(h, h)) * 0.1
    cases = {
        "diagonal": model(diag, z, z),
        "diagonal + diagonal hopping": model(diag, np.diag(rng.normal(size=n_orb)).astype(complex), z),
        "decoupled blocks": model(blocks, hop_blocks, z),
        "rank one": model(rank1, 0.3 * rank1, z),
        "identity": model(np.eye(n_orb, dtype=complex), z, z),
    }
    for name, p in cases.items():
        _check(tbk, p, k, None, None, orc.eigenval_array(p.R, p.hop, p.pos, k), f"{name} N={n_orb}")
    empty = pack_arrays(np.zeros((0, 3), dtype=np.int32), np.zeros((0, n_orb, n_orb), dtype=complex), pos)
    got = tbk.Evaluator(empty).eigenval_array(k)
    assert got.shape == (5, n_orb) and not got.any()


def test_eigenval_mesh_edge_shapes(tbk):
    """Mesh entry point corner cases: unit dimensions, a single class of lattice vectors (planar model in 3-D), the empty
    model, line lengths around the 64-point step of the line kernel, large and negative shifts."""
    from tbmodels_b200 import pack_arrays
    from oracle import workloads as wl

    orc = _oracle()
    rng = np.random.default_rng(77)
    p = wl.synthetic(17, 25, seed=31)
    planar = pack_arrays(np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, -1, 0]]),
                         np.stack([0.5 * np.eye(9), *(rng.normal(size=(3, 9, 9)) + 1j * rng.normal(size=(3, 9, 9)))]),
                         rng.random((9, 3)))
    empty = pack_arrays(np.zeros((0, 3), dtype=np.int32), np.zeros((0, 11, 11), dtype=complex), rng.random((11, 3)))
    for model, dims, shift in (
        (p, (1, 1, 40), None),
        (p, (1, 3, 64), None),
        (p, (2, 1, 65), (0.0, 0.0, 0.5)),
        (p, (1, 2, 1000), (-3.25, 17.5, 1e3)),
        (planar, (3, 4, 16), None),
        (empty, (2, 2, 8), None),
    ):
        k = wl.kgrid_points(dims, shift)
        want = orc.eigenval_array(model.R, model.hop, model.pos, k)
        ev = tbk.Evaluator(model)
        try:
            got = ev.eigenval_mesh(dims, shift)
        finally:
            ev.close()
        assert_eig_close(got, want, f"mesh {dims} shift {shift} N={model.size}")
