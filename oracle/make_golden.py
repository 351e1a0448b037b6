"""TEST INFRASTRUCTURE ONLY -- regenerate ``tests/golden/*.npz`` from the unmodified reference.

Run in the build container (the only place ``/root/reference`` exists):

    python -m oracle.make_golden

Every array written here is produced either by the reference's own code (``tbmodels.Model.hamilton`` /
``eigenval`` / ``supercell`` / ``from_wannier_files`` imported through ``oracle/ref_shim.py``) or read out of
the reference's own fixture files (``tests/samples/cli_eigenvals/*.hdf5``,
``tests/regression_data/test_hamilton|test_eigenval/*``).  Nothing from the product package is used to
compute expected values; ``tbmodels_b200.pack_model`` is used only to lay the reference's ``hop`` dict out as
arrays, and ``tests/test_oracle_golden.py`` re-checks that layout against the reference dicts.
"""
from __future__ import annotations

import itertools
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.ref_shim import REFERENCE_ROOT, import_reference  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
SAMPLES = os.path.join(REFERENCE_ROOT, "tests", "samples")
REGRESSION = os.path.join(REFERENCE_ROOT, "tests", "regression_data")

T_VALUES = [(t1, t2) for t1 in [-0.1, 0.2, 0.3] for t2 in [-0.2, 0.5]]  # reference tests/parameters.py:7
KPT = [(0.1, 0.2, 0.7), (-0.3, 0.5, 0.2), (0.0, 0.0, 0.0), (0.1, -0.9, -0.7)]  # reference tests/parameters.py:8


def packed_arrays(model):
    from tbmodels_b200 import pack_model

    p = pack_model(model)
    return dict(R=p.R, hop=p.hop, pos=p.pos)


def ref_outputs(model, k, with_h=True):
    out = {"k": np.asarray(k, dtype=float)}
    if with_h:
        out["H1"] = model.hamilton(k, convention=1)
        out["H2"] = model.hamilton(k, convention=2)
    out["eig"] = np.array(model.eigenval(k))
    return out


def simple_model(tb, t1, t2, dim=3, sparse=False):
    """The reference's fixture (tests/conftest.py:155-189), built with the reference's own API."""
    pos = [[0] * 2, [0.5] * 2]
    for p in pos:
        p.extend([0] * (dim - 2))
    model = tb.Model(pos=pos, occ=1, on_site=(1, -1), size=2, dim=None, sparse=sparse)
    for phase, r_part in zip([1, -1j, 1j, -1], itertools.product([0, -1], [0, -1])):
        R = list(r_part) + [0] * (dim - 2)
        model.add_hop(t1 * phase, 0, 1, R)
    for r_part in itertools.permutations([0, 1]):
        R = list(r_part) + [0] * (dim - 2)
        model.add_hop(t2, 0, 0, R)
        model.add_hop(-t2, 1, 1, R)
    return model


def find_block(raw: bytes, expected: np.ndarray, tol=1e-9):
    """Locate a contiguous little-endian f64 block in an HDF5 file that matches ``expected`` (h5py stores small
    datasets contiguously; complex128 as interleaved (r, i) pairs) and return the FILE's values."""
    flat = np.ascontiguousarray(expected).view(np.float64).ravel()
    n = flat.size
    arr = np.frombuffer(raw[: len(raw) // 8 * 8], dtype="<f8")
    cand = np.where(np.abs(arr[: arr.size - n + 1] - flat[0]) <= tol * max(1.0, abs(flat[0])))[0]
    for i in cand:
        blk = arr[i : i + n]
        if np.allclose(blk, flat, rtol=0, atol=tol):
            return blk.copy()
    return None


def find_scattered(raw: bytes, expected: np.ndarray, tol=1e-9):
    """The regression goldens were written by fsc.hdf5_io as nested ``builtins.list`` / ``builtins.number``
    groups: every number is its own scalar dataset (complex as an adjacent (r, i) pair), so the values are
    scattered through the file.  For each expected element return the FILE's value that matches it."""
    expected = np.asarray(expected)
    arr = np.frombuffer(raw[: len(raw) // 8 * 8], dtype="<f8")
    out = np.empty(expected.shape, dtype=expected.dtype)
    flat_out = out.reshape(-1)
    for n, v in enumerate(expected.reshape(-1)):
        if np.iscomplexobj(expected):
            hit = np.where((np.abs(arr[:-1] - v.real) <= tol) & (np.abs(arr[1:] - v.imag) <= tol))[0]
            if hit.size == 0:
                return None
            flat_out[n] = complex(arr[hit[0]], arr[hit[0] + 1])
        else:
            hit = np.where(np.abs(arr - v) <= tol)[0]
            if hit.size == 0:
                return None
            flat_out[n] = arr[hit[0]]
    return out


def main():
    warnings.simplefilter("ignore")
    tb = import_reference()
    from oracle import workloads as wl

    os.makedirs(GOLD, exist_ok=True)
    rng = np.random.default_rng(20240917)

    # ---- C1: silicon Wannier90 model ------------------------------------------------------------------
    si = tb.Model.from_wannier_files(
        hr_file=os.path.join(SAMPLES, "silicon_hr.dat"),
        wsvec_file=os.path.join(SAMPLES, "silicon_wsvec.dat"),
        xyz_file=os.path.join(SAMPLES, "silicon_centres.xyz"),
        win_file=os.path.join(SAMPLES, "silicon.win"),
    )
    k_si = np.concatenate([np.array(KPT), rng.uniform(-1.0, 1.0, size=(60, 3))])
    grid = wl.kgrid(20, 3)
    sub = grid.reshape(20, 20, 20, 3)[::2, ::2, ::2].reshape(-1, 3)  # 1000 points of the 20^3 mesh
    np.savez_compressed(
        os.path.join(GOLD, "silicon.npz"),
        **packed_arrays(si),
        **ref_outputs(si, k_si),
        k_grid=sub,
        eig_grid=np.array(si.eigenval(sub)),
    )

    # ---- reference known answers: CLI eigenvalues fixture (tests/test_cli_eigenvals.py:47-50, atol 1e-10) ----
    si_cli = tb.Model.from_wannier_files(
        hr_file=os.path.join(SAMPLES, "silicon_hr.dat"), wsvec_file=os.path.join(SAMPLES, "silicon_wsvec.dat")
    )
    k_cli = np.array([[x, x, 0.0] for x in np.linspace(0, 1, 11)])
    raw = open(os.path.join(SAMPLES, "cli_eigenvals", "silicon_eigenvals.hdf5"), "rb").read()
    eig_file = find_block(raw, np.array(si_cli.eigenval(k_cli)), tol=1e-8)
    k_file = find_block(raw, k_cli, tol=1e-12)
    assert eig_file is not None and k_file is not None, "CLI fixture block not found"
    np.savez_compressed(
        os.path.join(GOLD, "silicon_cli_eigenvals.npz"),
        **packed_arrays(si_cli),
        k=k_file.reshape(11, 3),
        eig=eig_file.reshape(11, 8),
    )

    # ---- reference regression goldens for the 2-band fixture (tests/test_hamilton.py:10-18, test_eigenval.py:10-14) ----
    reg = {}
    missing = []
    for ti, (t1, t2) in enumerate(T_VALUES):
        model = simple_model(tb, t1, t2)
        for ki, kpt in enumerate(KPT):
            for conv in (1, 2):
                name = f"test_hamilton/test_simple_hamilton[False-{conv}-t_values{ti}-kpt{ki}]"
                raw = open(os.path.join(REGRESSION, name), "rb").read()
                blk = find_scattered(raw, model.hamilton(kpt, convention=conv))
                if blk is None:
                    missing.append(name)
                else:
                    reg[f"H{conv}_t{ti}_k{ki}"] = blk
            name = f"test_eigenval/test_simple_eigenval[False-t_values{ti}-kpt{ki}]"
            raw = open(os.path.join(REGRESSION, name), "rb").read()
            blk = find_scattered(raw, model.eigenval(kpt))
            if blk is None:
                missing.append(name)
            else:
                reg[f"E_t{ti}_k{ki}"] = blk
    assert not missing, f"regression goldens not located: {missing}"
    np.savez_compressed(os.path.join(GOLD, "ref_regression.npz"), t_values=np.array(T_VALUES), kpt=np.array(KPT), **reg)

    # ---- 2-band fixture in 2, 3, 4 dimensions, dense and sparse storage --------------------------------
    simple = {}
    for dim in (2, 3, 4):
        k = rng.uniform(-1.0, 1.0, size=(16, dim))
        k[0] = 0.0
        if dim == 3:
            k[1:5] = np.array(KPT)
        for ti, (t1, t2) in enumerate(T_VALUES):
            m = simple_model(tb, t1, t2, dim=dim, sparse=(ti % 2 == 1))
            tag = f"d{dim}_t{ti}"
            for key, val in {**packed_arrays(m), **ref_outputs(m, k)}.items():
                simple[f"{tag}_{key}"] = val
    np.savez_compressed(os.path.join(GOLD, "simple_models.npz"), **simple)

    # ---- C2: Haldane ------------------------------------------------------------------------------------
    M, t1, t2, phi = 0.3, 1.0, 0.1, np.pi / 2
    hal = tb.Model(on_site=[M, -M], dim=2, occ=1, pos=[[1 / 3, 1 / 3], [2 / 3, 2 / 3]])
    for R in [(0, 0), (-1, 0), (0, -1)]:
        hal.add_hop(t1, 0, 1, R)
    for R in [(1, 0), (-1, 1), (0, -1)]:
        hal.add_hop(t2 * np.exp(1j * phi), 0, 0, R)
        hal.add_hop(t2 * np.exp(-1j * phi), 1, 1, R)
    k = np.random.default_rng(0).random((4096, 2))
    np.savez_compressed(os.path.join(GOLD, "haldane.npz"), **packed_arrays(hal), **ref_outputs(hal, k))

    # ---- C3 / C5 style synthetic models (generator output must equal the reference-constructed model) -------
    def ref_synthetic(n_orb, n_half, seed):
        q = wl.synthetic(n_orb, n_half, seed=seed)
        full = {}
        for R, mat in zip(q.R, q.hop):
            R = tuple(int(x) for x in R)
            if not any(R):
                full[R] = 2 * mat
            else:
                full[R] = mat
                full[tuple(-x for x in R)] = mat.conj().T
        return tb.Model(hop=full, pos=q.pos, contains_cc=True), q

    syn = {}
    for tag, n_orb, n_half, nk_h, nk_e in (
        ("c3", 36, 250, 4, 48),
        ("c5s", 128, 60, 1, 6),
        ("n3", 3, 7, 32, 32),
        ("n5", 5, 9, 32, 32),
        ("n7", 7, 12, 32, 32),
        ("n12", 12, 30, 16, 32),
        ("n17", 17, 20, 8, 32),
        ("n33", 33, 25, 4, 16),
        ("n50", 50, 12, 2, 12),
        ("n70", 70, 10, 1, 8),
    ):
        m, q = ref_synthetic(n_orb, n_half, 1234)
        pr = packed_arrays(m)
        assert np.array_equal(pr["R"], q.R) and np.array_equal(pr["hop"], q.hop), tag
        k = rng.uniform(-0.5, 1.5, size=(nk_e, 3))
        syn[f"{tag}_shape"] = np.array([n_orb, n_half])
        syn[f"{tag}_k"] = k
        syn[f"{tag}_eig"] = np.array(m.eigenval(k))
        syn[f"{tag}_H1"] = m.hamilton(k[:nk_h], convention=1)
        syn[f"{tag}_H2"] = m.hamilton(k[:nk_h], convention=2)
    np.savez_compressed(os.path.join(GOLD, "synthetic.npz"), **syn)

    # ---- C4: silicon supercells (N = 64 and N = 512) --------------------------------------------------------
    sc = {}
    for tag, size, nk in (("s222", (2, 2, 2), 8), ("s444", (4, 4, 4), 2)):
        m = si.supercell(size)
        k = rng.random((nk, 3))
        sc[f"{tag}_k"] = k
        sc[f"{tag}_eig"] = np.array(m.eigenval(k))
        p = packed_arrays(m)
        sc[f"{tag}_R"] = p["R"]
        sc[f"{tag}_pos"] = p["pos"]
        sc[f"{tag}_hop_absum"] = np.array([np.abs(p["hop"]).sum(), np.abs(p["hop"]).max()])
        if tag == "s222":
            sc[f"{tag}_H1"] = m.hamilton(k[:1], convention=1)
    np.savez_compressed(os.path.join(GOLD, "supercell.npz"), **sc)

    # ---- edge cases the reference tests exercise ------------------------------------------------------------
    edge = {}
    # empty model: hop = {} with size = 5 (tests/test_constructors.py:86-89) -> H == 0
    m = tb.Model(hop={}, size=5, dim=3)
    k = rng.random((3, 3))
    for key, val in {**packed_arrays(m), **ref_outputs(m, k)}.items():
        edge[f"empty_{key}"] = val
    # N = 1 (tests/test_slice.py:13)
    m = simple_model(tb, 0.2, 0.5).slice_orbitals([1])
    k = rng.uniform(-1, 1, size=(9, 3))
    for key, val in {**packed_arrays(m), **ref_outputs(m, k)}.items():
        edge[f"n1_{key}"] = val
    # 1-D, 2-orbital model with scalar / integer k (tests/test_convention.py:12-33)
    m = tb.Model(hop={(0,): [[2, 0], [-1j, 3]], (1,): [[0, 1j], [1.5, 0]]}, pos=((0.1,), (0.6,)), contains_cc=False)
    k = np.array([[0.2], [0.0], [-1.3], [2.0], [0.77]])
    for key, val in {**packed_arrays(m), **ref_outputs(m, k)}.items():
        edge[f"d1_{key}"] = val
    edge["d1_scalar_H1"] = m.hamilton(0.2, convention=1)
    edge["d1_scalar_eig"] = m.eigenval(0.2)
    edge["d1_int_H1"] = m.hamilton([[1], [2]], convention=1)
    # wide k range: integer shifts must not change H
    m = simple_model(tb, 0.3, -0.2)
    k = np.array(KPT) + np.array([[7, -13, 40]])
    for key, val in {**packed_arrays(m), **ref_outputs(m, k)}.items():
        edge[f"shift_{key}"] = val
    np.savez_compressed(os.path.join(GOLD, "edge_cases.npz"), **edge)

    # ---- k.p models (reference src/tbmodels/kdotp.py; tests/test_kdotp.py:19-37 and construct_kdotp :47-57) ----
    from tbmodels.kdotp import KdotpModel

    kd = {}
    kp = KdotpModel({(0, 0): np.eye(2), (1, 0): [[0, 1j], [-1j, 0]], (0, 2): [[0, 1], [1, 0]]})
    k = np.array([(0, 0), (0, 0.5), (1, 0), (0.3, -0.7)])
    kd["toy_powers"] = np.array(list(kp.taylor_coefficients.keys()))
    kd["toy_coeff"] = np.array(list(kp.taylor_coefficients.values()))
    kd["toy_k"] = k
    kd["toy_H"] = kp.hamilton(k)
    kd["toy_eig"] = np.array(kp.eigenval(k))
    for order in (0, 1, 2, 3):
        kp = si.construct_kdotp(KPT[0], order=order)
        k = rng.uniform(-0.05, 0.05, size=(24, 3))
        k[0] = 0.0
        kd[f"si{order}_powers"] = np.array(list(kp.taylor_coefficients.keys()))
        kd[f"si{order}_coeff"] = np.array(list(kp.taylor_coefficients.values()))
        kd[f"si{order}_k"] = k
        kd[f"si{order}_H"] = kp.hamilton(k)
        kd[f"si{order}_eig"] = np.array(kp.eigenval(k))
        if order == 0:  # tests/test_kdotp.py:55-57: the k.p model at k = 0 reproduces the tight-binding bands at kpt
            assert np.allclose(kp.eigenval((0, 0, 0)), si.eigenval(KPT[0]))
    np.savez_compressed(os.path.join(GOLD, "kdotp.npz"), **kd)

    # ---- reference regression goldens of tests/test_wannier.py (:18-46, :182-214) and tests/test_simple_model.py (:12-20) ----
    wg = {}
    cases = [
        ("hr_only_w90", dict(hr_file="wannier90_hr.dat", occ=28), "test_wannier_hr_only[wannier90_hr.dat]"),
        ("hr_only_w90v2", dict(hr_file="wannier90_hr_v2.dat", occ=28), "test_wannier_hr_only[wannier90_hr_v2.dat]"),
        ("hr_only_si", dict(hr_file="silicon_hr.dat", occ=28), "test_wannier_hr_only[silicon_hr.dat]"),
        ("hr_wsvec_si", dict(hr_file="silicon_hr.dat", wsvec_file="silicon_wsvec.dat"),
         "test_wannier_hr_wsvec[silicon_hr.dat-silicon_wsvec.dat]"),
        ("hr_wsvec_bi", dict(hr_file="bi_hr.dat", wsvec_file="bi_wsvec.dat"), "test_wannier_hr_wsvec[bi_hr.dat-bi_wsvec.dat]"),
        ("all_si", dict(hr_file="silicon_hr.dat", wsvec_file="silicon_wsvec.dat", xyz_file="silicon_centres.xyz",
                        win_file="silicon.win", pos_kind="wannier", distance_ratio_threshold=1.0),
         "test_wannier_all[silicon_hr.dat-silicon_wsvec.dat-silicon_centres.xyz-silicon.win-pos0-uc0-reciprocal_lattice0-wannier]"),
        ("all_bi", dict(hr_file="bi_hr.dat", wsvec_file="bi_wsvec.dat", xyz_file="bi_centres.xyz", win_file="bi.win",
                        pos_kind="wannier", distance_ratio_threshold=1.0),
         "test_wannier_all[bi_hr.dat-bi_wsvec.dat-bi_centres.xyz-bi.win-pos1-uc1-reciprocal_lattice1-wannier]"),
        ("all_bi_nearest", dict(hr_file="bi_hr.dat", wsvec_file="bi_wsvec.dat", xyz_file="bi_centres.xyz", win_file="bi.win",
                                pos_kind="nearest_atom", distance_ratio_threshold=1.0),
         "test_wannier_all[bi_hr.dat-bi_wsvec.dat-bi_centres.xyz-bi.win-pos2-uc2-reciprocal_lattice2-nearest_atom]"),
    ]
    for tag, kw, fname in cases:
        kw = {k: (os.path.join(SAMPLES, v) if k.endswith("_file") else v) for k, v in kw.items()}
        m = tb.Model.from_wannier_files(**kw)
        H = np.array([m.hamilton(k) for k in KPT])  # convention 2, as the reference test
        raw = open(os.path.join(REGRESSION, "test_wannier", fname), "rb").read()
        blk = find_scattered(raw, H, tol=1e-8)
        assert blk is not None, f"golden values not found in {fname}"
        for key, val in packed_arrays(m).items():
            wg[f"{tag}_{key}"] = val
        wg[f"{tag}_H2"] = blk
    wg["kpt"] = np.array(KPT)
    np.savez_compressed(os.path.join(GOLD, "ref_wannier.npz"), **wg)

    sm = {}
    for ti, (t1, t2) in enumerate(T_VALUES):
        model = simple_model(tb, t1, t2)
        for ki, kpt in enumerate(KPT):
            for what, val in (("hamilton", model.hamilton(kpt)), ("eigenval", model.eigenval(kpt))):
                raw = open(os.path.join(REGRESSION, "test_simple_model", f"test_simple[False-k{ki}-t1{ti}]{what}"), "rb").read()
                blk = find_scattered(raw, val)
                assert blk is not None, (ti, ki, what)
                sm[f"{what}_t{ti}_k{ki}"] = blk
    np.savez_compressed(os.path.join(GOLD, "ref_simple_model.npz"), **sm)

    total = sum(os.path.getsize(os.path.join(GOLD, f)) for f in os.listdir(GOLD))
    print(f"wrote {len(os.listdir(GOLD))} files, {total / 1e6:.2f} MB, to {GOLD}")


if __name__ == "__main__":
    main()
