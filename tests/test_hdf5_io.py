"""HDF5 plumbing of the `tbmodels eigenvals` path -- runs without a GPU.

The fixtures under tests/golden/cli_eigenvals/ are the reference's own data files (written by h5py / bands_inspect /
fsc.hdf5_io): the reader is pinned on them, the writer is pinned by emitting byte-identical datatype / dataspace /
fill-value messages and by the round trip through the reader.
"""
import glob
import os
import pathlib
import struct

import numpy as np
import pytest

from conftest import GOLDEN, load_golden

from tbmodels_b200 import _h5lite, io

CLI = os.path.join(GOLDEN, "cli_eigenvals")


def _oracle():
    from oracle import tb_oracle

    return tb_oracle


def test_reader_kpoints_and_eigenvals_fixtures():
    k = io.load_kpoints(os.path.join(CLI, "kpoints.hdf5"))
    assert k.shape == (11, 3)
    assert np.allclose(k, [[x, x, 0.0] for x in np.linspace(0, 1, 11)], atol=1e-15)
    k2, e = io.load_eigenvals(os.path.join(CLI, "silicon_eigenvals.hdf5"))
    assert np.array_equal(k2, k) and e.shape == (11, 8)
    # the CLI accepts an eigenvals_data file as k-point input (reference _cli.py:247-248)
    assert np.array_equal(io.load_kpoints(os.path.join(CLI, "silicon_eigenvals.hdf5")), k)
    # values quoted in SURVEY.md section 8 c3
    assert abs(e[0, 0] - (-5.821847625730381)) < 1e-12 and abs(e[0, 7] - 9.705551893206355) < 1e-12


def test_model_file_reproduces_the_reference_cli_answer():
    """Model read from silicon_model.hdf5 + oracle eigenvalues == silicon_eigenvals.hdf5 at the reference's own
    tolerance (tests/test_cli_eigenvals.py:47-50, atol 1e-10)."""
    p, meta = io.load_model(os.path.join(CLI, "silicon_model.hdf5"), with_meta=True)
    assert (p.size, p.dim) == (8, 3) and meta["sparse"] is False and meta["uc"].shape == (3, 3)
    # reduced-form invariants of Model.hop (reference :206-218, :281-298)
    for R in p.R:
        nz = [x for x in R if x != 0]
        assert not nz or nz[0] > 0
    i0 = [i for i, R in enumerate(p.R) if not R.any()]
    assert len(i0) == 1 and np.allclose(p.hop[i0[0]], p.hop[i0[0]].conj().T)
    k, want = io.load_eigenvals(os.path.join(CLI, "silicon_eigenvals.hdf5"))
    got = _oracle().eigenval_array(p.R, p.hop, p.pos, k)
    assert np.abs(got - want).max() <= 1e-10


def _tree_equal(a, b):
    if isinstance(a, dict):
        return isinstance(b, dict) and set(a) == set(b) and all(_tree_equal(a[k], b[k]) for k in a)
    if isinstance(a, str):
        return a == b
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and a.dtype.kind == b.dtype.kind and np.array_equal(a, b)


def test_writer_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    tree = {
        "type_tag": "bands_inspect.eigenvals_data",
        "eigenvals": rng.normal(size=(11, 8)),
        "kpoints_obj": {"type_tag": "bands_inspect.kpoints_explicit", "kpoints": rng.random((11, 3))},
        "flag": np.bool_(True),
        "n": np.int64(-7),
        "c": rng.normal(size=(3, 3)) + 1j * rng.normal(size=(3, 3)),
        "many": {str(i): {"R": np.array([i, 0, -i])} for i in range(95)},  # more links than one symbol node holds
        "empty": np.zeros((0, 3)),
        "unicode": "häm",
    }
    path = str(tmp_path / "t.hdf5")
    _h5lite.save(tree, path)
    assert _tree_equal(tree, _h5lite.load(path))


def _messages(path, *names):
    data = pathlib.Path(path).read_bytes()
    r = _h5lite._Reader(data)
    addr = r.root_header
    for n in names:
        addr = r.links(r.messages(addr))[n]
    return {t: data[b:b + s] for t, b, s in r.messages(addr) if t in (0x0001, 0x0003, 0x0005)}


def test_writer_emits_the_messages_h5py_emits(tmp_path):
    """Dataspace, datatype and fill-value messages byte-identical to the h5py-written fixtures for every kind of
    dataset on the path: float64 matrix, int64 scalar and vector, complex128 matrix, bool scalar, vlen string."""
    ref_model = os.path.join(CLI, "silicon_model.hdf5")
    tree = _h5lite.load(ref_model)
    mine = str(tmp_path / "m.hdf5")
    _h5lite.save(tree, mine)
    for names in (("pos",), ("uc",), ("size",), ("sparse",), ("type_tag",), ("hop", "0", "mat"), ("hop", "0", "R")):
        assert _messages(mine, *names) == _messages(ref_model, *names), names
    # superblock: same version, offset / length sizes and group B-tree parameters
    a, b = pathlib.Path(mine).read_bytes()[:24], pathlib.Path(ref_model).read_bytes()[:24]
    assert a == b
    assert _tree_equal(tree, _h5lite.load(mine))


def test_save_model_and_eigenvals_round_trip(tmp_path):
    p = io.load_model(os.path.join(CLI, "silicon_model.hdf5"))
    path = str(tmp_path / "model.hdf5")
    io.save_model(p, path, uc=np.eye(3), occ=4)
    q, meta = io.load_model(path, with_meta=True)
    assert meta["occ"] == 4 and np.array_equal(meta["uc"], np.eye(3))
    assert {tuple(r): h.tobytes() for r, h in zip(q.R, q.hop)} == {tuple(r): h.tobytes() for r, h in zip(p.R, p.hop)}
    assert np.array_equal(q.pos, p.pos)
    k, e = io.load_eigenvals(os.path.join(CLI, "silicon_eigenvals.hdf5"))
    out = str(tmp_path / "e.hdf5")
    io.save_eigenvals(out, k, e)
    k2, e2 = io.load_eigenvals(out)
    assert np.array_equal(k, k2) and np.array_equal(e, e2)
    assert _tree_equal(_h5lite.load(out), _h5lite.load(os.path.join(CLI, "silicon_eigenvals.hdf5")))


def test_loader_normalises_like_the_constructor(tmp_path):
    """contains_cc=False path of Model.__init__: negative-R keys are mapped to +R (conjugate transpose), the R = 0
    block is Hermitised, positions outside the unit cell shift the lattice vectors, zero blocks are dropped."""
    h = np.array([[0.0, 1.0 + 2.0j], [0.5j, 0.0]])
    tree = {"size": np.int64(2), "dim": np.int64(1), "sparse": np.bool_(False), "pos": np.array([[0.25], [1.5]]),
            "hop": {"0": {"R": np.array([-1]), "mat": h}, "1": {"R": np.array([0]), "mat": np.array([[1.0, 2.0], [0.0, 3.0]], dtype=complex)},
                    "2": {"R": np.array([5]), "mat": np.zeros((2, 2), dtype=complex)}}}
    path = str(tmp_path / "n.hdf5")
    _h5lite.save(tree, path)
    p = io.load_model(path)
    assert np.allclose(p.pos, [[0.25], [0.5]])
    hop = {tuple(r): m for r, m in zip(p.R, p.hop)}
    # orbital 1 sits in cell +1: element (i0, i1) of R moves to R + off[i1] - off[i0]
    want = {}
    off = [0, 1]
    for R, mat in (((-1,), h), ((0,), np.array([[1.0, 2.0], [0.0, 3.0]], dtype=complex))):
        for i0 in range(2):
            for i1 in range(2):
                if mat[i0, i1] != 0:
                    want.setdefault((R[0] + off[i1] - off[i0],), np.zeros((2, 2), dtype=complex))[i0, i1] += mat[i0, i1]
    red = {}
    for R, m in want.items():
        if R[0] > 0:
            red[R] = red.get(R, 0) + m
        elif R[0] < 0:
            red[(-R[0],)] = red.get((-R[0],), 0) + m.conj().T
        else:
            red[R] = red.get(R, 0) + 0.5 * m + 0.5 * m.conj().T
    red = {R: m for R, m in red.items() if np.any(m)}
    assert set(hop) == set(red) and all(np.allclose(hop[R], red[R]) for R in red)


def test_unsupported_files_fail_loudly(tmp_path):
    bad = tmp_path / "x.hdf5"
    bad.write_bytes(b"not an hdf5 file at all")
    with pytest.raises(_h5lite.H5Error):
        _h5lite.load(str(bad))
    with pytest.raises(_h5lite.H5Error):
        io.load_model(os.path.join(CLI, "kpoints.hdf5"))
    with pytest.raises(_h5lite.H5Error):
        io.load_kpoints(os.path.join(CLI, "silicon_model.hdf5"))


@pytest.mark.skipif(not os.path.isdir("/root/reference/tests/samples"), reason="reference checkout not present")
def test_reader_on_every_reference_hdf5_sample():
    """Every HDF5 file the reference ships (models with and without type_tag, symmetry groups, legacy objects, the
    regression goldens) parses; models load into packed arrays."""
    files = sorted(glob.glob("/root/reference/tests/samples/*.hdf5"))
    files += sorted(glob.glob("/root/reference/tests/regression_data/test_eigenval/*"))[:12]
    assert len(files) >= 12
    n_models = 0
    for f in files:
        tree = _h5lite.load(f)
        assert isinstance(tree, dict) and tree
        if "hop" in tree and "size" in tree:
            p = io.load_model(f)
            assert p.hop.shape == (p.n_R, p.size, p.size)
            n_models += 1
    assert n_models >= 4


@pytest.mark.skipif(not os.path.isdir("/root/reference/tests/regression_data"), reason="reference checkout not present")
def test_reader_agrees_with_committed_regression_goldens():
    """The eigenvalue goldens that oracle/make_golden.py recovered by a raw byte scan are the same numbers a proper
    parse of the HDF5 files yields."""
    d = load_golden("ref_regression.npz")
    keys = [k for k in d if k.startswith("E_")]
    assert keys
    files = sorted(glob.glob("/root/reference/tests/regression_data/test_eigenval/*"))
    vals = []
    for f in files:
        tree = _h5lite.load(f)
        leaves = []

        def walk(t):
            for v in t.values():
                walk(v) if isinstance(v, dict) else leaves.append(v)

        walk(tree)
        vals += [np.asarray(v, dtype=float).ravel() for v in leaves if not isinstance(v, str) and np.asarray(v).dtype.kind == "f"]
    flat = np.concatenate(vals)
    for k in keys[:16]:
        for x in np.asarray(d[k]).ravel():
            assert np.any(flat == x), k


@pytest.mark.parametrize("n_links", [0, 8, 9, 256, 257, 1001, 8200])
def test_writer_large_groups_use_multi_level_btrees(tmp_path, n_links):
    """Groups beyond one B-tree node (256 links) -- e.g. the hop group of the C5 model, 1001 entries -- get the two- /
    three-level v1 B-trees h5py writes (separator keys = first key of the child, siblings linked)."""
    tree = {"hop": {str(i): {"R": np.array([i, 0, -i])} for i in range(n_links)}, "size": np.int64(3)}
    path = str(tmp_path / "big.hdf5")
    _h5lite.save(tree, path)
    back = _h5lite.load(path)
    assert set(back["hop"]) == set(tree["hop"])
    assert all(np.array_equal(back["hop"][k]["R"], tree["hop"][k]["R"]) for k in tree["hop"])
    data = pathlib.Path(path).read_bytes()
    r = _h5lite._Reader(data)
    hop = r.links(r.messages(r.root_header))["hop"]
    bt = next(struct.unpack_from("<Q", data, b)[0] for t, b, _ in r.messages(hop) if t == 0x0011)
    level, used = struct.unpack_from("<BH", data, bt + 5)
    want_level = 0 if n_links <= 256 else (1 if n_links <= 8192 else 2)
    assert level == want_level and used <= 32
    if level > 0:  # children: key 0 of child i equals the parent's key i, siblings chained left to right
        prev = None
        for i in range(used):
            key, child = struct.unpack_from("<QQ", data, bt + 24 + 16 * i)
            left, right = struct.unpack_from("<QQ", data, child + 8)
            assert struct.unpack_from("<Q", data, child + 24)[0] == key
            assert left == (prev if prev is not None else 0xFFFFFFFFFFFFFFFF)
            prev = child
        assert right == 0xFFFFFFFFFFFFFFFF


def test_loader_reads_csr_hoppings(tmp_path):
    """`sparse=True` files (reference to_hdf5 :1052-1056: data / indices / indptr / shape per R) load to the same packed
    model as the dense form."""
    import scipy.sparse as sp

    rng = np.random.default_rng(3)
    mats = {(0, 0): None, (1, 0): None, (0, 1): None, (1, -1): None}
    dense = {}
    for R in mats:
        m = (rng.random((5, 5)) < 0.3) * (rng.normal(size=(5, 5)) + 1j * rng.normal(size=(5, 5)))
        if R == (0, 0):
            m = 0.25 * (m + m.conj().T)
        dense[R] = m
    pos = rng.random((5, 2))

    def tree(sparse):
        hop = {}
        for i, (R, m) in enumerate(dense.items()):
            g = {"R": np.array(R)}
            if sparse:
                c = sp.csr_matrix(m)
                g.update(data=c.data, indices=c.indices.astype(np.int64), indptr=c.indptr.astype(np.int64),
                         shape=np.array(c.shape))
            else:
                g["mat"] = m
            hop[str(i)] = g
        return {"type_tag": io.MODEL_TAG, "size": np.int64(5), "dim": np.int64(2), "pos": pos,
                "sparse": np.bool_(sparse), "hop": hop, "occ": np.int64(2)}

    pa, pb = str(tmp_path / "dense.hdf5"), str(tmp_path / "csr.hdf5")
    _h5lite.save(tree(False), pa)
    _h5lite.save(tree(True), pb)
    a, (b, meta) = io.load_model(pa), io.load_model(pb, with_meta=True)
    assert meta["sparse"] is True and meta["occ"] == 2
    assert np.array_equal(a.R, b.R) and np.array_equal(a.hop, b.hop) and np.array_equal(a.pos, b.pos)
    assert {tuple(r) for r in a.R} == set(dense)


def test_truncated_and_corrupt_files_raise_h5error(tmp_path):
    data = pathlib.Path(os.path.join(CLI, "kpoints.hdf5")).read_bytes()
    for cut in (100, 700, 2100, len(data) - 3000):
        f = tmp_path / f"cut{cut}.hdf5"
        f.write_bytes(data[:cut])
        with pytest.raises(_h5lite.H5Error):
            _h5lite.load(str(f))
    # a group whose only link points back at the root object header
    r = _h5lite._Reader(data)
    links = r.links(r.messages(r.root_header))
    bad = bytearray(data)
    snod = data.index(b"SNOD")
    struct.pack_into("<Q", bad, snod + 8 + 8, r.root_header)  # first entry's object header address -> root
    f = tmp_path / "cycle.hdf5"
    f.write_bytes(bytes(bad))
    with pytest.raises(_h5lite.H5Error):
        _h5lite.load(str(f))
    assert links  # (the untouched file still parses)
