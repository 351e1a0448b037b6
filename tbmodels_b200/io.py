"""HDF5 files of the ``tbmodels eigenvals`` path, read and written without h5py.

* :func:`load_model`      -- ``tbmodels.Model.from_hdf5_file`` (reference src/tbmodels/_tb_model.py:985-1041): datasets
  ``uc, occ, size, dim, pos, sparse`` and a group ``hop/<i>/`` holding ``R`` plus either ``mat`` (dense) or the csr triple
  ``data, indices, indptr, shape``; the stored hoppings go through the same normalisation as the constructor called
  with ``contains_cc=False`` (``_map_to_uc`` :222-245, ``_map_hop_positive_R`` :281-298, zero blocks dropped :207-210).
* :func:`load_kpoints`    -- ``bands_inspect.io.load`` on a ``kpoints_explicit`` or ``eigenvals_data`` file
  (bands_inspect 0.3.2, pinned in the reference's poetry.lock; call site src/tbmodels/_cli.py:246-248).
* :func:`save_eigenvals`  -- ``bands_inspect.io.save(EigenvalsData)`` (call site _cli.py:259): group layout
  ``type_tag, kpoints_obj/{type_tag, kpoints}, eigenvals`` exactly as in the reference's own fixture
  tests/samples/cli_eigenvals/silicon_eigenvals.hdf5.
* :func:`save_model`      -- ``Model.to_hdf5_file`` (:1043-1060 plus the ``type_tag`` fsc.hdf5_io adds).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

from . import _h5lite
from ._pack import PackedModel, pack_arrays

MODEL_TAG = "tbmodels.model"
KPOINTS_TAG = "bands_inspect.kpoints_explicit"
KPOINTS_TAG_LEGACY = "kpoints_explicit"  # written by early bands_inspect versions (the reference's kpoints.hdf5)
EIGENVALS_TAG = "bands_inspect.eigenvals_data"


def _first_nonzero_positive(R) -> Optional[bool]:
    for x in R:
        if x != 0:
            return x > 0
    return None  # R = 0


def _normalise_hop(hop: dict, pos: np.ndarray, size: int) -> Tuple[dict, np.ndarray]:
    """What ``Model.__init__(hop=..., pos=..., contains_cc=False)`` does to its arguments (:174-218)."""
    offsets = np.floor(pos).astype(int)
    if np.any(offsets != 0):  # _map_to_uc, uncommon case (:233-245)
        new_hop: dict = {}
        for R, mat in hop.items():
            for i0, i1 in zip(*np.nonzero(mat)):
                R_new = tuple(int(x) for x in (np.array(R, dtype=int) + offsets[i1] - offsets[i0]))
                new_hop.setdefault(R_new, np.zeros((size, size), dtype=complex))[i0, i1] += mat[i0, i1]
        hop, pos = new_hop, pos % 1
    out: dict = {}
    for R, mat in hop.items():  # _map_hop_positive_R (:281-298)
        sign = _first_nonzero_positive(R)
        if sign is None:
            key, val = R, 0.5 * mat + 0.5 * mat.conjugate().transpose()
        elif sign:
            key, val = R, mat
        else:
            key, val = tuple(-x for x in R), mat.transpose().conjugate()
        out[key] = out[key] + val if key in out else val
    return {R: m for R, m in out.items() if np.any(m)}, pos


def load_model(path: str, with_meta: bool = False):
    """Packed model (and, with ``with_meta``, a dict of ``uc / occ / sparse``) from a TBmodels HDF5 file."""
    tree = _h5lite.load(path)
    grp = tree.get("tb_model", tree)  # development-version files nest everything under 'tb_model' (:1014-1019)
    if "hop" not in grp or not isinstance(grp["hop"], dict):
        raise _h5lite.H5Error(f"'{path}' does not contain a TBmodels model (no 'hop' group)")
    if grp is tree and "type_tag" in tree and tree["type_tag"] != MODEL_TAG:
        raise _h5lite.H5Error(f"'{path}' holds an object of type '{tree['type_tag']}', not a {MODEL_TAG}")
    sparse = bool(grp.get("sparse", False))
    hop: dict = {}
    size = int(grp["size"]) if "size" in grp else None
    for _, g in sorted(grp["hop"].items(), key=lambda kv: int(kv[0])):
        R = tuple(int(x) for x in np.asarray(g["R"]).ravel())
        if sparse:
            n0, n1 = (int(x) for x in np.asarray(g["shape"]).ravel())
            mat = np.zeros((n0, n1), dtype=complex)
            indptr = np.asarray(g["indptr"]).astype(int)
            indices = np.asarray(g["indices"]).astype(int)
            data = np.asarray(g["data"])
            for row in range(n0):
                sl = slice(indptr[row], indptr[row + 1])
                np.add.at(mat[row], indices[sl], data[sl])  # csr semantics: duplicates add up
        else:
            mat = np.array(g["mat"], dtype=complex)
        hop[R] = hop[R] + mat if R in hop else mat
        if size is None:
            size = mat.shape[0]
    if size is None:
        raise _h5lite.H5Error("cannot determine the number of orbitals")
    if "dim" in grp:
        dim = int(grp["dim"])
    elif "pos" in grp:
        dim = np.asarray(grp["pos"]).shape[1]
    elif hop:
        dim = len(next(iter(hop)))
    else:
        dim = np.asarray(grp["uc"]).shape[1]
    pos = np.array(grp["pos"], dtype=float).reshape(size, dim) if "pos" in grp else np.zeros((size, dim))
    hop, pos = _normalise_hop(hop, pos, size)
    R_arr = np.array(list(hop), dtype=np.int32).reshape(len(hop), dim)
    mats = np.stack(list(hop.values())) if hop else np.zeros((0, size, size), dtype=complex)
    packed = pack_arrays(R_arr, mats, pos)
    if with_meta:
        meta = {"uc": np.array(grp["uc"]) if "uc" in grp else None,
                "occ": int(grp["occ"]) if "occ" in grp else None, "sparse": sparse}
        return packed, meta
    return packed


def save_model(packed: PackedModel, path: str, uc=None, occ=None) -> None:
    tree = {"type_tag": MODEL_TAG, "size": np.int64(packed.size), "dim": np.int64(packed.dim),
            "pos": np.asarray(packed.pos), "sparse": np.bool_(False),
            "hop": {str(i): {"R": np.asarray(packed.R[i], dtype=np.int64), "mat": np.asarray(packed.hop[i])}
                    for i in range(packed.n_R)}}
    if uc is not None:
        tree["uc"] = np.asarray(uc, dtype=float)
    if occ is not None:
        tree["occ"] = np.int64(occ)
    _h5lite.save(tree, path)


def load_kpoints(path: str) -> np.ndarray:
    """Explicit k-point list ``[n_k, dim]`` from a ``kpoints_explicit`` file, or the k-points of an ``eigenvals_data`` file
    (what the CLI does at _cli.py:246-248)."""
    tree = _h5lite.load(path)
    tag = tree.get("type_tag")
    if tag == EIGENVALS_TAG:
        tree = tree["kpoints_obj"]
        tag = tree.get("type_tag")
    if tag not in (KPOINTS_TAG, KPOINTS_TAG_LEGACY) or "kpoints" not in tree:
        raise _h5lite.H5Error(f"'{path}': unsupported k-point object '{tag}' (only explicit k-point lists are handled)")
    return np.array(tree["kpoints"], dtype=float)


def load_eigenvals(path: str) -> Tuple[np.ndarray, np.ndarray]:
    tree = _h5lite.load(path)
    if tree.get("type_tag") != EIGENVALS_TAG:
        raise _h5lite.H5Error(f"'{path}' is not a {EIGENVALS_TAG} file")
    return np.array(tree["kpoints_obj"]["kpoints"], dtype=float), np.array(tree["eigenvals"], dtype=float)


def save_eigenvals(path: str, kpoints, eigenvals) -> None:
    kpoints = np.asarray(kpoints, dtype=float)
    eigenvals = np.asarray(eigenvals, dtype=float)
    if kpoints.ndim != 2 or eigenvals.ndim != 2 or len(kpoints) != len(eigenvals):
        raise ValueError("kpoints must be [n_k, dim] and eigenvals [n_k, n_bands]")
    _h5lite.save({"type_tag": EIGENVALS_TAG, "kpoints_obj": {"type_tag": KPOINTS_TAG, "kpoints": kpoints},
                  "eigenvals": eigenvals}, path)
