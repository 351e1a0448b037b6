"""Host-side packing of a TBmodels ``Model`` into the three dense arrays the device consumes.

Source of truth is what the reference hot path reads:

* ``model.hop``  -- ``defaultdict(R-tuple -> N x N complex ndarray or csr wrapper)`` holding the *half set*
  (first non-zero component of R positive) plus HALF of the R = 0 block
  (reference src/tbmodels/_tb_model.py:206-218, 247-279); iterated in dict order at :1111;
* ``model.pos``  -- ``[N, dim]`` reduced orbital positions (:186-191), used by convention 1 (:1124-1128);
* ``model.size``, ``model.dim``.

The half-set semantics are preserved verbatim: the device adds the Hermitian conjugate itself, exactly as
``H += H.conjugate().transpose()`` does at :1123.  Sparse matrices are densified once here instead of once
per R per call (``_array_cast``, :1326-1331).
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass, field

import numpy as np


def _dense(mat) -> np.ndarray:
    """Dense complex128 copy of a hopping matrix (ndarray, scipy sparse, or the reference's csr wrapper)."""
    if hasattr(mat, "toarray"):  # avoids np.array(csr), which warns under numpy 2 (reference _sparse_matrix.py:20-21)
        mat = mat.toarray()
    return np.ascontiguousarray(mat, dtype=np.complex128)


@dataclass(frozen=True)
class PackedModel:
    """Dense, contiguous snapshot of the state ``hamilton`` / ``eigenval`` read."""

    R: np.ndarray  # int32 [n_R, dim]
    hop: np.ndarray  # complex128 [n_R, N, N]   (half-set semantics, R = 0 halved)
    pos: np.ndarray  # float64 [N, dim]
    _digest: list = field(default_factory=list, repr=False, compare=False)

    def __post_init__(self):
        if self.R.dtype != np.int32 or self.hop.dtype != np.complex128 or self.pos.dtype != np.float64:
            raise TypeError("PackedModel arrays must be int32 / complex128 / float64")
        n_R, dim = self.R.shape
        size = self.pos.shape[0]
        if self.pos.shape != (size, dim) or self.hop.shape != (n_R, size, size):
            raise ValueError(
                f"inconsistent packed shapes: R {self.R.shape}, hop {self.hop.shape}, pos {self.pos.shape}"
            )

    @property
    def size(self) -> int:
        return self.pos.shape[0]

    @property
    def dim(self) -> int:
        return self.pos.shape[1]

    @property
    def n_R(self) -> int:
        return self.R.shape[0]

    def digest(self) -> bytes:
        """Content hash used to decide whether a cached device copy is still valid."""
        if not self._digest:
            h = hashlib.blake2b(digest_size=16)
            h.update(np.array([self.size, self.dim, self.n_R], dtype=np.int64).tobytes())
            h.update(self.R.tobytes())
            h.update(self.hop.tobytes())
            h.update(self.pos.tobytes())
            self._digest.append(h.digest())
        return self._digest[0]


def pack_arrays(R, hop, pos) -> PackedModel:
    """Build a :class:`PackedModel` from raw arrays (any integer / complex / float dtypes)."""
    pos = np.ascontiguousarray(pos, dtype=np.float64)
    if pos.ndim != 2:
        raise ValueError("pos must be [N, dim]")
    size, dim = pos.shape
    R = np.asarray(R)
    if R.size == 0:
        R = np.zeros((0, dim), dtype=np.int32)
    if not np.all(np.asarray(R) == np.rint(R)):
        raise ValueError("lattice vectors R must be integers")
    R = np.ascontiguousarray(R, dtype=np.int32).reshape(-1, dim)
    hop = np.ascontiguousarray(hop, dtype=np.complex128).reshape(R.shape[0], size, size)
    return PackedModel(R=R, hop=hop, pos=pos)


def pack_model(model) -> PackedModel:
    """Pack any object exposing ``hop``, ``pos``, ``size`` and ``dim`` like ``tbmodels.Model``.

    All-zero matrices (which appear when a missing key of the ``defaultdict`` is merely read,
    reference :206) are dropped: they contribute exact zeros to the sum at :1117-1122.
    """
    size = int(model.size)
    dim = int(model.dim)
    pos = np.ascontiguousarray(np.asarray(model.pos, dtype=np.float64)).reshape(size, dim)
    keys = []
    mats = []
    for R, mat in model.hop.items():
        dense = _dense(mat)
        if dense.shape != (size, size):
            raise ValueError(f"hopping matrix of shape {dense.shape} found, should be ({size},{size})")
        if len(R) != dim:
            raise ValueError(f"The length of R = {R} does not match the dimensionality of the system ({dim})")
        if not dense.any():
            continue
        keys.append(tuple(int(x) for x in R))
        mats.append(dense)
    if keys:
        R_arr = np.array(keys, dtype=np.int32).reshape(len(keys), dim)
        hop_arr = np.stack(mats).astype(np.complex128, copy=False)
    else:
        R_arr = np.zeros((0, dim), dtype=np.int32)
        hop_arr = np.zeros((0, size, size), dtype=np.complex128)
    return PackedModel(R=np.ascontiguousarray(R_arr), hop=np.ascontiguousarray(hop_arr), pos=pos)


def hop_dict(packed: PackedModel) -> dict:
    """Inverse of :func:`pack_model`: the ``{R-tuple: matrix}`` dict the reference stores."""
    return {tuple(int(x) for x in r): packed.hop[i].copy() for i, r in enumerate(packed.R)}
