"""TEST / BENCH INPUT GENERATORS -- the workloads of BASELINE.json, generated directly in packed (half-set) form.

Not part of the product: ``tbmodels_b200`` never imports this module.  Model construction is OUT OF SCOPE for the
device path and stays the reference's Python; these generators restate the reference's
construction rules with numpy only so that ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` can build their
synthetic input models without importing ``tbmodels``.  ``oracle/make_golden.py`` asserts array equality of every
generator with the unmodified reference, and ``tests/test_host_side.py`` re-checks that live when the reference is
importable (``/root/reference`` or ``oracle/_ref``).

Reference rules restated:
* ``Model.add_hop`` (src/tbmodels/_tb_model.py:1196-1215): R = 0 stores overlap/2 at (i, j) and its conjugate
  /2 at (j, i); first non-zero component of R positive stores overlap at (i, j) under R; otherwise the
  conjugate at (j, i) under -R.
* ``Model.add_on_site`` (:1217-1234) and the ``on_site`` constructor argument (:211-218): half the on-site
  energy goes on the diagonal of the R = 0 matrix.
* ``_reduce_hop`` (:247-279) for ``contains_cc=True`` input: keeps R with positive first non-zero component as
  is, halves R = 0.
* ``supercell`` (:1645-1724) followed by ``_map_hop_positive_R`` (:281-298).
"""
from __future__ import annotations

import itertools

import numpy as np

from tbmodels_b200._pack import PackedModel, pack_arrays


class HopBuilder:
    """Accumulates hoppings with the semantics of ``Model.add_hop`` / ``add_on_site``."""

    def __init__(self, size: int, dim: int, pos=None):
        self.size = size
        self.dim = dim
        self.pos = np.zeros((size, dim)) if pos is None else np.array(pos, dtype=float).reshape(size, dim)
        self.hop: dict = {}

    def _mat(self, R):
        if R not in self.hop:
            self.hop[R] = np.zeros((self.size, self.size), dtype=complex)
        return self.hop[R]

    def add_hop(self, overlap, orbital_1: int, orbital_2: int, R) -> None:
        R = tuple(int(x) for x in R)
        if len(R) != self.dim:
            raise ValueError(f"Dimension of R ({len(R)}) does not match the model dimension ({self.dim})")
        overlap = complex(overlap)
        nonzero = [x for x in R if x != 0]
        if not nonzero:
            m = self._mat(R)
            m[orbital_1, orbital_2] += overlap / 2.0
            m[orbital_2, orbital_1] += overlap.conjugate() / 2.0
        elif nonzero[0] > 0:
            self._mat(R)[orbital_1, orbital_2] += overlap
        else:
            R = tuple(-x for x in R)
            self._mat(R)[orbital_2, orbital_1] += overlap.conjugate()

    def add_on_site(self, on_site) -> None:
        if len(on_site) != self.size:
            raise ValueError(f"The number of on-site energy terms should be {self.size}, but is {len(on_site)}.")
        zero = tuple([0] * self.dim)
        for orbital, energy in enumerate(on_site):
            self.add_hop(energy / 2.0, orbital, orbital, zero)

    def packed(self) -> PackedModel:
        keys = [R for R, m in self.hop.items() if m.any()]
        R = np.array(keys, dtype=np.int32).reshape(len(keys), self.dim)
        hop = np.stack([self.hop[k] for k in keys]) if keys else np.zeros((0, self.size, self.size), complex)
        return pack_arrays(R, hop, self.pos)


def haldane(m=0.3, t1=1.0, t2=0.1, phi=np.pi / 2) -> PackedModel:
    """C2: 2-band Haldane honeycomb model, D = 2, stored keys (0,0), (1,0), (0,1), (1,-1)  (SURVEY 8 d2)."""
    b = HopBuilder(2, 2, pos=[[1 / 3, 1 / 3], [2 / 3, 2 / 3]])
    b.add_on_site([m, -m])
    for R in [(0, 0), (-1, 0), (0, -1)]:
        b.add_hop(t1, 0, 1, R)
    for R in [(1, 0), (-1, 1), (0, -1)]:
        b.add_hop(t2 * np.exp(1j * phi), 0, 0, R)
        b.add_hop(t2 * np.exp(-1j * phi), 1, 1, R)
    return b.packed()


def simple_model(t1, t2, dim=3, pos=None) -> PackedModel:
    """The 2-band fixture every reference hot-path test uses (reference tests/conftest.py:155-189)."""
    if dim < 2:
        raise ValueError("dimension must be at least 2")
    if pos is None:
        pos = [[0.0] * dim, [0.5, 0.5] + [0.0] * (dim - 2)]
    b = HopBuilder(2, dim, pos=pos)
    b.add_on_site((1, -1))
    for phase, r_part in zip([1, -1j, 1j, -1], itertools.product([0, -1], [0, -1])):
        b.add_hop(t1 * phase, 0, 1, list(r_part) + [0] * (dim - 2))
    for r_part in itertools.permutations([0, 1]):
        R = list(r_part) + [0] * (dim - 2)
        b.add_hop(t2, 0, 0, R)
        b.add_hop(-t2, 1, 1, R)
    return b.packed()


def shortest_half_vectors(n_half: int, dim: int = 3) -> np.ndarray:
    """The ``n_half`` shortest non-zero integer vectors whose first non-zero component is positive,
    sorted by |R|^2 with lexicographic tie-break."""
    reach = 1
    while True:
        rng_ = range(-reach, reach + 1)
        cand = []
        for R in itertools.product(rng_, repeat=dim):
            nz = [x for x in R if x != 0]
            if nz and nz[0] > 0:
                cand.append(R)
        # complete shells only: every vector with |R|^2 <= reach^2 is inside the cube
        cand = [R for R in cand if sum(x * x for x in R) <= reach * reach]
        if len(cand) >= n_half:
            cand.sort(key=lambda R: (sum(x * x for x in R), R))
            return np.array(cand[:n_half], dtype=np.int32)
        reach += 1


def synthetic(n_orb: int, n_half: int, seed: int = 1234, dim: int = 3, decay: float = 0.3) -> PackedModel:
    """C3 / C5: random Hermitian-consistent Wannier-like model.

    Stored form (what ``_reduce_hop`` keeps of the full +-R input): ``hop[0] = H_0 / 2`` with ``H_0``
    Hermitian, ``hop[R] = e^{-decay |R|} (A + iB)`` for the ``n_half`` shortest half-set vectors.
    """
    rng = np.random.default_rng(seed)
    Rs = shortest_half_vectors(n_half, dim)
    a0 = rng.normal(size=(n_orb, n_orb)) + 1j * rng.normal(size=(n_orb, n_orb))
    h0 = 0.5 * (a0 + a0.conj().T)
    mats = [0.5 * h0]
    for R in Rs:
        amp = np.exp(-decay * np.sqrt(float(np.dot(R, R))))
        mats.append(amp * (rng.normal(size=(n_orb, n_orb)) + 1j * rng.normal(size=(n_orb, n_orb))))
    pos = rng.random((n_orb, dim))
    R_all = np.concatenate([np.zeros((1, dim), dtype=np.int32), Rs])
    return pack_arrays(R_all, np.stack(mats), pos)


def supercell(packed: PackedModel, size) -> PackedModel:
    """Packed-form restatement of ``Model.supercell`` (reference :1645-1724) + ``_map_hop_positive_R`` (:281-298)."""
    size_array = np.array(size, dtype=int)
    dim = packed.dim
    if size_array.shape != (dim,):
        raise ValueError(f"The given 'size' has incorrect shape {size_array.shape}, should be {(dim,)}.")
    n = packed.size
    vol = int(np.prod(size_array))
    new_size = n * vol
    uc_offsets = [np.array(o) for o in itertools.product(*[range(s) for s in size_array])]
    reduced = packed.pos / size_array
    new_pos = np.concatenate([reduced + off / size_array for off in uc_offsets])
    mult = np.array([int(np.prod(size_array[i:])) for i in range(1, dim + 1)]) * n
    raw: dict = {}
    for uc1_idx, uc1_pos in enumerate(uc_offsets):
        o1 = uc1_idx * n
        for R, mat in zip(packed.R, packed.hop):
            full = uc1_pos + R
            o2 = int(np.inner(mult, full % size_array))
            new_R = tuple(int(x) for x in np.floor(full / size_array).astype(int))
            if new_R not in raw:
                raw[new_R] = np.zeros((new_size, new_size), dtype=complex)
            raw[new_R][o1 : o1 + n, o2 : o2 + n] += mat
    # contains_cc=False: fold onto the half set
    half: dict = {}
    for R, mat in raw.items():
        nz = [x for x in R if x != 0]
        if not nz:
            key, val = R, 0.5 * mat + 0.5 * mat.conjugate().transpose()
        elif nz[0] > 0:
            key, val = R, mat
        else:
            key, val = tuple(-x for x in R), mat.transpose().conjugate()
        if key in half:
            half[key] = half[key] + val
        else:
            half[key] = val.copy()
    keys = [k for k, m in half.items() if m.any()]
    R_arr = np.array(keys, dtype=np.int32).reshape(len(keys), dim)
    return pack_arrays(R_arr, np.stack([half[k] for k in keys]), new_pos)


def kgrid(n: int, dim: int = 3) -> np.ndarray:
    """Uniform ``n^dim`` mesh in [0, 1)^dim, 'ij' ordering."""
    axes = [np.linspace(0.0, 1.0, n, endpoint=False)] * dim
    return np.stack(np.meshgrid(*axes, indexing="ij"), axis=-1).reshape(-1, dim)


def kgrid_points(dims, shift=None) -> np.ndarray:
    """Explicit points of the regular mesh ``k_d = (i_d + shift_d) / dims[d]`` in 'ij' / C order: what
    ``Evaluator.eigenval_mesh`` evaluates without the array."""
    axes = [(np.arange(int(n)) + (0.0 if shift is None else float(shift[d]))) / int(n) for d, n in enumerate(dims)]
    return np.stack(np.meshgrid(*axes, indexing="ij"), axis=-1).reshape(-1, len(dims))


def load_packed(path) -> PackedModel:
    """Load a packed model saved with ``np.savez(path, R=..., hop=..., pos=...)``."""
    with np.load(path) as f:
        return pack_arrays(f["R"], f["hop"], f["pos"])
