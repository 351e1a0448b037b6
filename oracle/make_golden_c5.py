"""TEST INFRASTRUCTURE ONLY -- golden vectors of the BENCHMARKED C5 model (synthetic N = 128, 1001 stored R, seed 1234;
BASELINE.json configs[4]) from the unmodified reference.

    python -m oracle.make_golden_c5        (build container)   ->  tests/golden/c5_full.npz

Kept apart from make_golden.py so that the existing golden files stay byte-identical.  The model is built by the
reference's own constructor (``contains_cc=True`` input through ``_reduce_hop``, src/tbmodels/_tb_model.py:247-279)
and must equal the packed arrays of ``oracle.workloads.synthetic(128, 1000, seed=1234)`` -- the model bench.py times.
Committed: 64 k-points, their eigenvalues (reference ``Model.eigenval``) and ``Model.hamilton`` at two of them for both
conventions (the full 64 would be 32 MB).
"""
from __future__ import annotations

import os
import warnings

import numpy as np

from oracle.make_golden import GOLD, packed_arrays
from oracle.ref_shim import import_reference


def main():
    warnings.simplefilter("ignore")
    tb = import_reference()
    from oracle import workloads as wl

    q = wl.synthetic(128, 1000, seed=1234)
    full = {}
    for R, mat in zip(q.R, q.hop):
        R = tuple(int(x) for x in R)
        if not any(R):
            full[R] = 2 * mat
        else:
            full[R] = mat
            full[tuple(-x for x in R)] = mat.conj().T
    m = tb.Model(hop=full, pos=q.pos, contains_cc=True)
    pr = packed_arrays(m)
    assert np.array_equal(pr["R"], q.R) and np.array_equal(pr["hop"], q.hop) and np.array_equal(pr["pos"], q.pos)
    k = np.random.default_rng(50128).uniform(-0.5, 1.5, size=(64, 3))
    np.savez_compressed(
        os.path.join(GOLD, "c5_full.npz"),
        shape=np.array([128, 1000, 1234]),
        k=k,
        eig=np.array(m.eigenval(k)),
        H1=m.hamilton(k[:2], convention=1),
        H2=m.hamilton(k[:2], convention=2),
    )
    print("wrote c5_full.npz", q.n_R, "stored R")


if __name__ == "__main__":
    main()
