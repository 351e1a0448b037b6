"""Tiny runs of the two-stage tridiagonalisation (eig_band.cu), meant to be wrapped in compute-sanitizer."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tbmodels_b200 as tbk  # noqa: E402
from oracle import workloads as wl  # noqa: E402

rng = np.random.default_rng(0)
os.environ["TBK_TRIDIAG_TWOSTAGE"] = "12"
for threads in ("257", "512"):
    os.environ["TBK_BAND_T"] = threads
    for n_orb, nk in ((13, 3), (26, 5), (40, 2), (70, 3)):
        p = wl.synthetic(n_orb, 3, seed=n_orb)
        ev = tbk.Evaluator(p, device=0)
        k = rng.random((nk, 3))
        e = ev.eigenval_array(k)
        ref = np.linalg.eigvalsh(ev.hamilton(k, convention=2))
        print("two-stage", threads, n_orb, float(np.abs(e - ref).max()), flush=True)
        ev.close()
os.environ.pop("TBK_TRIDIAG_TWOSTAGE")
os.environ.pop("TBK_BAND_T")
os.environ["TBK_WORKSPACE_MB"] = "3"
os.environ["TBK_BAND_GROUP_MB"] = "1"
p = wl.synthetic(230, 2, seed=1)
ev = tbk.Evaluator(p, device=0)
k = rng.random((11, 3))
e = ev.eigenval_array(k)
print("two-stage default N=230", float(np.abs(e - np.linalg.eigvalsh(ev.hamilton(k, convention=2))).max()), flush=True)
ev.close()
print("sanitize two-stage done")
