"""Tiny run of every kernel family, meant to be wrapped in compute-sanitizer (memcheck / racecheck / synccheck)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tbmodels_b200 as tbk  # noqa: E402
from tbmodels_b200 import workloads as wl  # noqa: E402

rng = np.random.default_rng(0)
cases = [
    ("haldane", wl.haldane(), 300),
    ("n3", wl.synthetic(3, 5), 200),
    ("n8", wl.synthetic(8, 9), 200),
    ("n12", wl.synthetic(12, 9), 70),
    ("n17", wl.synthetic(17, 9), 70),
    ("n36", wl.synthetic(36, 20), 150),
    ("n70", wl.synthetic(70, 5), 9),
    ("n130", wl.synthetic(130, 3), 3),
    ("n170", wl.synthetic(170, 2), 2),
]
for name, p, nk in cases:
    ev = tbk.Evaluator(p, device=0)
    k = rng.random((nk, p.dim))
    e = ev.eigenval_array(k)
    h = ev.hamilton(k[: min(nk, 20)], convention=1)
    print(name, ev.path, float(e.sum()), float(np.abs(h).sum()), flush=True)
    ev.close()
print("sanitize smoke done")
