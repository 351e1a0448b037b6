"""Tiny run of every kernel family, meant to be wrapped in compute-sanitizer (memcheck / racecheck / synccheck)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tbmodels_b200 as tbk  # noqa: E402
from oracle import workloads as wl  # noqa: E402

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
    ("simple3d", wl.simple_model(0.3, -0.2), 100),      # N = 2, D = 3: trigonometric-product kernel, 27 functions
    ("n1", wl.synthetic(1, 3, dim=2), 100),             # N = 1
    ("blocked n36", wl.synthetic(36, 5), 5),            # blocked tridiagonalisation forced onto small sizes
    ("blocked n70", wl.synthetic(70, 3), 3),
]
for name, p, nk in cases:
    if name.startswith("blocked"):
        os.environ["TBK_TRIDIAG_PANEL_MIN"] = "2"
    else:
        os.environ.pop("TBK_TRIDIAG_PANEL_MIN", None)
    ev = tbk.Evaluator(p, device=0)
    k = rng.random((nk, p.dim))
    e = ev.eigenval_array(k)
    h = ev.hamilton(k[: min(nk, 20)], convention=1)
    print(name, ev.path, float(e.sum()), float(np.abs(h).sum()), flush=True)
    ev.close()
# ---- round 2: eigenvectors, construct_kdotp, device-packed supercell, blocked -> staged hand-over, mesh host pipeline ----
os.environ.pop("TBK_TRIDIAG_PANEL_MIN", None)
for n_orb, nk in ((5, 9), (36, 7), (64, 3), (90, 2)):
    p = wl.synthetic(n_orb, 4, seed=n_orb)
    ev = tbk.Evaluator(p, device=0)
    w, v = ev.eigh(rng.random((nk, 3)))
    print("eigh", n_orb, float(w.sum()), float(np.abs(v).sum()), flush=True)
    ev.close()
p = wl.synthetic(12, 10, seed=3)
ev = tbk.Evaluator(p, device=0)
tc = ev.construct_kdotp((0.1, 0.2, 0.3), 2)
print("kdotp", len(tc), float(sum(np.abs(m).sum() for m in tc.values())), flush=True)
print("mesh host", float(ev.eigenval_mesh((3, 4, 16)).sum()), flush=True)
ev.close()
sup = tbk.KModel.from_packed(wl.synthetic(5, 6, seed=4)).supercell((2, 1, 2))
print("supercell", sup.size, float(np.array(sup.eigenval(rng.random((6, 3)))).sum()), flush=True)
sup.evaluator().close()
for n_orb in (128, 170, 264):
    p = wl.synthetic(n_orb, 2, seed=n_orb)
    ev = tbk.Evaluator(p, device=0)
    print("hand-over", n_orb, float(ev.eigenval_array(rng.random((3, 3))).sum()), flush=True)
    ev.close()
print("sanitize smoke done")
