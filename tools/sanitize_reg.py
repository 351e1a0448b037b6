"""Register-resident tridiagonalisation (eig_tridiag_reg.cu) at every template instantiation and hand-over size, meant to
be wrapped in compute-sanitizer; also prints the error against numpy so a plain run is a quick parity check."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["TBK_TRIDIAG_REG_MIN"] = "2"
os.environ["TBK_TRIDIAG_REG_MAX"] = "48"
os.environ["TBK_FORCE_GEMM"] = "1"
import tbmodels_b200 as tbk  # noqa: E402
from oracle import tb_oracle as orc  # noqa: E402
from oracle import workloads as wl  # noqa: E402

rng = np.random.default_rng(0)
worst = 0.0
for n in (2, 3, 9, 12, 13, 16, 20, 21, 24, 25, 28, 29, 32, 33, 34, 35, 36, 37, 40, 41, 44, 45, 48, 49, 64, 97):
    p = wl.synthetic(n, 4, seed=n)
    nk = 37 if n <= 48 else 9
    k = rng.random((nk, p.dim))
    ev = tbk.Evaluator(p, device=0)
    e = ev.eigenval_array(k)
    want = orc.eigenval_array(p.R, p.hop, p.pos, k)
    err = float(np.abs(e - want).max() / np.abs(want).max())
    worst = max(worst, err)
    print(f"N={n:3d} rel err {err:.2e}", flush=True)
    ev.close()
print("sanitize reg done, worst", worst)
assert worst < 1e-10
