"""Eigensolver timing sweep over N with the dispatch overridden through the tuning environment variables.

    python tools/tridiag_sweep.py N:nk ...      (GPU box)
Prints ms per 1000 matrices for: default dispatch, the unblocked single-launch kernels, the blocked kernel with
128 / 256 / 512 threads per matrix.
"""
import os
import sys
import time

import numpy as np
import torch

import tbmodels_b200 as tbk
from tbmodels_b200 import workloads as wl

VARIANTS = {
    "default": {},
    "smem/old": {"TBK_TRIDIAG_NOPANEL": "1", "TBK_TRIDIAG_PANEL_MIN": "100000", "TBK_TRIDIAG_STAGES": "0"},
    "panel128": {"TBK_TRIDIAG_PANEL_MIN": "2", "TBK_PANEL_T": "128"},
    "panel256": {"TBK_TRIDIAG_PANEL_MIN": "2", "TBK_PANEL_T": "256"},
    "panel512": {"TBK_TRIDIAG_PANEL_MIN": "2", "TBK_PANEL_T": "512"},
}
KEYS = sorted({k for v in VARIANTS.values() for k in v})


def main():
    for spec in sys.argv[1:]:
        n, nk = (int(x) for x in spec.split(":"))
        p = wl.synthetic(n, 3, seed=n)
        ev = tbk.Evaluator(p, device=0)
        ev.profile(True)
        k = torch.rand((nk, p.dim), dtype=torch.float64, device="cuda")
        out = torch.empty((nk, n), dtype=torch.float64, device="cuda")
        ref = None
        line = [f"N={n:4d} nk={nk:7d}"]
        for name, env in VARIANTS.items():
            for key in KEYS:
                os.environ.pop(key, None)
            os.environ.update(env)
            try:
                for _ in range(2):
                    ev.eigenval_device(k, out=out)
                ev.profile_read()
                reps = 3
                for _ in range(reps):
                    ev.eigenval_device(k, out=out)
                torch.cuda.synchronize()
                pr = ev.profile_read()
                ms = {c: v[0] / reps for c, v in pr.items() if v[1]}
                got = out.cpu().numpy()
                if ref is None:
                    ref = got
                err = float(np.abs(got - ref).max())
                line.append(f"{name}: tridiag {ms.get('tridiag', float('nan')) * 1000 / nk:8.3f} ms/1k (ql {ms.get('ql', 0) * 1000 / nk:6.3f}) d={err:.1e}")
            except Exception as e:  # noqa: BLE001
                line.append(f"{name}: FAILED {e}")
        print(" | ".join(line), flush=True)
        ev.close()


if __name__ == "__main__":
    main()
