"""Throughput of the fused small-N path (N <= 8) at large batch: silicon (N = 8, 95 R) and synthetic N = 3..8."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tbmodels_b200 as tbk  # noqa: E402
from oracle import workloads as wl  # noqa: E402

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cases = [("silicon N=8 95R", wl.load_packed(os.path.join(root, "tests", "golden", "silicon.npz")), 2_000_000)]
for n in (3, 4, 6, 8):
    cases.append((f"synthetic N={n} 13R (nearest cells)", wl.synthetic(n, 13, seed=n), 4_000_000))
for name, p, nk in cases:
    ev = tbk.Evaluator(p, device=0)
    k = torch.rand((nk, p.dim), dtype=torch.float64, device="cuda")
    out = torch.empty((nk, p.size), dtype=torch.float64, device="cuda")
    for _ in range(2):
        ev.eigenval_device(k, out=out)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        ev.eigenval_device(k, out=out)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print(f"{name:40s} path={ev.path:12s} {nk / dt:.3e} k/s  ({dt * 1e3:.2f} ms)", flush=True)
    ev.close()
