"""Throughput of Evaluator.eigh_device for a few sizes (GPU box)."""
import sys
import numpy as np
import torch
import tbmodels_b200 as tbk
from oracle import workloads as wl

for spec in sys.argv[1:]:
    n, nk = (int(x) for x in spec.split(":"))
    p = wl.synthetic(n, 3, seed=n)
    ev = tbk.Evaluator(p, device=0)
    k = torch.rand((nk, 3), dtype=torch.float64, device="cuda")
    w = torch.empty((nk, n), dtype=torch.float64, device="cuda")
    v = torch.empty((nk, n, n), dtype=torch.complex128, device="cuda")
    ev.profile(True)
    for _ in range(2):
        ev.eigh_device(k, out_w=w, out_v=v)
    ev.profile_read()
    for _ in range(3):
        ev.eigh_device(k, out_w=w, out_v=v)
    pr = ev.profile_read()
    ev.check()
    ms = pr["eigh"][0] / 3
    ee = ev.eigenval_device(k)
    err = float((w - ee).abs().max())
    print(f"N={n:4d} nk={nk:7d} eigh {ms:9.3f} ms = {ms * 1000 / nk:8.4f} ms/1k  ({nk / ms * 1e3:.3e} k/s)  |w - eigenval| {err:.1e}", flush=True)
    ev.close()
