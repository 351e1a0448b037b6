"""TEST INFRASTRUCTURE ONLY -- golden vectors of ``Model.construct_kdotp`` from the unmodified reference.

    python -m oracle.make_golden_kdotp        (build container)   ->  tests/golden/construct_kdotp.npz

For a few models (the 2-band fixture of the reference suite in 3-D, the Haldane model, the silicon Wannier model,
synthetic N = 5 / 12 / 36 models) and expansion points: the power tuples and coefficient matrices of the reference's
``construct_kdotp(k, order)`` (src/tbmodels/_tb_model.py:942-982), plus ``KdotpModel.eigenval`` of the constructed model at
a few offsets (src/tbmodels/kdotp.py:84-100).  The packed arrays travel with the file so the GPU test needs nothing else.
"""
from __future__ import annotations

import os
import warnings

import numpy as np

from oracle.make_golden import GOLD, packed_arrays, simple_model
from oracle.ref_shim import import_reference


def ref_from_packed(tb, q):
    return tb.Model(hop={tuple(int(x) for x in r): np.array(h) for r, h in zip(q.R, q.hop)}, pos=np.array(q.pos),
                    size=q.size, dim=q.dim, contains_cc=False)


def main():
    warnings.simplefilter("ignore")
    tb = import_reference()
    from oracle import workloads as wl

    rng = np.random.default_rng(942)
    cases = {
        "simple3d": (simple_model(tb, 0.2, 0.5, dim=3), 3),
        "haldane": (ref_from_packed(tb, wl.haldane()), 2),
        "silicon": (ref_from_packed(tb, wl.load_packed(os.path.join(GOLD, "silicon.npz"))), 2),
        "syn5": (ref_from_packed(tb, wl.synthetic(5, 7, seed=11)), 3),
        "syn12": (ref_from_packed(tb, wl.synthetic(12, 30, seed=12)), 2),
        "syn36": (ref_from_packed(tb, wl.synthetic(36, 60, seed=13)), 1),
        "syn2d": (ref_from_packed(tb, wl.synthetic(9, 12, seed=14, dim=2)), 4),
    }
    out = {"names": np.array(sorted(cases))}
    for name, (m, order) in cases.items():
        pr = packed_arrays(m)
        ks = np.vstack([np.zeros(m.dim), rng.uniform(-0.5, 1.5, size=(2, m.dim))])
        out[f"{name}_R"], out[f"{name}_hop"], out[f"{name}_pos"] = pr["R"], pr["hop"], pr["pos"]
        out[f"{name}_order"] = np.array(order)
        out[f"{name}_k"] = ks
        for i, k in enumerate(ks):
            kp = m.construct_kdotp(k, order)
            keys = list(kp.taylor_coefficients)
            out[f"{name}_powers"] = np.array(keys, dtype=np.int32).reshape(len(keys), m.dim)
            out[f"{name}_coeff{i}"] = np.stack([kp.taylor_coefficients[key] for key in keys])
            dk = rng.uniform(-0.05, 0.05, size=(4, m.dim))
            out[f"{name}_dk{i}"] = dk
            out[f"{name}_eig{i}"] = np.array(kp.eigenval(dk))
    np.savez_compressed(os.path.join(GOLD, "construct_kdotp.npz"), **out)
    print("wrote construct_kdotp.npz:", ", ".join(out["names"]))


if __name__ == "__main__":
    main()
