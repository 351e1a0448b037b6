"""Eigensolver timing sweep over N with the dispatch overridden through the tuning environment variables.

    python tools/tridiag_sweep.py [--variants a,b,..] N:nk ...      (GPU box)
Prints ms per 1000 matrices per variant.  The hooks are captured when a handle is created, so every variant gets its own
Evaluator.
"""
import os
import sys

import numpy as np
import torch

import tbmodels_b200 as tbk
from oracle import workloads as wl

VARIANTS = {
    "default": {},
    "noreg": {"TBK_TRIDIAG_REG_MAX": "0"},
    "reg48": {"TBK_TRIDIAG_REG_MAX": "48", "TBK_TRIDIAG_REG_MIN": "2"},
    "bw8": {"TBK_TRIDIAG_REG_BW": "8"},
    "nola": {"TBK_TRIDIAG_REG_BW": "1"},
    "la168": {"TBK_TRIDIAG_REG_BW": "2"},
    "reg40": {"TBK_TRIDIAG_REG_MAX": "40"},
    "smem/old": {"TBK_TRIDIAG_REG_MAX": "0", "TBK_TRIDIAG_NOPANEL": "1", "TBK_TRIDIAG_PANEL_MIN": "100000", "TBK_TRIDIAG_STAGES": "0"},
    "panel128": {"TBK_TRIDIAG_PANEL_MIN": "2", "TBK_PANEL_T": "128"},
    "panel256": {"TBK_TRIDIAG_PANEL_MIN": "2", "TBK_PANEL_T": "256"},
    "panel512": {"TBK_TRIDIAG_PANEL_MIN": "2", "TBK_PANEL_T": "512"},
    "bisect": {"TBK_QL_BISECT_MIN": "2"},
    "ql129": {"TBK_QL_BISECT_MIN": "129"},
    "qlglobal": {"TBK_QL_GLOBAL_MIN": "2"},
    "stop0": {"TBK_TRIDIAG_REG_STOP": "0"},
    "stop8": {"TBK_TRIDIAG_REG_STOP": "8"},
    "stop12": {"TBK_TRIDIAG_REG_STOP": "12"},
    "stop16": {"TBK_TRIDIAG_REG_STOP": "16"},
    "half16": {"TBK_TRIDIAG_REG_MIN": "2"},
    "nopanel": {"TBK_TRIDIAG_NOPANEL": "1"},
    "pstop0": {"TBK_TRIDIAG_PANEL_STOP": "0"},
    "pstop128": {"TBK_TRIDIAG_PANEL_STOP": "128"},
    "pstop112": {"TBK_TRIDIAG_PANEL_STOP": "112"},
    "pstop144": {"TBK_TRIDIAG_PANEL_STOP": "144"},
    "pstop80": {"TBK_TRIDIAG_PANEL_STOP": "80"},
    "pstop96": {"TBK_TRIDIAG_PANEL_STOP": "96"},
    "noovl": {"TBK_QL_OVERLAP": "0"},
    "panel200": {"TBK_TRIDIAG_PANEL_MIN": "200"},
    "panel120": {"TBK_TRIDIAG_PANEL_MIN": "120"},
    "g1_512c4": {"TBK_TRIDIAG_G1": "512", "TBK_TRIDIAG_CS1": "4"},
    "g1_512c8": {"TBK_TRIDIAG_G1": "512", "TBK_TRIDIAG_CS1": "8"},
    "g1_256c8": {"TBK_TRIDIAG_G1": "256", "TBK_TRIDIAG_CS1": "8"},
    "g1_256c2": {"TBK_TRIDIAG_G1": "256", "TBK_TRIDIAG_CS1": "2"},
    "g1_128c4": {"TBK_TRIDIAG_G1": "128", "TBK_TRIDIAG_CS1": "4"},
    "g512c4": {"TBK_TRIDIAG_G": "512", "TBK_TRIDIAG_CS": "4"},
    "g512c8": {"TBK_TRIDIAG_G": "512", "TBK_TRIDIAG_CS": "8"},
    "g256c8": {"TBK_TRIDIAG_G": "256", "TBK_TRIDIAG_CS": "8"},
    "g256c2": {"TBK_TRIDIAG_G": "256", "TBK_TRIDIAG_CS": "2"},
    "st50": {"TBK_TRIDIAG_STAGES": "50"},
    "st80": {"TBK_TRIDIAG_STAGES": "80"},
    "st67": {"TBK_TRIDIAG_STAGES": "67"},
    "st75": {"TBK_TRIDIAG_STAGES": "75"},
    "st85": {"TBK_TRIDIAG_STAGES": "85"},
    "st90": {"TBK_TRIDIAG_STAGES": "90"},
    "mid0": {"TBK_TRIDIAG_REG_MID": "0"},
    "mid20": {"TBK_TRIDIAG_REG_MID": "20"},
    "mid24": {"TBK_TRIDIAG_REG_MID": "24"},
    "one": {"TBK_TRIDIAG_TWOSTAGE": "0"},
    "chase1": {"TBK_BAND_CHASE": "1"},
    "chase4": {"TBK_BAND_CHASE": "4"},
    "two_chase1": {"TBK_TRIDIAG_TWOSTAGE": "12", "TBK_BAND_CHASE": "1"},
    "two": {"TBK_TRIDIAG_TWOSTAGE": "12"},
    "t128": {"TBK_BAND_T": "128"},
    "two256": {"TBK_TRIDIAG_TWOSTAGE": "12", "TBK_BAND_T": "256"},
    "two512": {"TBK_TRIDIAG_TWOSTAGE": "12", "TBK_BAND_T": "512"},
    "w4736": {"TBK_BAND_WAVE": "4736"},
    "w7104": {"TBK_BAND_WAVE": "7104"},
    "w8192": {"TBK_BAND_WAVE": "8192"},
    "w100000": {"TBK_BAND_WAVE": "100000"},
    "g256": {"TBK_BAND_GROUP_MB": "256"},
    "g512": {"TBK_BAND_GROUP_MB": "512"},
    "g768": {"TBK_BAND_GROUP_MB": "768"},
    "g1024": {"TBK_BAND_GROUP_MB": "1024"},
    "g1280": {"TBK_BAND_GROUP_MB": "1280"},
    "g2048": {"TBK_BAND_GROUP_MB": "2048"},
    "two257": {"TBK_TRIDIAG_TWOSTAGE": "12", "TBK_BAND_T": "257"},
    "two_s1": {"TBK_TRIDIAG_TWOSTAGE": "12", "TBK_BAND_STAGE2": "0"},
}
KEYS = sorted({k for v in VARIANTS.values() for k in v})


def main():
    args = sys.argv[1:]
    names = ["default", "noreg", "reg48"]
    if args and args[0] == "--variants":
        names = args[1].split(",")
        args = args[2:]
    for spec in args:
        n, nk = (int(x) for x in spec.split(":"))
        p = wl.synthetic(n, 3, seed=n)
        k = torch.rand((nk, p.dim), dtype=torch.float64, device="cuda")
        out = torch.empty((nk, n), dtype=torch.float64, device="cuda")
        ref = None
        line = [f"N={n:4d} nk={nk:7d}"]
        for name in names:
            for key in KEYS:
                os.environ.pop(key, None)
            os.environ.update(VARIANTS[name])
            try:
                ev = tbk.Evaluator(p, device=0)
                ev.profile(True)
                for _ in range(2):
                    ev.eigenval_device(k, out=out)
                ev.profile_read()
                reps = 3
                for _ in range(reps):
                    ev.eigenval_device(k, out=out)
                torch.cuda.synchronize()
                pr = ev.profile_read()
                ev.check()
                ms = {c: v[0] / reps for c, v in pr.items() if v[1]}
                got = out.cpu().numpy()
                if ref is None:
                    ref = got
                err = float(np.abs(got - ref).max())
                line.append(f"{name}: tridiag {ms.get('tridiag', float('nan')) * 1000 / nk:8.3f} ms/1k (ql {ms.get('ql', 0) * 1000 / nk:6.3f}) d={err:.1e}")
                ev.close()
            except Exception as e:  # noqa: BLE001
                line.append(f"{name}: FAILED {e}")
        print(" | ".join(line), flush=True)
    for key in KEYS:
        os.environ.pop(key, None)


if __name__ == "__main__":
    main()
