"""Shared fixtures.  ``-m gpu`` tests need a B200 and call through the C ABI; everything else runs on CPU."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# parity bounds stated by BASELINE.json's north_star
H_TOL = 1e-11  # max|dH| <= H_TOL * max|H_R|
EIG_TOL = 1e-10  # max|d lambda| <= EIG_TOL * spectral radius


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name)) as f:
        return {k: f[k] for k in f.files}


@pytest.fixture(scope="session")
def golden():
    return load_golden


def packed_from(d, prefix=""):
    from tbmodels_b200 import pack_arrays

    return pack_arrays(d[prefix + "R"], d[prefix + "hop"], d[prefix + "pos"])


def h_scale(packed):
    return max(float(np.abs(packed.hop).max()) if packed.hop.size else 0.0, 1e-300)


def assert_h_close(got, want, packed, what=""):
    err = float(np.abs(np.asarray(got) - np.asarray(want)).max()) if np.asarray(want).size else 0.0
    bound = H_TOL * h_scale(packed)
    assert err <= bound, f"{what}: max|dH| = {err:.3e} > {bound:.3e}"


def assert_eig_close(got, want, what=""):
    got = np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape, f"{what}: shape {got.shape} != {want.shape}"
    rho = max(float(np.abs(want).max()) if want.size else 0.0, 1e-300)
    err = float(np.abs(got - want).max()) if want.size else 0.0
    assert err <= EIG_TOL * rho, f"{what}: max|d lambda| = {err:.3e} > {EIG_TOL * rho:.3e}"


def gpu_available():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False
