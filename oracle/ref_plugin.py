"""TEST INFRASTRUCTURE ONLY -- pytest plugin that lets the reference's OWN test-suite run in this environment.

    python -m pytest <reference tests dir> -p oracle.ref_plugin [...]

Loaded with ``-p`` it runs before collection: the reference package becomes importable through ``oracle/ref_shim.py``
(h5py / fsc.hdf5_io stand-ins, numpy-2 aliases; regression data read through this repository's HDF5 reader) and, with
``TBK_REF_INSTALL=1``, ``tbmodels_b200.install()`` routes ``Model.hamilton`` / ``Model.eigenval`` /
``Model.construct_kdotp`` and ``KdotpModel.hamilton`` / ``eigenval`` of that package to the GPU -- so every assertion the
reference's own tests make about those methods (regression goldens, batched == per-k, supercell band folding, sparse ==
dense, slicing, arithmetic, k.p expansion ...) is made about the CUDA path.  Used by ``tests/test_reference_suite.py``.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402

tbmodels = ref_shim.import_reference()
if os.environ.get("TBK_REF_INSTALL") == "1":
    import tbmodels_b200

    tbmodels_b200.install()


def pytest_terminal_summary(terminalreporter):
    if os.environ.get("TBK_REF_INSTALL") == "1":
        from tbmodels_b200 import _patch

        terminalreporter.write_line("tbk-installed-calls: " + " ".join(f"{k}={v}" for k, v in _patch.stats.items()))


def pytest_report_header(config):
    return f"reference tbmodels from {ref_shim.reference_src()}, GPU methods installed: {os.environ.get('TBK_REF_INSTALL') == '1'}"
