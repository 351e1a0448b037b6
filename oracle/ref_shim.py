"""TEST INFRASTRUCTURE ONLY -- import the unmodified reference ``tbmodels`` from /root/reference.

This module is part of the parity oracle: only ``tests/``, ``oracle/make_golden.py`` and the
``cpu_baseline`` leg of ``bench.py`` may use anything under ``oracle/``.  The product package
``tbmodels_b200`` never imports it.

``/root/reference`` exists only in the build container; on the GPU box the shim falls back to the byte-identical, sha256
verified archive of the package that ``oracle/build_ref.py`` places under ``oracle/_ref/`` (git-ignored, shipped like
``libtbk.so``; imported straight from the zip).

The reference (v1.4.4) does not import on this image as-is; the hot path itself needs only numpy and
scipy.  The shims:

* ``importlib.metadata.version("tbmodels")`` -> "1.4.4"       (src/tbmodels/__init__.py:7)
* stub ``h5py`` and ``fsc.hdf5_io`` modules                     (src/tbmodels/_tb_model.py:21,26; io.py:10-11)
* ``np.complex_`` / ``np.float_`` aliases removed in numpy 2    (src/tbmodels/_tb_model.py:1131,1150)
"""
from __future__ import annotations

import importlib
import importlib.metadata
import os
import sys
import types
import warnings

import numpy as np

REFERENCE_ROOT = os.environ.get("TBK_REFERENCE_ROOT", "/root/reference")
_REF_COPY = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def reference_src() -> str | None:
    """Directory to put on ``sys.path``: the reference tree itself, else the verified copy under ``oracle/_ref``."""
    if os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "tbmodels")):
        return os.path.join(REFERENCE_ROOT, "src")
    from . import build_ref

    return build_ref.package_path()


def reference_available() -> bool:
    return reference_src() is not None


def _install_stubs() -> None:
    if not hasattr(np, "complex_"):
        np.complex_ = np.complex128  # type: ignore[attr-defined]
    if not hasattr(np, "float_"):
        np.float_ = np.float64  # type: ignore[attr-defined]

    orig_version = importlib.metadata.version

    def version(name):  # noqa: D401
        if name == "tbmodels":
            return "1.4.4"
        return orig_version(name)

    importlib.metadata.version = version  # type: ignore[assignment]

    if "h5py" not in sys.modules:
        try:
            importlib.import_module("h5py")
        except ImportError:
            h5 = types.ModuleType("h5py")
            h5.File = object  # type: ignore[attr-defined]
            h5.Group = object  # type: ignore[attr-defined]
            sys.modules["h5py"] = h5
    if "fsc.hdf5_io" not in sys.modules:
        try:
            importlib.import_module("fsc.hdf5_io")
        except ImportError:
            fsc = types.ModuleType("fsc")
            hio = types.ModuleType("fsc.hdf5_io")

            class HDF5Enabled:  # plain base class, no behaviour needed on the hot path
                pass

            class SimpleHDF5Mapping(HDF5Enabled):
                pass

            def subscribe_hdf5(*_a, **_k):
                return lambda cls: cls

            def _unavailable(*_a, **_k):
                raise RuntimeError("fsc.hdf5_io is stubbed: HDF5 I/O is not available in this container")

            def _load(path):
                """Read-only stand-in for ``fsc.hdf5_io.load`` on the files the reference's test-suite keeps as
                regression data (type tags builtins.number / list / str and tbmodels.model): parsed with the HDF5
                reader of this repository (tbmodels_b200/_h5lite.py, no h5py), decoded by type tag.  ``OSError`` for a
                missing file, like h5py (tests/conftest.py:46-50 relies on it)."""
                if not os.path.exists(path):
                    raise OSError(f"Unable to open file {path!r}")
                from tbmodels_b200 import _h5lite

                def decode(node):
                    if not isinstance(node, dict) or "type_tag" not in node:
                        raise ValueError("no type_tag")  # tbmodels.io.load falls back to the legacy decoder on ValueError
                    tag = str(node["type_tag"])
                    if tag in ("builtins.number", "builtins.str"):
                        v = node["value"]
                        return str(v) if tag == "builtins.str" else v
                    if tag == "builtins.list":
                        return [decode(node[str(i)]) for i in range(len(node) - 1)]
                    if tag == "tbmodels.model":
                        tb = sys.modules["tbmodels"]
                        return tb.Model.from_hdf5({k: v for k, v in node.items() if k != "type_tag"})
                    raise RuntimeError(f"type tag {tag!r} is not supported by the test shim")

                return decode(_h5lite.load(path))

            hio.HDF5Enabled = HDF5Enabled
            hio.SimpleHDF5Mapping = SimpleHDF5Mapping
            hio.subscribe_hdf5 = subscribe_hdf5
            hio.save = _unavailable
            hio.load = _load
            hio.to_hdf5 = _unavailable
            hio.from_hdf5 = _unavailable
            hio.to_hdf5_file = _unavailable
            hio.from_hdf5_file = _unavailable
            fsc.hdf5_io = hio
            sys.modules["fsc"] = fsc
            sys.modules["fsc.hdf5_io"] = hio


def import_reference():
    """Return the reference ``tbmodels`` module (unmodified sources, shimmed imports)."""
    if "tbmodels" in sys.modules:
        return sys.modules["tbmodels"]
    src = reference_src()
    if src is None:
        raise ImportError(f"reference package found neither at {REFERENCE_ROOT} nor (verified archive) under {_REF_COPY}")
    _install_stubs()
    if src not in sys.path:
        sys.path.insert(0, src)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return importlib.import_module("tbmodels")
