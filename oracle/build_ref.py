"""TEST INFRASTRUCTURE ONLY -- recipe that makes the UNMODIFIED reference available on the GPU box.

    python -m oracle.build_ref          (build container; /root/reference must exist)

The reference (Z2PackDev/TBmodels v1.4.4) is pure Python, so "building" it means packing byte-identical copies of its
package ``/root/reference/src/tbmodels`` and of its own test-suite ``/root/reference/tests`` (python files, samples,
regression data) into two archives under ``oracle/_ref/`` -- ``tbmodels_ref.zip`` and ``tests_ref.zip`` -- next to a manifest
with the sha256 of every member.  ``oracle/_ref/`` is git-ignored (reference sources never enter the history) but NOT
gpurun-ignored: the archives travel to the GPU box like the built ``libtbk.so``.  They are archives on purpose: no file of
the reference exists as a source file in this working tree; the package is imported straight from the zip
(``zipimport``), the tests are unpacked into a temporary directory for the duration of a test run.
``oracle/ref_shim.py`` imports the package from ``/root/reference/src`` when that exists and from the verified archive
otherwise; it is used by

* ``tests/test_gpu_reference_class.py`` -- ``tbmodels_b200.install()`` on the real ``tbmodels.Model`` on a B200;
* ``tests/test_reference_suite.py``     -- the reference's OWN tests against the installed GPU methods;
* ``bench.py --impl reference`` and its ``cpu_baseline`` leg -- the reference's own ``Model.eigenval`` on the host cores.

The product package never imports anything from here.
"""
from __future__ import annotations

import atexit
import hashlib
import json
import os
import shutil
import tempfile
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("TBK_REFERENCE_ROOT", "/root/reference")
REF_SRC = os.path.join(REF_ROOT, "src", "tbmodels")
REF_TESTS = os.path.join(REF_ROOT, "tests")  # the reference's own test-suite (+ its samples and regression data)
DEST = os.path.join(HERE, "_ref")
PKG_ZIP = os.path.join(DEST, "tbmodels_ref.zip")
TESTS_ZIP = os.path.join(DEST, "tests_ref.zip")
MANIFEST = os.path.join(DEST, "MANIFEST.json")
_SKIP = ("__pycache__",)


def _sha_bytes(data: bytes) -> str:
    return hashlib.sha256(data).hexdigest()


def _pack(src_dir: str, zip_path: str, prefix: str) -> dict:
    """Deterministic archive of ``src_dir`` (sorted members, fixed timestamps); returns {member: sha256}."""
    manifest = {}
    with zipfile.ZipFile(zip_path, "w", zipfile.ZIP_DEFLATED) as zf:
        for root, dirs, files in os.walk(src_dir):
            dirs[:] = sorted(d for d in dirs if d not in _SKIP)
            for name in sorted(files):
                if name.endswith(".pyc") or name.startswith("coverage"):
                    continue
                path = os.path.join(root, name)
                rel = os.path.join(prefix, os.path.relpath(path, src_dir)) if prefix else os.path.relpath(path, src_dir)
                with open(path, "rb") as f:
                    data = f.read()
                info = zipfile.ZipInfo(rel, date_time=(2020, 1, 1, 0, 0, 0))
                info.compress_type = zipfile.ZIP_DEFLATED
                info.external_attr = 0o644 << 16
                zf.writestr(info, data)
                manifest[rel] = _sha_bytes(data)
    return manifest


def build(verbose: bool = True) -> str | None:
    """Pack the reference package and its tests into ``oracle/_ref`` (no-op when the reference tree is absent)."""
    if not os.path.isdir(REF_SRC):
        if verbose:
            print(f"oracle.build_ref: {REF_SRC} not found; keeping whatever is under {DEST}")
        return DEST if os.path.exists(PKG_ZIP) else None
    os.makedirs(DEST, exist_ok=True)
    for stale in ("tbmodels", "tests", "MANIFEST_tests.json"):  # layouts of earlier versions of this recipe
        p = os.path.join(DEST, stale)
        if os.path.isdir(p):
            shutil.rmtree(p)
        elif os.path.exists(p):
            os.remove(p)
    man = {"source": REF_ROOT, "version": "1.4.4", "package": _pack(REF_SRC, PKG_ZIP, "tbmodels")}
    if os.path.isdir(REF_TESTS):
        man["tests"] = _pack(REF_TESTS, TESTS_ZIP, "")
    with open(MANIFEST, "w") as f:
        json.dump(man, f, indent=1, sort_keys=True)
    if verbose:
        print(f"oracle.build_ref: packed {len(man['package'])} files of the unmodified reference package into {PKG_ZIP}"
              + (f" and {len(man.get('tests', {}))} files of its test-suite into {TESTS_ZIP}" if "tests" in man else ""))
    return DEST


def _verify_zip(zip_path: str, members: dict) -> bool:
    if not os.path.exists(zip_path):
        return False
    try:
        with zipfile.ZipFile(zip_path) as zf:
            names = set(zf.namelist())
            return set(members) == names and all(_sha_bytes(zf.read(n)) == h for n, h in members.items())
    except (OSError, zipfile.BadZipFile):
        return False


_verified: dict = {}


def verify() -> bool:
    """True if the package archive exists and every member still has the sha256 recorded when it was packed."""
    if "package" not in _verified:
        ok = False
        if os.path.exists(MANIFEST):
            with open(MANIFEST) as f:
                ok = _verify_zip(PKG_ZIP, json.load(f).get("package", {}))
        _verified["package"] = ok
    return _verified["package"]


def package_path() -> str | None:
    """What to put on ``sys.path`` to import the archived reference package (a zip: imported through zipimport)."""
    return PKG_ZIP if verify() else None


_tests_tmp: list = []


def tests_dir() -> str | None:
    """The reference's own test-suite as a directory: the live tree when present, else the verified archive unpacked
    into a temporary directory (removed at interpreter exit)."""
    if os.path.isdir(REF_TESTS):
        return REF_TESTS
    if _tests_tmp:
        return _tests_tmp[0]
    if not os.path.exists(MANIFEST):
        return None
    with open(MANIFEST) as f:
        members = json.load(f).get("tests")
    if not members or not _verify_zip(TESTS_ZIP, members):
        return None
    tmp = tempfile.mkdtemp(prefix="tbk_ref_tests_")
    with zipfile.ZipFile(TESTS_ZIP) as zf:
        zf.extractall(tmp)
    atexit.register(shutil.rmtree, tmp, ignore_errors=True)
    _tests_tmp.append(tmp)
    return tmp


if __name__ == "__main__":
    build()
