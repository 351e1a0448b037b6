"""TEST INFRASTRUCTURE ONLY -- recipe that makes the UNMODIFIED reference importable on the GPU box.

    python -m oracle.build_ref          (build container; /root/reference must exist)

The reference (Z2PackDev/TBmodels v1.4.4) is pure Python, so "building" it means placing a byte-identical copy of its
package ``/root/reference/src/tbmodels`` under ``oracle/_ref/tbmodels`` (git-ignored: reference sources never enter the
history; NOT gpurun-ignored: the directory travels to the GPU box like the built ``libtbk.so``).  A manifest with the
sha256 of every file is written next to it so tests can check that nothing was edited.  ``oracle/ref_shim.py`` imports
the package from ``/root/reference/src`` when that exists and from ``oracle/_ref`` otherwise; it is used by

* ``tests/test_gpu_reference_class.py`` -- ``tbmodels_b200.install()`` on the real ``tbmodels.Model`` on a B200;
* ``bench.py --impl reference`` and its ``cpu_baseline`` leg -- the reference's own ``Model.eigenval`` on the host cores.

The product package never imports anything from here.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("TBK_REFERENCE_ROOT", "/root/reference")
REF_SRC = os.path.join(REF_ROOT, "src", "tbmodels")
REF_TESTS = os.path.join(REF_ROOT, "tests")  # the reference's own test-suite (+ its samples and regression data)
DEST = os.path.join(HERE, "_ref")


def _sha(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def build(verbose: bool = True) -> str | None:
    """Copy the reference package to ``oracle/_ref/tbmodels`` (no-op when the reference tree is absent)."""
    if not os.path.isdir(REF_SRC):
        if verbose:
            print(f"oracle.build_ref: {REF_SRC} not found; keeping whatever is under {DEST}")
        return DEST if os.path.isdir(os.path.join(DEST, "tbmodels")) else None
    pkg = os.path.join(DEST, "tbmodels")
    if os.path.isdir(pkg):
        shutil.rmtree(pkg)
    os.makedirs(DEST, exist_ok=True)
    shutil.copytree(REF_SRC, pkg, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    manifest = {}
    for root, _dirs, files in os.walk(pkg):
        for name in sorted(files):
            p = os.path.join(root, name)
            rel = os.path.relpath(p, pkg)
            manifest[rel] = _sha(p)
            assert manifest[rel] == _sha(os.path.join(REF_SRC, rel)), rel
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF_SRC, "version": "1.4.4", "sha256": manifest}, f, indent=1, sort_keys=True)
    n_tests = 0
    if os.path.isdir(REF_TESTS):  # the reference's own tests, run against the installed GPU methods by tests/test_reference_suite.py
        tdst = os.path.join(DEST, "tests")
        if os.path.isdir(tdst):
            shutil.rmtree(tdst)
        shutil.copytree(REF_TESTS, tdst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "coverage*.sh"))
        tman = {}
        for root, _dirs, files in os.walk(tdst):
            for name in sorted(files):
                if name.endswith(".py"):
                    pth = os.path.join(root, name)
                    tman[os.path.relpath(pth, tdst)] = _sha(pth)
        with open(os.path.join(DEST, "MANIFEST_tests.json"), "w") as f:
            json.dump({"source": REF_TESTS, "sha256": tman}, f, indent=1, sort_keys=True)
        n_tests = len(tman)
    if verbose:
        print(f"oracle.build_ref: copied {len(manifest)} files of the unmodified reference to {pkg}"
              + (f" and its test-suite ({n_tests} python files + data) to {os.path.join(DEST, 'tests')}" if n_tests else ""))
    return DEST


def tests_dir() -> str | None:
    """The reference's own test-suite: the live tree when present, else the verified copy under ``oracle/_ref/tests``."""
    if os.path.isdir(REF_TESTS):
        return REF_TESTS
    tdst = os.path.join(DEST, "tests")
    mf = os.path.join(DEST, "MANIFEST_tests.json")
    if not (os.path.isdir(tdst) and os.path.exists(mf)):
        return None
    with open(mf) as f:
        tman = json.load(f)["sha256"]
    ok = all(os.path.exists(os.path.join(tdst, rel)) and _sha(os.path.join(tdst, rel)) == h for rel, h in tman.items())
    return tdst if ok else None


def verify() -> bool:
    """True if ``oracle/_ref/tbmodels`` exists and every file still has the sha256 recorded at copy time."""
    mf = os.path.join(DEST, "MANIFEST.json")
    if not os.path.exists(mf):
        return False
    with open(mf) as f:
        manifest = json.load(f)["sha256"]
    return all(os.path.exists(os.path.join(DEST, "tbmodels", rel)) and _sha(os.path.join(DEST, "tbmodels", rel)) == h
               for rel, h in manifest.items())


if __name__ == "__main__":
    build()
