"""Copy the reference's HDF5 test DATA for the `eigenvals` CLI path into tests/golden/cli_eigenvals/ (run in the build
container, where /root/reference exists).  These are data fixtures of tests/test_cli_eigenvals.py, not sources:

    silicon_model.hdf5      input model           (tests/test_cli_eigenvals.py:40)
    kpoints.hdf5            explicit k-point list (:22)
    silicon_eigenvals.hdf5  expected output, also accepted as k-point input (:22, :47)
"""
import os
import shutil

SRC = "/root/reference/tests/samples/cli_eigenvals"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "cli_eigenvals")

if __name__ == "__main__":
    os.makedirs(DST, exist_ok=True)
    for name in ("silicon_model.hdf5", "kpoints.hdf5", "silicon_eigenvals.hdf5"):
        shutil.copyfile(os.path.join(SRC, name), os.path.join(DST, name))
        os.chmod(os.path.join(DST, name), 0o644)
        print("copied", name, os.path.getsize(os.path.join(DST, name)))
