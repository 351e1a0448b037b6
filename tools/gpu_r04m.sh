#!/bin/bash
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT; TAG=${TAG:-r04m}
timeout 1700 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "twostage or size_boundaries or blocked or unblocked_large or c4 or supercell or staged or mesh" > $OUT/${TAG}_pytest.log 2>&1; echo "rc=$?"; tail -6 $OUT/${TAG}_pytest.log
echo "== memcheck"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_twostage.py > $OUT/${TAG}_memcheck.log 2>&1; echo "rc=$?"; tail -4 $OUT/${TAG}_memcheck.log
echo "== racecheck"; timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_twostage.py > $OUT/${TAG}_racecheck.log 2>&1; echo "rc=$?"; tail -4 $OUT/${TAG}_racecheck.log
