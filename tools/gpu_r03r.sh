#!/bin/bash
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT; TAG=${TAG:-r03r}
echo "== plain"; timeout 600 python tools/sanitize_smoke.py > $OUT/${TAG}_smoke_plain.log 2>&1; echo "rc=$?"; tail -3 $OUT/${TAG}_smoke_plain.log
echo "== memcheck"; timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_smoke.py > $OUT/${TAG}_memcheck.log 2>&1; echo "rc=$?"; tail -3 $OUT/${TAG}_memcheck.log
echo "== racecheck"; timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_smoke.py > $OUT/${TAG}_racecheck.log 2>&1; echo "rc=$?"; tail -3 $OUT/${TAG}_racecheck.log
