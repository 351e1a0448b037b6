#!/bin/bash
# final-state evidence for the two-stage reduction: ncu of 4 band_reduce launches + 1 band_chase launch (N = 512, 4084
# matrices), per-phase cycles, sanitizer
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT; TAG=${TAG:-r04v}
PYTHONPATH=. timeout 1200 ncu --set full --clock-control none --import-source on -k regex:band_ -c 5 -f -o $OUT/${TAG}_c4_twostage python tools/tridiag_sweep.py --variants default 512:4084 > $OUT/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 $OUT/${TAG}_ncu.log
echo "== memcheck"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_twostage.py > $OUT/${TAG}_memcheck.log 2>&1; echo "rc=$?"; tail -2 $OUT/${TAG}_memcheck.log
echo "== racecheck"; timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_twostage.py > $OUT/${TAG}_racecheck.log 2>&1; echo "rc=$?"; tail -2 $OUT/${TAG}_racecheck.log
TAG=$TAG SIZES="512:148 256:296" bash tools/gpu_band_timing.sh
