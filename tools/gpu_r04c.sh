#!/bin/bash
# two-stage reduction: stage split + ncu of both kernels
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT
PYTHONPATH=. timeout 900 python tools/tridiag_sweep.py --variants default,two,two_s1 256:2368 512:296 512:2368 > $OUT/r04c_sweep.log 2>&1
echo "rc=$?"; cat $OUT/r04c_sweep.log | tail -40
PYTHONPATH=. timeout 900 ncu --set full --clock-control none --import-source on -k regex:band_ -c 2 -f -o $OUT/r04c_band python tools/tridiag_sweep.py --variants two 512:296 > $OUT/r04c_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 $OUT/r04c_ncu.log
