#!/bin/bash
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT
PYTHONPATH=. timeout 900 ncu --set full --clock-control none --import-source on -k regex:band_reduce -c 1 -f -o $OUT/${TAG:-r04e}_band python tools/tridiag_sweep.py --variants two_s1 512:296 > $OUT/${TAG:-r04e}_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 $OUT/${TAG:-r04e}_ncu.log
