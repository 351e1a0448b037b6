#!/bin/bash
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT
PYTHONPATH=. timeout 1200 ncu --set full --clock-control none --import-source on -k regex:band_chase -c 1 -f -o $OUT/${TAG:-r04p}_chase python tools/tridiag_sweep.py --variants default 512:4736 > $OUT/${TAG:-r04p}_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 $OUT/${TAG:-r04p}_ncu.log
TAG=${TAG:-r04p} SIZES="512:148" bash tools/gpu_band_timing.sh
