#!/bin/bash
# two-stage reduction: first correctness + timing sweep
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT
PYTHONPATH=. timeout 900 python tools/tridiag_sweep.py --variants ${VARIANTS:-default,two} ${SIZES:-12:64 17:64 24:64 33:64 40:64 64:64 100:64 161:592 200:592 256:592 300:296 384:296 512:296} > $OUT/${TAG:-r04b}_sweep.log 2>&1
echo "rc=$?"; cat $OUT/${TAG:-r04b}_sweep.log | tail -40
