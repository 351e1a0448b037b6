#!/bin/bash
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT
PYTHONPATH=. timeout 900 python tools/tridiag_sweep.py --variants default,two,two_s1 17:64 33:64 100:64 161:592 200:592 256:2368 300:296 512:296 512:2368 > $OUT/${TAG:-r04g}_sweep.log 2>&1
echo "rc=$?"; cat $OUT/${TAG:-r04g}_sweep.log | tail -40
