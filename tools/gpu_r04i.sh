#!/bin/bash
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT
PYTHONPATH=. timeout 900 python tools/tridiag_sweep.py --variants default,two,two_s1 33:64 161:592 200:2368 256:2368 320:2368 384:2368 512:2368 512:8192 > $OUT/${TAG:-r04o}_sweep.log 2>&1
echo "rc=$?"; cat $OUT/${TAG:-r04o}_sweep.log | tail -40
