#!/bin/bash
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT; TAG=${TAG:-r02s}
PYTHONPATH=. timeout 900 python tools/tridiag_sweep.py --variants default,st75,st80,st85,st90 48:65536 64:32768 80:16384 96:16384 112:16384 128:16384 160:8192 > $OUT/${TAG}_sweep.log 2>&1; cut -c1-400 $OUT/${TAG}_sweep.log
