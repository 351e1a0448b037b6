#!/bin/bash
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT; TAG=${TAG:-r02q}
echo "== sweep panel vs staged smem"; PYTHONPATH=. timeout 900 python tools/tridiag_sweep.py --variants default,panel200 112:16384 120:16384 128:16384 136:8192 144:8192 160:8192 > $OUT/${TAG}_sweep.log 2>&1; cut -c1-250 $OUT/${TAG}_sweep.log
echo "== done"
