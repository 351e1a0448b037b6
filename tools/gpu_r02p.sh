#!/bin/bash
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT; TAG=${TAG:-r02p}
echo "== smoke"; timeout 600 python __graft_entry__.py --smoke > $OUT/${TAG}_smoke.log 2>&1; echo "rc=$?"; tail -5 $OUT/${TAG}_smoke.log
echo "== sweep panel vs staged smem at N~128"; PYTHONPATH=. timeout 600 python tools/tridiag_sweep.py --variants default,nopanel 112:16384 120:16384 128:16384 144:8192 > $OUT/${TAG}_sweep.log 2>&1; cut -c1-250 $OUT/${TAG}_sweep.log
echo "== done"
