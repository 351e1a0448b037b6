#!/bin/bash
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT; TAG=${TAG:-r02m}
echo "== sweep ql global"; PYTHONPATH=. timeout 900 python tools/tridiag_sweep.py --variants default,qlglobal 36:262144 64:65536 96:32768 128:16384 128:65536 256:4096 512:2048 > $OUT/${TAG}_sweep.log 2>&1; cut -c1-250 $OUT/${TAG}_sweep.log
echo "== done"
