#!/bin/bash
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT
PYTHONPATH=. timeout 900 python tools/tridiag_sweep.py --variants ${VARIANTS:-default,two_s1} ${SIZES:-512:2368} > $OUT/${TAG:-r04k}_sweep.log 2>&1
echo "rc=$?"; cat $OUT/${TAG:-r04k}_sweep.log | tail -40
PYTHONPATH=. timeout 900 python tools/tridiag_sweep.py --variants two256,two257,two512 ${SIZES2:-200:2368 256:2368} > $OUT/${TAG:-r04k}_sweep2.log 2>&1
echo "rc=$?"; cat $OUT/${TAG:-r04k}_sweep2.log | tail -40
