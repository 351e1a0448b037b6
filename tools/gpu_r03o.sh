#!/bin/bash
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT; TAG=${TAG:-r03o}
PYTHONPATH=. timeout 900 python tools/tridiag_sweep.py --variants pstop80,pstop96,pstop112,pstop128,pstop144,default 164:7104 200:4096 256:2368 384:1184 512:1184 > $OUT/${TAG}_sweep.log 2>&1; cut -c1-560 $OUT/${TAG}_sweep.log
