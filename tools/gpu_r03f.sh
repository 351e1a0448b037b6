#!/bin/bash
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT; TAG=${TAG:-r03f}
PYTHONPATH=. timeout 900 python tools/tridiag_sweep.py --variants default,st80,st67 72:28416 80:14208 96:14208 112:14208 128:14208 144:7104 160:7104 > $OUT/${TAG}_sweep.log 2>&1; cut -c1-330 $OUT/${TAG}_sweep.log

