#!/bin/bash
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT; TAG=${TAG:-r03d}
PYTHONPATH=. timeout 900 python tools/tridiag_sweep.py --variants default,g1_512c4,g1_512c8,g1_256c8,g1_256c2,g1_128c4 128:14208 144:7104 160:7104 > $OUT/${TAG}_sweep.log 2>&1; cut -c1-420 $OUT/${TAG}_sweep.log
