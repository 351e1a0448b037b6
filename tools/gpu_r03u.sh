#!/bin/bash
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT; TAG=${TAG:-r03u}
echo "== sweep"; PYTHONPATH=. timeout 900 python tools/tridiag_sweep.py --variants default,lalite 24:262144 28:131072 32:131072 36:227328 40:65536 48:65536 > $OUT/${TAG}_sweep.log 2>&1; cut -c1-260 $OUT/${TAG}_sweep.log
