#!/bin/bash
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT; TAG=${TAG:-r02r}
echo "== sweep G/CS/stages at N = 128, 144, 64, 96"; PYTHONPATH=. timeout 900 python tools/tridiag_sweep.py --variants default,g512c4,g512c8,g256c8,g256c2,st50,st80 64:32768 96:16384 128:16384 144:8192 > $OUT/${TAG}_sweep.log 2>&1; cut -c1-480 $OUT/${TAG}_sweep.log
echo "== pytest subset"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "size_boundaries or synthetic_golden or c5 or blocked or staged or degenerate or unblocked" > $OUT/${TAG}_pytest.log 2>&1; echo "rc=$?"; tail -4 $OUT/${TAG}_pytest.log
echo "== done"
