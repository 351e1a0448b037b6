#!/bin/bash
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT; TAG=${TAG:-r03n}
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "size_boundaries or degenerate or blocked or unblocked or oversize or supercell_c4 or c4" > $OUT/${TAG}_pytest.log 2>&1; echo "rc=$?"; tail -6 $OUT/${TAG}_pytest.log
echo "== sweep"; PYTHONPATH=. timeout 900 python tools/tridiag_sweep.py --variants default,pstop0,pstop128 164:7104 200:4096 256:2368 384:1184 512:1184 > $OUT/${TAG}_sweep.log 2>&1; cut -c1-330 $OUT/${TAG}_sweep.log
