#!/bin/bash
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT; TAG=${TAG:-r02l}
echo "== sweep ql/bisect"; PYTHONPATH=. timeout 900 python tools/tridiag_sweep.py --variants ql129,bisect 48:65536 64:32768 80:16384 96:16384 100:16384 104:16384 112:8192 120:8192 128:8192 > $OUT/${TAG}_sweep.log 2>&1; cut -c1-250 $OUT/${TAG}_sweep.log
echo "== pytest subset"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "size_boundaries or synthetic_golden or degenerate or c5 or blocked or mesh" > $OUT/${TAG}_pytest.log 2>&1; echo "rc=$?"; tail -4 $OUT/${TAG}_pytest.log
echo "== done"
