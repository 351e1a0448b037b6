#!/bin/bash
# QL vs bisection crossover; full GPU suite; default bench
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT; TAG=${TAG:-r02k}
echo "== sweep ql/bisect"; PYTHONPATH=. timeout 900 python tools/tridiag_sweep.py --variants default,bisect 36:131072 48:65536 64:32768 80:16384 96:16384 112:8192 128:8192 > $OUT/${TAG}_sweep.log 2>&1; cut -c1-250 $OUT/${TAG}_sweep.log
echo "== pytest gpu"; timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider > $OUT/${TAG}_pytest.log 2>&1; echo "rc=$?"; tail -4 $OUT/${TAG}_pytest.log
echo "== done"
