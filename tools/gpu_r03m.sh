#!/bin/bash
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT; TAG=${TAG:-r03m}
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "ql_two or size_boundaries or degenerate or eigh" > $OUT/${TAG}_pytest.log 2>&1; echo "rc=$?"; tail -6 $OUT/${TAG}_pytest.log
echo "== sweep"; PYTHONPATH=. timeout 900 python tools/tridiag_sweep.py --variants default,qlpair 24:262144 36:227328 48:65536 64:56832 80:28416 96:28416 112:14208 128:14208 128:16384 > $OUT/${TAG}_sweep.log 2>&1; cut -c1-260 $OUT/${TAG}_sweep.log
