#!/bin/bash
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT; TAG=${TAG:-r02y}
echo "== plain parity"; timeout 600 python tools/sanitize_reg.py > $OUT/${TAG}_reg_plain.log 2>&1; echo "rc=$?"; tail -2 $OUT/${TAG}_reg_plain.log
echo "== sweep"; PYTHONPATH=. timeout 900 python tools/tridiag_sweep.py --variants default,nola,reg40 36:131072 37:65536 40:65536 44:65536 48:65536 56:32768 64:32768 > $OUT/${TAG}_sweep.log 2>&1; cut -c1-330 $OUT/${TAG}_sweep.log
echo "== pytest gpu (full)"; timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider > $OUT/${TAG}_pytest.log 2>&1; echo "rc=$?"; tail -4 $OUT/${TAG}_pytest.log
echo "== done"
