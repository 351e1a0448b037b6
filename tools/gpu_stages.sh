#!/bin/bash
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-st}
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider -k "size_boundaries or synthetic_golden or batch_invariance or tridiag_variants or c3_properties or mesh or staged or unblocked" > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $OUT/${TAG}_pytest.log
for r in 0 67 75 60 50; do
  echo "ratio $r"
  TBK_TRIDIAG_STAGES=$r PYTHONPATH=. python tools/tridiag_sweep.py 30:65536 36:131072 48:32768 64:32768 96:16384 110:8192 2>&1 | cut -c1-75 | tail -6
done
