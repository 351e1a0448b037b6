#!/bin/bash
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT; TAG=${TAG:-r03y}
for lim in 82 48 32 16; do echo "== smem max $lim"; TBK_EIGH_SMEM_MAX=$lim PYTHONPATH=. timeout 600 python tools/eigh_bench.py 24:65536 36:65536 48:32768 56:16384 64:16384 82:8192 2>&1 | tee -a $OUT/${TAG}_eigh.log; done
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "eigh" > $OUT/${TAG}_pytest.log 2>&1; echo "rc=$?"; tail -3 $OUT/${TAG}_pytest.log
