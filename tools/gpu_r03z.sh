#!/bin/bash
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT; TAG=${TAG:-r03z}
for lim in 48 0; do echo "== smem max $lim"; TBK_EIGH_SMEM_MAX=$lim PYTHONPATH=. timeout 600 python tools/eigh_bench.py 4:262144 8:262144 12:262144 16:131072 20:131072 24:65536 36:262144 2>&1 | tee -a $OUT/${TAG}_eigh.log; done
