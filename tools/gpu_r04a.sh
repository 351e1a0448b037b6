#!/bin/bash
# eigh on global scratch by default: eigh tests + throughput table + bench c3_eigh extra
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT
echo skip-tests
PYTHONPATH=. timeout 600 python tools/eigh_bench.py 4:400000 8:400000 12:400000 16:200000 24:200000 36:200000 48:100000 64:50000 82:20000 128:8000 200:2000 > $OUT/r04a_eigh.log 2>&1; cat $OUT/r04a_eigh.log
