#!/bin/bash
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
run() {  # workload nk G CS
  TBK_TRIDIAG_G=$3 TBK_TRIDIAG_CS=$4 timeout 300 python bench.py --workload $1 --no-extra --no-cpu --no-peaks --steps 2 --warmup 3 --nk $2 > $OUT/sweep.json 2>$OUT/sweep.err
  python -c "
import json
try:
    d=json.load(open('$OUT/sweep.json')); print('$1 G=$3 CS=$4', '%.4g'%d['value'], {k:round(v,2) for k,v in d['kernel_ms_per_step'].items()})
except Exception as e: print('$1 G=$3 CS=$4 FAILED', open('$OUT/sweep.err').read()[-300:])"
}
for gc in "0 1" "32 1" "64 2" "128 4" "64 1" "128 2"; do run c3 524288 $gc; done
for gc in "0 1" "128 1" "256 2" "512 4" "256 4" "256 1" "512 8"; do run c5 16384 $gc; done
run c4 2048 0 1
