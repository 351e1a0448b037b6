#!/bin/bash
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider -k "supercell or size_boundaries" 2>&1 | tail -4
run() {  # workload nk mpb
  TBK_TRIDIAG_MPB=$3 timeout 600 python bench.py --workload $1 --no-extra --no-cpu --no-peaks --steps 2 --warmup 3 --nk $2 > $OUT/sweep.json 2>$OUT/sweep.err
  python -c "
import json
try:
    d=json.load(open('$OUT/sweep.json')); print('$1 MPB=$3', '%.4g'%d['value'], {k:round(v,2) for k,v in d['kernel_ms_per_step'].items()})
except Exception as e: print('$1 MPB=$3 FAILED', open('$OUT/sweep.err').read()[-300:])"
}
run c4 2048 0
for m in 4 5 6 8 9 12 16; do run c3 524288 $m; done
