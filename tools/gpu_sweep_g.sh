#!/bin/bash
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
for g in ${SWEEP:-1 32}; do
  TBK_TRIDIAG_G=$g timeout 300 python bench.py --workload c3 --no-extra --no-cpu --no-peaks --steps 2 --warmup 3 --nk 524288 > $OUT/sweep_g$g.json 2>$OUT/sweep_g$g.err
  python -c "
import json; d=json.load(open('$OUT/sweep_g$g.json')); print('G=$g', d['value'], d['kernel_ms_per_step'])"
done
