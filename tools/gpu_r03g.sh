#!/bin/bash
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT; TAG=${TAG:-r03g}
timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:"tridiag_smem_kernel" -s 15 -c 2 -o $OUT/${TAG}_c5_smem python bench.py --workload c5 --nk 14208 --steps 1 --warmup 3 --no-cpu --no-peaks --no-extra > $OUT/${TAG}_c5.log 2>&1; tail -1 $OUT/${TAG}_c5.log | cut -c1-100
timeout 600 python bench.py --workload c5 --nk 16384 --steps 3 --warmup 3 --no-cpu --no-peaks --no-extra > $OUT/${TAG}_bench_c5.json 2>/dev/null
python -c "import sys,json; d=json.loads(open('$OUT/${TAG}_bench_c5.json').read().strip().splitlines()[-1]); print('c5', d['value'], d['ms_per_step'], d['kernel_ms_per_step'])"
