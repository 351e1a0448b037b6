#!/bin/bash
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
for r in 67 60 75 55; do
  TBK_TRIDIAG_STAGES=$r timeout 600 python bench.py --workload c3 --nk 1048576 --no-extra --no-cpu --no-peaks --steps 3 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ratio $r', round(d['kernel_ms_per_step']['tridiag'],2), round(d['ms_per_step'],1))"
done
for mb in 1024 4096 8192; do
  TBK_WORKSPACE_MB=$mb timeout 600 python bench.py --workload c3 --nk 2097152 --no-extra --no-cpu --no-peaks --steps 2 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('workspace $mb MB', round(d['ms_per_step'],1), d['kernel_ms_per_step'])"
done
