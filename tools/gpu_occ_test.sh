#!/bin/bash
# Blocked tridiagonalisation, N = 512: time per matrix as a function of how many matrices (CTAs) run concurrently.
# 16 CTAs (everything L2 resident, no contention) 14.7 ms, 148 CTAs 18.0 ms: the kernel is latency bound inside the CTA.
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
for nk in 16 37 74 148 296; do
  timeout 600 python bench.py --workload c4 --nk $nk --no-extra --no-cpu --no-peaks --steps 3 --warmup 2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('nk $nk', d['kernel_ms_per_step'])"
done
