#!/bin/bash
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "size_boundaries or synthetic_golden or batch_invariance or tridiag_variants or staged or c3_properties or degenerate" 2>&1 | tail -2
timeout 600 python bench.py --workload c3 --nk 1048576 --no-extra --no-cpu --no-peaks --steps 3 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c3', d['kernel_ms_per_step'])"
PYTHONPATH=. python tools/tridiag_sweep.py 24:65536 48:32768 64:32768 96:16384 110:8192 2>&1 | cut -c1-75 | tail -5
