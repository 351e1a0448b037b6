cd "${GRAFT_REPO_ROOT}"
for nk in 16 37 74 148 296; do
  timeout 600 python bench.py --workload c4 --nk $nk --no-extra --no-cpu --no-peaks --steps 3 --warmup 2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('nk $nk', d['kernel_ms_per_step'])"
done
