#!/bin/bash
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
for mb in 16 64 256 1024; do
  TBK_HOST_CHUNK_MB=$mb timeout 600 python bench.py --no-extra --no-cpu --no-peaks --steps 5 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunk MB $mb e2e %.4g'%d['e2e']['value'], 'value %.4g'%d['value'])"
done
