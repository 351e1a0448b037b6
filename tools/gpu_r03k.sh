#!/bin/bash
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT; TAG=${TAG:-r03k}
echo "== pytest mesh"; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "mesh or exports" > $OUT/${TAG}_pytest.log 2>&1; echo "rc=$?"; tail -4 $OUT/${TAG}_pytest.log
echo "== bench c3 mesh"; timeout 600 python bench.py --workload c3 --mesh --no-extra --no-cpu --steps 3 --warmup 3 > $OUT/${TAG}_bench_c3m.json 2> $OUT/${TAG}_bench_c3m.err; tail -3 $OUT/${TAG}_bench_c3m.err
python -c "import sys,json; d=json.loads(open('$OUT/${TAG}_bench_c3m.json').read().strip().splitlines()[-1]); print('c3 mesh', d['value'], d['ms_per_step'], 'e2e', d['e2e'])"
