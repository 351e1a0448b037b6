#!/bin/bash
# round 2, session 2 baseline: full GPU suite on HEAD + C3 bench at full size + smoke
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT; TAG=r02c
echo "== smoke"; timeout 600 python __graft_entry__.py --smoke > $OUT/${TAG}_smoke.log 2>&1; echo "rc=$?"; tail -4 $OUT/${TAG}_smoke.log
echo "== pytest gpu"; timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider > $OUT/${TAG}_pytest.log 2>&1; echo "rc=$?"; tail -8 $OUT/${TAG}_pytest.log
echo "== bench c3"; timeout 600 python bench.py --workload c3 --no-extra --no-cpu --steps 3 --warmup 3 > $OUT/${TAG}_bench_c3.json 2> $OUT/${TAG}_bench_c3.err; tail -2 $OUT/${TAG}_bench_c3.err
python -c "import sys,json; d=json.loads(open('$OUT/${TAG}_bench_c3.json').read().strip().splitlines()[-1]); print('c3', d['value'], d['kernel_ms_per_step'], d['fp64_peak_tflops'])"
nvidia-smi topo -m > $OUT/${TAG}_topo.txt 2>&1; lscpu > $OUT/${TAG}_lscpu.txt 2>&1; numactl -H >> $OUT/${TAG}_lscpu.txt 2>&1
echo "== done"
