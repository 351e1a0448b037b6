#!/bin/bash
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT; TAG=${TAG:-r02x}
echo "== plain parity"; timeout 600 python tools/sanitize_reg.py > $OUT/${TAG}_reg_plain.log 2>&1; echo "rc=$?"; tail -2 $OUT/${TAG}_reg_plain.log
echo "== sweep"; PYTHONPATH=. timeout 900 python tools/tridiag_sweep.py --variants default,nola,bw8 24:131072 28:131072 32:131072 36:131072 40:65536 48:65536 > $OUT/${TAG}_sweep.log 2>&1; cut -c1-330 $OUT/${TAG}_sweep.log
echo "== pytest subset"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "register_tridiag or size_boundaries or synthetic_golden or batch_invariance or tridiag_variants or staged or c3_properties or degenerate" > $OUT/${TAG}_pytest.log 2>&1; echo "rc=$?"; tail -5 $OUT/${TAG}_pytest.log
echo "== racecheck"; timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_reg.py > $OUT/${TAG}_reg_racecheck.log 2>&1; echo "rc=$?"; tail -2 $OUT/${TAG}_reg_racecheck.log
echo "== bench c3"; timeout 600 python bench.py --workload c3 --nk 2097152 --no-extra --no-cpu --steps 3 --warmup 3 > $OUT/${TAG}_bench_c3.json 2> $OUT/${TAG}_bench_c3.err; tail -2 $OUT/${TAG}_bench_c3.err
python -c "import sys,json; d=json.loads(open('$OUT/${TAG}_bench_c3.json').read().strip().splitlines()[-1]); print('c3', d['value'], d['kernel_ms_per_step'])"
echo "== done"
