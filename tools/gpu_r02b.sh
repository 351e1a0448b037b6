#!/bin/bash
# round 2, second GPU session: merged-reduction register kernel; QL profile at N = 36
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT; TAG=r02b
echo "== reg parity (plain)"; timeout 600 python tools/sanitize_reg.py > $OUT/${TAG}_reg_plain.log 2>&1; echo "rc=$?"; tail -2 $OUT/${TAG}_reg_plain.log
echo "== racecheck"; timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_reg.py > $OUT/${TAG}_reg_racecheck.log 2>&1; echo "rc=$?"; tail -2 $OUT/${TAG}_reg_racecheck.log
echo "== pytest subset"; timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider -k "size_boundaries or synthetic_golden or batch_invariance or tridiag_variants or staged or c3_properties or degenerate" > $OUT/${TAG}_pytest.log 2>&1; echo "rc=$?"; tail -3 $OUT/${TAG}_pytest.log
echo "== sweep"; PYTHONPATH=. timeout 900 python tools/tridiag_sweep.py --variants default,bw8,noreg 24:131072 28:131072 32:131072 36:131072 40:65536 48:65536 64:32768 > $OUT/${TAG}_sweep.log 2>&1; cat $OUT/${TAG}_sweep.log | cut -c1-260
echo "== bench c3"; timeout 600 python bench.py --workload c3 --nk 1048576 --no-extra --no-cpu --no-peaks --steps 3 --warmup 3 > $OUT/${TAG}_bench_c3.json 2> $OUT/${TAG}_bench_c3.err; tail -2 $OUT/${TAG}_bench_c3.err
python -c "import sys,json; d=json.loads(open('$OUT/${TAG}_bench_c3.json').read().strip().splitlines()[-1]); print('c3', d['value'], d['kernel_ms_per_step'])"
echo "== ncu reg + ql"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tridiag_reg|ql_smem" -s 4 -c 2 -f -o $OUT/${TAG}_prof_c3 python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu --no-peaks --no-extra --nk 131072 > $OUT/${TAG}_ncu.log 2>&1; tail -2 $OUT/${TAG}_ncu.log; ls -la $OUT/${TAG}*.ncu-rep
echo "== done"
