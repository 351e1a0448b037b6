#!/bin/bash
# One GPU session: smoke, parity tests, sanitizer, bench, ncu launch list + full captures.  Logs -> gpurun_out/.
# Every step has its own timeout so a hung kernel cannot eat the whole lease.
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out
mkdir -p $OUT
TAG=${1:-r1}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/${TAG}_gpu.txt 2>&1
echo "== build+smoke" ; timeout 600 python __graft_entry__.py --smoke > $OUT/${TAG}_smoke.log 2>&1 ; echo "smoke rc=$?"
tail -5 $OUT/${TAG}_smoke.log
echo "== peaks" ; timeout 120 python -c "import tbmodels_b200 as t; print(t.fp64_peaks(4000))" > $OUT/${TAG}_peaks.log 2>&1 ; cat $OUT/${TAG}_peaks.log
echo "== pytest gpu" ; timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -rA -p no:cacheprovider > $OUT/${TAG}_pytest_gpu.log 2>&1 ; echo "pytest rc=$?"
grep -E "passed|failed|error" $OUT/${TAG}_pytest_gpu.log | tail -3
if [ "${SKIP_SANITIZER:-0}" != "1" ]; then
echo "== sanitizer" ; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_smoke.py > $OUT/${TAG}_memcheck.log 2>&1 ; echo "memcheck rc=$?"
tail -4 $OUT/${TAG}_memcheck.log
fi
echo "== bench" ; timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err ; echo "bench rc=$?"
tail -c 3000 $OUT/${TAG}_bench.json ; tail -5 $OUT/${TAG}_bench.err
for w in c3 c1 ${EXTRA_WORKLOADS}; do
  echo "== bench $w" ; timeout 900 python bench.py --workload $w --no-extra > $OUT/${TAG}_bench_$w.json 2> $OUT/${TAG}_bench_$w.err ; echo "bench $w rc=$?"
  tail -c 2500 $OUT/${TAG}_bench_$w.json ; tail -3 $OUT/${TAG}_bench_$w.err
done
if [ "${SKIP_NCU:-0}" != "1" ]; then
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_c2.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu --no-peaks --nk 20000000 > $OUT/${TAG}_ncu_c2.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_c3.csv \
   python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu --no-peaks --no-extra --nk 262144 > $OUT/${TAG}_ncu_c3.log 2>&1
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hk_small -s 3 -c 1 -f -o $OUT/${TAG}_prof_hk_small \
   python bench.py --steps 2 --warmup 3 --no-cpu --no-peaks --no-extra --nk 20000000 > $OUT/${TAG}_ncu_full_c2.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"hk_gemm|tridiag|ql_" -s 9 -c 3 -f -o $OUT/${TAG}_prof_c3 \
   python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu --no-peaks --no-extra --nk 131072 > $OUT/${TAG}_ncu_full_c3.log 2>&1
ls -la $OUT
fi
echo "== done"
