#!/bin/bash
# Full GPU session: smoke, parity tests, sanitizer, benches (all workloads), ncu launch lists + full captures.
# Logs -> gpurun_out/.  Every step has its own timeout so a hung kernel cannot eat the whole lease.
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out
mkdir -p $OUT
TAG=${1:-r1}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/${TAG}_gpu.txt 2>&1
echo "== build+smoke" ; timeout 600 python __graft_entry__.py --smoke > $OUT/${TAG}_smoke.log 2>&1 ; echo "smoke rc=$?"
tail -3 $OUT/${TAG}_smoke.log
echo "== peaks" ; timeout 120 python -c "import tbmodels_b200 as t; print(t.fp64_peaks(4000))" > $OUT/${TAG}_peaks.log 2>&1 ; cat $OUT/${TAG}_peaks.log
echo "== pytest gpu" ; timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -rA -p no:cacheprovider > $OUT/${TAG}_pytest_gpu.log 2>&1 ; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" $OUT/${TAG}_pytest_gpu.log | tail -5
if [ "${SKIP_SANITIZER:-0}" != "1" ]; then
echo "== sanitizer" ; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_smoke.py > $OUT/${TAG}_memcheck.log 2>&1 ; echo "memcheck rc=$?"
tail -2 $OUT/${TAG}_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_smoke.py > $OUT/${TAG}_racecheck.log 2>&1 ; echo "racecheck rc=$?"
tail -2 $OUT/${TAG}_racecheck.log
fi
echo "== bench (default)" ; timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err ; echo "bench rc=$?"
tail -3 $OUT/${TAG}_bench.err
echo "== bench reference arm" ; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err ; echo "rc=$?"
for spec in "c1 8000" "c3 2097152" "c5 16384" "c4 2048"; do
  set -- $spec
  echo "== bench $1" ; timeout 900 python bench.py --workload $1 --nk $2 --no-extra > $OUT/${TAG}_bench_$1.json 2> $OUT/${TAG}_bench_$1.err ; echo "rc=$?"
  tail -2 $OUT/${TAG}_bench_$1.err
done
python - <<PY
import json,glob
for f in sorted(glob.glob("$OUT/${TAG}_bench*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        r=d.get("roofline") or {}
        print(f.split("/")[-1], "value=%.4g"%d["value"], "e2e=%.4g"%d["e2e"]["value"], "ms=%.4g"%d["ms_per_step"], r.get("kernel"), "frac=%s"%r.get("frac"), d.get("kernel_ms_per_step"), "cpu=%s"%((d.get("cpu_baseline") or {}).get("value")))
    except Exception as e: print(f, "parse fail", e)
PY
if [ "${SKIP_NCU:-0}" != "1" ]; then
echo "== ncu launch lists"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_c2.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu --no-peaks --nk 20000000 > $OUT/${TAG}_ncu_c2.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_c3.csv \
   python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu --no-peaks --no-extra --nk 262144 > $OUT/${TAG}_ncu_c3.log 2>&1
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"hk_small|hk_basis" -s 3 -c 1 -f -o $OUT/${TAG}_prof_hk_small \
   python bench.py --steps 2 --warmup 3 --no-cpu --no-peaks --no-extra --nk 20000000 > $OUT/${TAG}_ncu_full_c2.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"hk_gemm|hk_phase|tridiag|ql_" -s 12 -c 4 -f -o $OUT/${TAG}_prof_c3 \
   python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu --no-peaks --no-extra --nk 131072 > $OUT/${TAG}_ncu_full_c3.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"hk_gemm|tridiag|ql_" -s 9 -c 3 -f -o $OUT/${TAG}_prof_c5 \
   python bench.py --workload c5 --steps 1 --warmup 3 --no-cpu --no-peaks --no-extra --nk 4096 > $OUT/${TAG}_ncu_full_c5.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"tridiag|bisect" -s 6 -c 2 -f -o $OUT/${TAG}_prof_c4 \
   python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu --no-peaks --no-extra --nk 296 > $OUT/${TAG}_ncu_full_c4.log 2>&1
ls -la $OUT | grep ncu-rep
fi
echo "== done"
