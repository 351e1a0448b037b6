#!/bin/bash
# Short GPU session: parity tests + benches (+ optional ncu full capture of the C3 kernels).
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-q}
timeout 300 python -c "import tbmodels_b200 as t; print(t.fp64_peaks(4000))" > $OUT/${TAG}_peaks.log 2>&1 ; cat $OUT/${TAG}_peaks.log
echo "== pytest gpu" ; timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -rA -p no:cacheprovider > $OUT/${TAG}_pytest_gpu.log 2>&1 ; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" $OUT/${TAG}_pytest_gpu.log | tail -12
for w in c2 c3 ${EXTRA_WORKLOADS}; do
  echo "== bench $w" ; timeout 900 python bench.py --workload $w --no-extra ${BENCH_FLAGS} > $OUT/${TAG}_bench_$w.json 2> $OUT/${TAG}_bench_$w.err ; echo "bench $w rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("$OUT/${TAG}_bench_$w.json"))
    print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, "e2e", d["e2e"]["value"], "roof", {k:d["roofline"][k] for k in ("kernel","achieved","peak","frac","kernel_share_of_step")}, d["kernel_ms_per_step"], d["clocks"])
except Exception as e: print("parse fail", e)
PY
  tail -3 $OUT/${TAG}_bench_$w.err
done
if [ "${RUN_NCU:-0}" = "1" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hk_small -s 3 -c 1 -f -o $OUT/${TAG}_prof_hk_small \
   python bench.py --steps 2 --warmup 3 --no-cpu --no-peaks --no-extra --nk 20000000 > $OUT/${TAG}_ncu_full_c2.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"hk_gemm|hk_phase|tridiag|ql_" -s 12 -c 4 -f -o $OUT/${TAG}_prof_c3 \
   python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu --no-peaks --no-extra --nk 131072 > $OUT/${TAG}_ncu_full_c3.log 2>&1
tail -3 $OUT/${TAG}_ncu_full_c3.log
fi
echo "== done"
