#!/bin/bash
# C2 (2-band nearest-cell model): parity of the small-N tests + bench with and without the trigonometric-product kernel.
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-c2}
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider -k "haldane or simple or edge or regression or general_path or batch_invariance or device_pointer or pipeline or mutation or patch" > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/${TAG}_pytest.log
for v in basis kp2 nobasis; do
  unset TBK_NO_BASIS TBK_BASIS_KP
  if [ $v = nobasis ]; then export TBK_NO_BASIS=1; fi
  if [ $v = kp2 ]; then export TBK_BASIS_KP=2; fi
  timeout 600 python bench.py --workload c2 --no-extra --no-cpu --steps 10 --warmup 3 > $OUT/${TAG}_bench_$v.json 2> $OUT/${TAG}_bench_$v.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/${TAG}_bench_$v.json"))
    print("$v", "value %.4g"%d["value"], "ms %.4f"%d["ms_per_step"], "e2e %.3g"%d["e2e"]["value"], d["roofline"]["frac"], d["kernel_ms_per_step"])
except Exception as e: print("$v parse fail", e)
PY
done
unset TBK_NO_BASIS TBK_BASIS_KP
if [ "${RUN_NCU:-0}" = "1" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"hk_basis|hk_small" -s 3 -c 1 -f -o $OUT/${TAG}_prof_c2 \
   python bench.py --steps 2 --warmup 3 --no-cpu --no-peaks --no-extra --nk 20000000 > $OUT/${TAG}_ncu_c2.log 2>&1
tail -2 $OUT/${TAG}_ncu_c2.log
fi
echo "== done"
