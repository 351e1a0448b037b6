#!/bin/bash
# Regular-mesh entry point: parity tests + C3 / C5 k-grid timing through eigenval_mesh vs the explicit-k path.
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-mesh}
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider -k "mesh" > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/${TAG}_pytest.log
run() {  # name workload nk flags...
  local name=$1 wlk=$2 nk=$3; shift 3
  timeout 900 python bench.py --workload $wlk --nk $nk --no-extra --no-cpu --no-peaks --steps 3 --warmup 3 "$@" > $OUT/${TAG}_bench_$name.json 2> $OUT/${TAG}_bench_$name.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$OUT/${TAG}_bench_$name.json") if l.startswith("{")][-1])
    print("$name", "value %.4g"%d["value"], "ms %.1f"%d["ms_per_step"], d["kernel_ms_per_step"], d["config"].get("mesh"))
except Exception as e:
    print("$name parse fail", e); print(open("$OUT/${TAG}_bench_$name.err").read()[-1500:])
PY
}
run c3_mesh c3 2097152 --mesh
run c3_expl c3 2097152
run c5_mesh c5 16384 --mesh
run c5_expl c5 16384
echo "== done"
