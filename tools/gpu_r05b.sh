#!/bin/bash
# validation of the final state: smoke, all GPU tests, the default bench line (with its extras)
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT; TAG=${TAG:-r05b}
echo "== smoke"; timeout 600 python __graft_entry__.py --smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/${TAG}_smoke.log
echo "== pytest gpu"; S0=$SECONDS; timeout 2400 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$? in $((SECONDS-S0)) s"
grep -E "^(FAILED|ERROR)|passed|failed" $OUT/${TAG}_pytest_gpu.log | tail -8
echo "== bench (default)"; S0=$SECONDS; timeout 1500 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$? in $((SECONDS-S0)) s"; tail -2 $OUT/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open("$OUT/${TAG}_bench.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("metric","value","ms_per_step","gpu_launches")}); print("e2e", d.get("e2e",{}).get("value")); print("roofline", {k:d["roofline"].get(k) for k in ("kernel","frac","achieved","peak")})
for k,v in (d.get("extra") or {}).items():
    if isinstance(v, dict): print(k, v.get("value"), v.get("kernel_ms_per_step"))
PY
