#!/bin/bash
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT; TAG=${TAG:-r04u}
timeout 1200 python bench.py --workload c4 --no-extra --steps 3 --warmup 3 > $OUT/${TAG}_bench_c4.json 2> $OUT/${TAG}_bench_c4.err; echo "rc=$?"; python - <<PY
import json
d=json.loads(open("$OUT/${TAG}_bench_c4.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","kernel_ms_per_step","roofline","gpu_launches")}); print(d.get("e2e")); print(d.get("cpu_baseline",{}).get("value"))
PY
