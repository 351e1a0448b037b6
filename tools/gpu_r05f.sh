#!/bin/bash
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT; TAG=${TAG:-r05f}
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "twostage or size_boundaries or one_stage or c5 or smoke" > $OUT/${TAG}_pytest.log 2>&1; echo "rc=$?"; tail -3 $OUT/${TAG}_pytest.log
timeout 900 python bench.py --workload c5 --no-extra --no-cpu > $OUT/${TAG}_bench_c5.json 2> $OUT/${TAG}_bench_c5.err; echo "rc=$?"
python - <<PY
import json
d=json.loads(open("$OUT/${TAG}_bench_c5.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","kernel_ms_per_step","gpu_launches")}); print(d.get("e2e",{}).get("value")); print(d["roofline"])
PY
