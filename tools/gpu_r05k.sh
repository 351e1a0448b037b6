#!/bin/bash
# four GPUs: C4 (1e5 k-points, N = 512, two-stage reduction + fused peer-store gather) under torchrun
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT; TAG=${TAG:-r05k}; N=${NGPU:-4}
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload c4 --no-extra --no-cpu --no-peaks --steps 2 --warmup 3 > $OUT/${TAG}_bench_c4_n$N.json 2> $OUT/${TAG}_bench_c4_n$N.err; echo "rc=$?"
python - <<PY
import json
try:
    d=json.loads([l for l in open("$OUT/${TAG}_bench_c4_n$N.json") if l.startswith("{")][-1])
    print({k:d[k] for k in ("value","n_gpus","ms_per_step","gpu_launches")}, "e2e", d["e2e"]["value"], d["kernel_ms_per_step"], str(d.get("extra"))[:500])
except Exception as e: print("parse fail", e)
PY
tail -2 $OUT/${TAG}_bench_c4_n$N.err
