#!/bin/bash
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-m}; N=${2:-2}
nvidia-smi --query-gpu=index,name --format=csv | head -9
echo "== pytest multi" ; timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py::test_tridiag_variants_agree -m gpu -q -p no:cacheprovider 2>&1 | tail -5
for w in c2 c3; do
echo "== bench $w N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $w --no-extra > $OUT/${TAG}_bench_${w}_n$N.json 2> $OUT/${TAG}_bench_${w}_n$N.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("$OUT/${TAG}_bench_${w}_n$N.json") if l.startswith("{")][-1])
    print({k:d[k] for k in ("value","n_gpus","ms_per_step","gpu_launches")}, "e2e", d["e2e"]["value"], d["kernel_ms_per_step"], d["extra"], d["clocks"])
except Exception as e: print("parse fail", e)
PY
tail -3 $OUT/${TAG}_bench_${w}_n$N.err
done
echo "== reference arm under torchrun"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --impl reference --steps 2 --warmup 1 2>/dev/null | grep '^{' | cut -c1-300
echo "== done"
