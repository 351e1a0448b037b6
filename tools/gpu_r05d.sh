#!/bin/bash
# two GPUs: sharded / fused-gather tests incl. the two-stage size, C5 bench under torchrun (the bench asserts that the
# fused gather equals the NCCL gather and re-evaluates sampled rows of every shard)
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT; TAG=${TAG:-r05d}; N=2
nvidia-smi --query-gpu=index,name --format=csv | head -4
echo "== pytest multi"; timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -p no:cacheprovider > $OUT/${TAG}_pytest_multi_2gpu.log 2>&1; echo "rc=$?"; tail -4 $OUT/${TAG}_pytest_multi_2gpu.log
for w in c5 c4; do
echo "== bench $w N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $w --no-extra --no-cpu --steps 2 --warmup 3 > $OUT/${TAG}_bench_${w}_n$N.json 2> $OUT/${TAG}_bench_${w}_n$N.err; echo "rc=$?"
python - <<PY
import json
try:
    d=json.loads([l for l in open("$OUT/${TAG}_bench_${w}_n$N.json") if l.startswith("{")][-1])
    print({k:d[k] for k in ("value","n_gpus","ms_per_step","gpu_launches")}, "e2e", d["e2e"]["value"], d["kernel_ms_per_step"], str(d.get("extra"))[:600])
except Exception as e: print("parse fail", e)
PY
tail -2 $OUT/${TAG}_bench_${w}_n$N.err
done
