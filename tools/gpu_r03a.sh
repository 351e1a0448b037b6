#!/bin/bash
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT; TAG=${TAG:-r03a}
NG=${NG:-2}
echo "== pytest multi"; timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -p no:cacheprovider -rs -x > $OUT/${TAG}_pytest_multi.log 2>&1; echo "rc=$?"; tail -3 $OUT/${TAG}_pytest_multi.log
echo "== bench N=$NG"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --no-extra > $OUT/${TAG}_bench_n$NG.json 2> $OUT/${TAG}_bench_n$NG.err; echo "rc=$?"; tail -3 $OUT/${TAG}_bench_n$NG.err
python - <<PY
import json
d=json.loads([l for l in open('$OUT/${TAG}_bench_n$NG.json').read().strip().splitlines() if l.startswith('{')][-1])
print('main', d['n_gpus'], d['value'], d['ms_per_step'], 'nccl variant', d['nccl_variant_ms_per_step'], 'gather', d['allgather_ms'], d['gather_fallback_reason'], 'checked', d['gather_rows_checked_bit_exact'], 'e2e', d['e2e']['value'])
print(d['kernel_ms_per_step'])
PY
echo "== done"
