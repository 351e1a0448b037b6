#!/bin/bash
# 8 GPUs: default bench (strong scaling of C3 256^3, fused gather inside the timed region) + multi-GPU pytest
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT; TAG=${TAG:-r02z}
NG=${NG:-8}
nvidia-smi topo -m > $OUT/${TAG}_topo.txt 2>&1
echo "== bench N=$NG"; SECONDS=0; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG > $OUT/${TAG}_bench_n$NG.json 2> $OUT/${TAG}_bench_n$NG.err; echo "rc=$? elapsed ${SECONDS}s"; tail -5 $OUT/${TAG}_bench_n$NG.err
python - <<PY
import json
d=json.loads([l for l in open('$OUT/${TAG}_bench_n$NG.json').read().strip().splitlines() if l.startswith('{')][-1])
print('main', d['n_gpus'], d['value'], d['ms_per_step'], 'nccl variant', d['nccl_variant_ms_per_step'], 'gather', d['allgather_ms'], d['gather'], d['gather_fallback_reason'], 'checked', d['gather_rows_checked_bit_exact'], 'e2e', d['e2e']['value'])
print(d['kernel_ms_per_step'], d['clocks'])
for k,v in d['extra'].items():
    if 'value' in v: print(k, v['value'], v['ms_per_step'], 'nccl', v.get('nccl_variant_ms_per_step'), 'gather', v['allgather_ms'])
    else: print(k, {p:(q['value'], q['ms_per_step'], q['nccl_variant_ms_per_step']) for p,q in v['points'].items()})
PY
echo "== done"
