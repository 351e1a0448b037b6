#!/bin/bash
# C5 strong scaling (BASELINE.json config 5): a fixed 2^17-point k-set split over N GPUs; run once per N under gpurun --gpus N.
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT; N=${1:-1}; TOTAL=${2:-131072}
NK=$((TOTAL / N))
for mode in "" "--mesh"; do
  tag=c5_strong_n${N}${mode:+_mesh}
  if [ "$N" = "1" ]; then
    timeout 900 python bench.py --workload c5 --nk $NK --no-extra --no-cpu --no-peaks --steps 2 --warmup 3 $mode > $OUT/$tag.json 2> $OUT/$tag.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --workload c5 --nk $NK --no-extra --no-cpu --no-peaks --steps 2 --warmup 3 $mode > $OUT/$tag.json 2> $OUT/$tag.err
  fi
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$OUT/$tag.json") if l.startswith("{")][-1])
    print("$tag", "n_gpus", d["n_gpus"], "k/s %.4g"%d["value"], "ms/step %.1f"%d["ms_per_step"], d["kernel_ms_per_step"], d["extra"])
except Exception as e:
    print("$tag parse fail", e); print(open("$OUT/$tag.err").read()[-800:])
PY
done
