#!/bin/bash
# round 2: new bench contract (C3 256^3 default, strong scaling, gather inside) at N=1; full GPU suite under filterwarnings=error
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT; TAG=${TAG:-r02d}
echo "skip pytest"
echo "== bench default"; SECONDS=0; timeout 1500 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "rc=$?"; echo "elapsed ${SECONDS}s"; tail -3 $OUT/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open('$OUT/${TAG}_bench.json').read().strip().splitlines()[-1])
print('main', d['value'], d['ms_per_step'], d['kernel_ms_per_step'], 'e2e', d['e2e']['value'])
print('roofline', {k:d['roofline'][k] for k in ('kernel','achieved','peak','frac')})
print('cpu', d['cpu_baseline'] and (d['cpu_baseline']['value'], d['cpu_baseline']['kind']))
for k,v in d['extra'].items():
    if 'value' in v: print(k, v['value'], v['ms_per_step'], v.get('kernel_ms_per_step'), 'e2e', v['e2e']['value'], 'cpu', v.get('cpu_baseline',{}).get('value'))
    else: print(k, {p:(q['value'], q['kernel_ms_per_step']) for p,q in v['points'].items()})
PY
echo skip ref
echo "== done"
