#!/bin/bash
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-c45}
for spec in "c5 16384" "c4 2048"; do
  set -- $spec
  echo "== bench $1 nk=$2"
  timeout 1200 python bench.py --workload $1 --nk $2 --no-extra --steps 2 --warmup 3 > $OUT/${TAG}_bench_$1.json 2> $OUT/${TAG}_bench_$1.err ; echo "rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("$OUT/${TAG}_bench_$1.json"))
    print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, "e2e", d["e2e"]["value"], "roof", {k:d["roofline"][k] for k in ("kernel","achieved","peak","frac","kernel_share_of_step")}, d["kernel_ms_per_step"], "cpu", d["cpu_baseline"]["value"] if d["cpu_baseline"] else None)
except Exception as e: print("parse fail", e)
PY
  tail -3 $OUT/${TAG}_bench_$1.err
done
echo "== hamilton path timing"
timeout 600 python - <<'PY'
import torch, time, numpy as np
import tbmodels_b200 as tbk
from oracle import workloads as wl
for name,p,nk in (("haldane",wl.haldane(),20_000_000),("c3",wl.synthetic(36,250),131072)):
    ev=tbk.Evaluator(p,device=0); ev.profile(True)
    k=torch.rand((nk,p.dim),dtype=torch.float64,device="cuda")
    out=torch.empty((nk,p.size,p.size),dtype=torch.complex128,device="cuda")
    for conv in (2,1):
        for _ in range(2): ev.hamilton_device(k,convention=conv,out=out)
        ev.profile_read()
        torch.cuda.synchronize(); t0=time.perf_counter()
        for _ in range(3): ev.hamilton_device(k,convention=conv,out=out)
        torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/3
        pr=ev.profile_read()
        print(name,"conv",conv,"k/s %.3e"%(nk/dt),"out GB/s %.0f"%(nk*p.size**2*16/dt/1e9),{c:round(v[0]/3,3) for c,v in pr.items() if v[1]})
PY
echo "== done"
