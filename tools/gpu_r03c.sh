#!/bin/bash
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT; TAG=${TAG:-r03c}
timeout 600 ncu --metrics gpu__time_duration.sum,launch__grid_size,launch__block_size,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"tridiag|ql_|hk_gemm_kernel" -s 21 -c 12 --csv --log-file $OUT/${TAG}_c5_launches.csv python bench.py --workload c5 --nk 14208 --steps 1 --warmup 3 --no-cpu --no-peaks --no-extra > $OUT/${TAG}_c5.log 2>&1; tail -1 $OUT/${TAG}_c5.log | cut -c1-200
python - <<PY
import csv
rows=list(csv.reader(open('$OUT/${TAG}_c5_launches.csv')))
hdr=None; cur={}
out=[]
for r in rows:
    if hdr is None:
        if 'Kernel Name' in r: hdr=r
        continue
    if len(r)!=len(hdr): continue
    d=dict(zip(hdr,r))
    key=(d['ID'], d['Kernel Name'].split('(')[0][-50:])
    cur.setdefault(key,{})[d['Metric Name']]=d['Metric Value']+' '+d['Metric Unit']
for k,v in cur.items(): print(k[0], k[1], {a.split('.')[0].replace('sm__','').replace('smsp__','').replace('launch__',''):b for a,b in v.items()})
PY
