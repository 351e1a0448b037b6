#!/bin/bash
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT; TAG=${TAG:-r03x}
PYTHONPATH=. timeout 600 python tools/eigh_bench.py 8:262144 16:131072 24:131072 32:65536 36:65536 48:32768 64:16384 82:8192 96:2048 128:1024 > $OUT/${TAG}_eigh.log 2>&1; cat $OUT/${TAG}_eigh.log
timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:"eigh_kernel" -s 2 -c 1 -o $OUT/${TAG}_eigh36 env PYTHONPATH=. python tools/eigh_bench.py 36:14208 > $OUT/${TAG}_ncu.log 2>&1; tail -2 $OUT/${TAG}_ncu.log
