#!/bin/bash
# ncu of the final two-stage kernels: pipelined bulge chasing at N = 512 (one resident wave), both stages at N = 128 (C5)
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT; TAG=${TAG:-r05c}
PYTHONPATH=. timeout 600 ncu --set full --clock-control none --import-source on -k regex:band_chase_pipe -c 1 -f -o $OUT/${TAG}_chase512 python tools/tridiag_sweep.py --variants default 512:1776 > $OUT/${TAG}_ncu1.log 2>&1; echo "rc=$?"
PYTHONPATH=. timeout 600 ncu --set full --clock-control none --import-source on -k regex:band_reduce -c 1 -f -o $OUT/${TAG}_reduce128 python tools/tridiag_sweep.py --variants default 128:16384 > $OUT/${TAG}_ncu2.log 2>&1; echo "rc=$?"
PYTHONPATH=. timeout 600 ncu --set full --clock-control none --import-source on -k regex:band_chase_pipe -c 1 -f -o $OUT/${TAG}_chase128 python tools/tridiag_sweep.py --variants default 128:16384 > $OUT/${TAG}_ncu3.log 2>&1; echo "rc=$?"
grep -h "nk=" $OUT/${TAG}_ncu?.log
