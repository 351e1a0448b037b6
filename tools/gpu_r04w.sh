#!/bin/bash
# final one-stage / two-stage table + ncu of the bulge-chasing kernel at a full wave
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT; TAG=${TAG:-r04w}
PYTHONPATH=. timeout 1500 python tools/tridiag_sweep.py --variants one,two 161:18944 200:18944 224:18944 256:18944 288:9472 320:9472 384:9472 448:9472 512:9472 640:4736 700:2368 816:2368 > $OUT/${TAG}_sweep.log 2>&1
echo "rc=$?"; cat $OUT/${TAG}_sweep.log
PYTHONPATH=. timeout 900 python tools/tridiag_sweep.py --variants one,two 256:296 512:296 > $OUT/${TAG}_sweep_small.log 2>&1
cat $OUT/${TAG}_sweep_small.log
PYTHONPATH=. timeout 1200 ncu --set full --clock-control none --import-source on -k regex:band_chase -c 1 -f -o $OUT/${TAG}_chase python tools/tridiag_sweep.py --variants default 512:9189 > $OUT/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 $OUT/${TAG}_ncu.log
