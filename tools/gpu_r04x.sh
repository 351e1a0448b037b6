#!/bin/bash
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT; TAG=${TAG:-r04x}
PYTHONPATH=. timeout 900 python tools/tridiag_sweep.py --variants two_chase1,two 13:7 17:64 26:63 40:64 100:64 > $OUT/${TAG}_small.log 2>&1; cat $OUT/${TAG}_small.log
PYTHONPATH=. timeout 1500 python tools/tridiag_sweep.py --variants chase1,default,one 256:296 512:296 224:18944 256:18944 320:9472 512:9472 700:2368 > $OUT/${TAG}_sweep.log 2>&1; cat $OUT/${TAG}_sweep.log
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "twostage or size_boundaries" > $OUT/${TAG}_pytest.log 2>&1; echo "rc=$?"; tail -3 $OUT/${TAG}_pytest.log
