#!/bin/bash
# Per-phase cycle counts of the band reduction (debug build with -DTBK_BAND_TIMING, rebuilt on the box only).
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT; TAG=${TAG:-bt}
cp tbmodels_b200/libtbk.so /tmp/libtbk.keep
TBK_BUILD_DEFINES=-DTBK_BAND_TIMING python -m tbmodels_b200.build --force > $OUT/${TAG}_build.log 2>&1
for spec in ${SIZES:-512:148 256:296}; do
  PYTHONPATH=. timeout 600 python tools/tridiag_sweep.py --variants two_s1 $spec 2>&1 | grep "band timing" | tail -1
done | tee $OUT/${TAG}_timing.txt
cp /tmp/libtbk.keep tbmodels_b200/libtbk.so
