#!/bin/bash
# Per-phase cycle counts of the blocked tridiagonalisation (debug build with -DTBK_PANEL_TIMING, rebuilt on the box only).
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-pt}
cp tbmodels_b200/libtbk.so /tmp/libtbk.keep
TBK_BUILD_DEFINES=-DTBK_PANEL_TIMING python -m tbmodels_b200.build --force > /dev/null 2>&1
for spec in "c4 148" "c5 592"; do
  set -- $spec
  timeout 600 python bench.py --workload $1 --nk $2 --no-extra --no-cpu --no-peaks --steps 1 --warmup 1 2>/dev/null | grep "panel timing" | tail -18 | sort > $OUT/${TAG}_$1.txt
  grep phases $OUT/${TAG}_$1.txt; grep "warp  0\|warp 15\|warp  7" $OUT/${TAG}_$1.txt
done
cp /tmp/libtbk.keep tbmodels_b200/libtbk.so
for pfd in 0 1 2 4 8; do
  TBK_PANEL_PFD=$pfd timeout 600 python bench.py --workload c4 --nk 1184 --no-extra --no-cpu --no-peaks --steps 2 --warmup 2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pfd $pfd', d['kernel_ms_per_step'])"
done
