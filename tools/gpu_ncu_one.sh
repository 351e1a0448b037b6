#!/bin/bash
# One ncu --set full capture of one kernel of one workload: tools/gpu_ncu_one.sh TAG WORKLOAD NK KERNEL_REGEX [SKIP] [ENV...]
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT
TAG=$1; WL=$2; NK=$3; KRE=$4; SKIP=${5:-1}; shift 5
env "$@" timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$KRE" -s $SKIP -c 1 -f -o $OUT/${TAG} \
   python bench.py --workload $WL --steps 1 --warmup 1 --no-cpu --no-peaks --no-extra --nk $NK > $OUT/${TAG}.log 2>&1
tail -3 $OUT/${TAG}.log
ls -la $OUT/${TAG}.ncu-rep
