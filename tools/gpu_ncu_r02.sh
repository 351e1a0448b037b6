#!/bin/bash
# round-2 ncu captures: --set full of the dominant kernels of every workload + the launch list of the default bench command
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT; TAG=${TAG:-r02n}
NCU="ncu --set full --clock-control none --import-source on -f"
B="python bench.py --steps 1 --warmup 3 --no-cpu --no-peaks --no-extra"
echo "== c3"; timeout 900 $NCU -k regex:"hk_gemm_kernel|tridiag_reg|ql_smem" -s 12 -c 4 -o $OUT/${TAG}_c3 $B --workload c3 --nk 113664 > $OUT/${TAG}_c3.log 2>&1; tail -1 $OUT/${TAG}_c3.log
echo "== c2"; timeout 900 $NCU -k regex:"hk_basis" -s 3 -c 1 -o $OUT/${TAG}_c2 $B --workload c2 > $OUT/${TAG}_c2.log 2>&1; tail -1 $OUT/${TAG}_c2.log
echo "== c5"; timeout 900 $NCU -k regex:"hk_gemm_kernel|tridiag_panel|ql_smem" -s 9 -c 3 -o $OUT/${TAG}_c5 $B --workload c5 --nk 16384 > $OUT/${TAG}_c5.log 2>&1; tail -1 $OUT/${TAG}_c5.log
echo "== c4"; timeout 1200 $NCU -k regex:"hk_gemm_kernel|tridiag_panel|bisect" -s 9 -c 3 -o $OUT/${TAG}_c4 $B --workload c4 --nk 888 > $OUT/${TAG}_c4.log 2>&1; tail -1 $OUT/${TAG}_c4.log
echo "== c3 mesh"; timeout 900 $NCU -k regex:"mesh_lines|hk_gemm_kernel" -s 6 -c 2 -o $OUT/${TAG}_c3mesh $B --workload c3 --nk 113664 --mesh > $OUT/${TAG}_c3mesh.log 2>&1; tail -1 $OUT/${TAG}_c3mesh.log
echo "== launch list of the default bench command (explicit C3, 2 steps)"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $OUT/${TAG}_launches_c3_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extra --nk 2097152 > $OUT/${TAG}_launches.log 2>&1; tail -1 $OUT/${TAG}_launches.log | cut -c1-300
ls -la $OUT/${TAG}*
echo "== done"
