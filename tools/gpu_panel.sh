#!/bin/bash
# Blocked tridiagonalisation: parity at every size that reaches it + C4 / C5 timing against the unblocked kernels.
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-p}
echo "== pytest (default dispatch)"
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider -k "size_boundaries or supercell or oversize or c5_small or synthetic_golden or batch_invariance" > $OUT/${TAG}_pytest_a.log 2>&1; echo "rc=$?"; tail -4 $OUT/${TAG}_pytest_a.log
echo "== pytest (panel forced from N=9)"
TBK_TRIDIAG_PANEL_MIN=9 timeout 900 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider -k "size_boundaries or synthetic_golden or c5_small or batch_invariance" > $OUT/${TAG}_pytest_b.log 2>&1; echo "rc=$?"; tail -4 $OUT/${TAG}_pytest_b.log
run() {  # name workload nk env...
  local name=$1 wlk=$2 nk=$3; shift 3
  env "$@" timeout 900 python bench.py --workload $wlk --nk $nk --no-extra --no-cpu --no-peaks --steps 2 --warmup 3 > $OUT/${TAG}_bench_$name.json 2> $OUT/${TAG}_bench_$name.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/${TAG}_bench_$name.json"))
    print("$name", "value %.4g"%d["value"], "ms %.1f"%d["ms_per_step"], d["kernel_ms_per_step"])
except Exception as e: print("$name parse fail", e)
PY
}
run c4_panel c4 2048 X=1
run c5_lpr16 c5 16384 X=1
run c5_lpr16_minb4 c5 16384 TBK_PANEL_MINB4=1
run c5_lpr16_t128 c5 16384 TBK_PANEL_T=128
run c5_lpr32 c5 16384 TBK_PANEL_LPR=32
PYTHONPATH=. python tools/tridiag_sweep.py 120:8192 164:4096 200:2048 256:2048 2>&1 | tail -4
TBK_PANEL_LPR=16 PYTHONPATH=. python tools/tridiag_sweep.py 200:2048 256:2048 2>&1 | tail -2
echo "== bench c1 large batch"
timeout 600 python bench.py --workload c1 --nk 2000000 --no-extra --no-cpu --no-peaks --steps 3 --warmup 2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c1 2e6', d['value'], d['kernel_ms_per_step'])"
echo "== done"
