#!/bin/bash
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT; TAG=${TAG:-r02u}
echo "== pytest subset"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "background_ql or batch_invariance or host_pipeline or mesh or c3_properties or device_pointer" > $OUT/${TAG}_pytest.log 2>&1; echo "rc=$?"; tail -15 $OUT/${TAG}_pytest.log
for ov in 1 0; do
echo "== bench c3 overlap=$ov"; TBK_QL_OVERLAP=$ov timeout 600 python bench.py --workload c3 --nk 2097152 --no-extra --no-cpu --steps 3 --warmup 3 > $OUT/${TAG}_bench_c3_ov$ov.json 2> $OUT/${TAG}_bench_c3_ov$ov.err; tail -2 $OUT/${TAG}_bench_c3_ov$ov.err
python -c "import sys,json; d=json.loads(open('$OUT/${TAG}_bench_c3_ov$ov.json').read().strip().splitlines()[-1]); print('c3', d['value'], d['ms_per_step'], d['kernel_ms_per_step'], 'e2e', d['e2e']['value'])"
echo "== bench c3 mesh overlap=$ov"; TBK_QL_OVERLAP=$ov timeout 600 python bench.py --workload c3 --nk 2097152 --mesh --no-extra --no-cpu --steps 3 --warmup 3 > $OUT/${TAG}_bench_c3m_ov$ov.json 2> $OUT/${TAG}_bench_c3m_ov$ov.err; tail -2 $OUT/${TAG}_bench_c3m_ov$ov.err
python -c "import sys,json; d=json.loads(open('$OUT/${TAG}_bench_c3m_ov$ov.json').read().strip().splitlines()[-1]); print('c3 mesh', d['value'], d['ms_per_step'], d['kernel_ms_per_step'])"
echo "== bench c5 overlap=$ov"; TBK_QL_OVERLAP=$ov timeout 600 python bench.py --workload c5 --nk 65536 --no-extra --no-cpu --steps 2 --warmup 3 > $OUT/${TAG}_bench_c5_ov$ov.json 2> $OUT/${TAG}_bench_c5_ov$ov.err; tail -2 $OUT/${TAG}_bench_c5_ov$ov.err
python -c "import sys,json; d=json.loads(open('$OUT/${TAG}_bench_c5_ov$ov.json').read().strip().splitlines()[-1]); print('c5', d['value'], d['ms_per_step'], d['kernel_ms_per_step'])"
done
echo "== done"
