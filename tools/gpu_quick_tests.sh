#!/bin/bash
# run a pytest -k selection on the GPU box:  K="expr" TAG=name bash tools/gpu_quick_tests.sh
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT; TAG=${TAG:-quick}
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider -x ${EXTRA} ${K:+-k "$K"} > $OUT/${TAG}_pytest.log 2>&1; echo "rc=$?"; tail -40 $OUT/${TAG}_pytest.log
