#!/bin/bash
set +e
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out; mkdir -p $OUT; TAG=${TAG:-r05l}
timeout 300 python __graft_entry__.py --smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/${TAG}_smoke.log
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "not reference_suite and not multi" > $OUT/${TAG}_pytest.log 2>&1; echo "rc=$?"; tail -2 $OUT/${TAG}_pytest.log
