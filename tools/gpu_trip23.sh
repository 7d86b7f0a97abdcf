#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== head clocks"; timeout 200 python tools/head_clocks.py | tee gpurun_out/head_clocks8.jsonl
echo "== fork choice / spec tests"; timeout 600 python -m pytest tests/test_gpu_forkchoice.py tests/test_gpu_spec.py tests/test_gpu_epoch.py -x -q -m gpu 2>&1 | tail -4
echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
