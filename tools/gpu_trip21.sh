#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== fork choice tests"; timeout 900 python -m pytest tests/test_gpu_forkchoice.py tests/test_gpu_spec.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -4
echo "== head clocks"; timeout 300 python tools/head_clocks.py | tee gpurun_out/head_clocks6.jsonl
