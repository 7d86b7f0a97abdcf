#!/bin/bash
# One GPU trip: gather/guard tests first (under a short timeout: a hung mbarrier must not hang the box), then the whole -m gpu
# suite, the bench line, and the ncu launch list of the bench command.  Everything lands in gpurun_out/.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== gather tests"; timeout 300 python -m pytest tests/test_gpu_gather.py -x -q -m gpu 2>&1 | tail -15
rc=$?
if [ $rc -ne 0 ]; then echo "gather tests failed/hung (rc=$rc): continuing with B2_K2_TMA=0"; export B2_K2_TMA=0; fi
echo "== full gpu suite"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 6000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
