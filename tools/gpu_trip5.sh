#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
echo "== decompress microbench (v2 byte loads / v3 branch-free 128-bit loads)"
for b in v2 v3 v2 v3; do timeout 120 tools/decompress_bench_$b.bin | tee -a gpurun_out/decompress_bench2.jsonl; done
echo "== head clocks"; timeout 300 python tools/head_clocks.py | tee -a gpurun_out/head_clocks2.jsonl
echo "== fork choice + spec + epoch tests"; timeout 1200 python -m pytest tests/test_gpu_forkchoice.py tests/test_gpu_spec.py tests/test_gpu_epoch.py tests/test_gpu_fullsize.py tests/test_gpu_bls.py -x -q -m gpu 2>&1 | tail -6
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-configs > gpurun_out/bench5.json 2> gpurun_out/bench5.err; tail -3 gpurun_out/bench5.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench5.json").read().strip().splitlines()[-1]); print("step %.2f sync %.2f e2e %.2f"%(d["ms_per_step"], d["ms_per_step_unpipelined"], d["e2e"]["ms_per_step"]), d["stage_ms"], "head p50 %.1f p99 %.1f"%(d["get_head_p50_us"],d["get_head_p99_us"]))
PY
