#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
echo "== decompress microbench: window-4 schedule (v2) vs schedule with the a^255 run token (v4)"
for b in v2 v4 v2 v4; do timeout 120 tools/decompress_bench_$b.bin | tee -a gpurun_out/decompress_bench3.jsonl; done
echo "== single-GPU suite"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench16.json 2> gpurun_out/bench16.err; tail -3 gpurun_out/bench16.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench16.json").read().strip().splitlines()[-1]); print("step %.2f sync %.2f e2e %.2f"%(d["ms_per_step"], d["ms_per_step_unpipelined"], d["e2e"]["ms_per_step"]), d["stage_ms"], "head p50 %.1f p99 %.1f"%(d["get_head_p50_us"],d["get_head_p99_us"]))
PY
