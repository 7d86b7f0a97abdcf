#!/bin/bash
# trip 2: gather microbenchmark, RLC tests, bench A/B per-aggregate vs RLC
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
echo "== gather microbench"; timeout 120 tools/gather_bench.bin 2>&1 | tee gpurun_out/gather_bench.jsonl
echo "== rlc tests"; timeout 900 python -m pytest tests/test_gpu_bls.py tests/test_gpu_epoch.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -15
echo "== bench rlc"; timeout 600 python bench.py --steps 20 --warmup 5 --rlc --no-cpu-baseline --no-extra-configs > gpurun_out/bench_rlc.json 2> gpurun_out/bench_rlc.err; python - <<'PY'
import json
for f in ("gpurun_out/bench_rlc.json",):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["ms_per_step"], d["ms_per_step_unpipelined"], d["e2e"]["ms_per_step"], d["stage_ms"])
    except Exception as e:
        print(f, "failed", e); print(open(f.replace('.json','.err')).read()[-2000:])
PY
