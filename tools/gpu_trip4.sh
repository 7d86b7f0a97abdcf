#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
echo "== decompress microbench (old headers / new local / new smem / new local + byte loads)"
for b in old v0 v1 v2; do timeout 120 tools/decompress_bench_$b.bin | tee -a gpurun_out/decompress_bench.jsonl; done
echo "== head clocks"; timeout 300 python tools/head_clocks.py | tee -a gpurun_out/head_clocks.jsonl; B2_HEAD_FUSED=0 timeout 300 python tools/head_clocks.py | tee -a gpurun_out/head_clocks.jsonl
echo "== full gpu suite"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
for v in "B2_POW_SMEM=1" "B2_HEAD_FUSED=0"; do
  echo "== bench $v"; env $v timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-configs > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, "step %.2f sync %.2f e2e %.2f"%(d["ms_per_step"], d["ms_per_step_unpipelined"], d["e2e"]["ms_per_step"]), d["stage_ms"], "head p50 %.1f p99 %.1f"%(d["get_head_p50_us"],d["get_head_p99_us"]), "gather tma %.3f ldg %.3f"%(d["roofline"]["gather"]["frac"], d["roofline"]["gather"]["plain_ldg_form"]["frac"]))
    except Exception as e:
        print(f, "failed", e)
PY
