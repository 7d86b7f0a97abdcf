#!/bin/bash
# multi-GPU trip (gpurun --gpus N): the sharded-epoch parity test on 2 GPUs, then the strong-scaling bench line at N ranks
set -u
N=${1:-2}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
if [ "$N" -le 2 ]; then
  echo "== multirank parity test"; timeout 900 python -m pytest tests/test_gpu_multirank.py -x -q -m gpu 2>&1 | tail -8
fi
echo "== bench strong N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_strong_n$N.json 2> gpurun_out/bench_strong_n$N.err
tail -3 gpurun_out/bench_strong_n$N.err
if [ "${2:-}" = "weak" ]; then
  echo "== bench weak N=$N"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --scaling weak > gpurun_out/bench_weak_n$N.json 2> gpurun_out/bench_weak_n$N.err
fi
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_*_n*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["scaling"], "N", d["n_gpus"], "value %.3g"%d["value"], "step %.2f sync %.2f e2e %.2f"%(d["ms_per_step"], d["ms_per_step_unpipelined"], d["e2e"]["ms_per_step"]), d["stage_ms"], "head p50 %.1f p99 %.1f"%(d["get_head_p50_us"], d["get_head_p99_us"]), "depth", d["config"]["pipeline_depth"])
    except Exception as e:
        print(f, "failed", e)
PY
