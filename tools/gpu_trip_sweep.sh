#!/bin/bash
# strong-scaling sweep at N ranks: tail form x pipeline depth (bench lines only)
set -u
N=${1:-2}; shift
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
port=29520
for cfg in "$@"; do
  form=${cfg%%:*}; depth=${cfg##*:}
  port=$((port+1))
  echo "== N=$N tail=$form depth=$depth"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 20 --warmup 5 \
      --tail-form $form --depth $depth --no-extra-configs --no-cpu-baseline > gpurun_out/sweep_n${N}_${form}_d${depth}.json 2> gpurun_out/sweep_n${N}_${form}_d${depth}.err
  python - "$N" "$form" "$depth" <<'PY'
import json,sys
n,form,depth=sys.argv[1:4]
f="gpurun_out/sweep_n%s_%s_d%s.json"%(n,form,depth)
try:
    d=json.loads(open(f).read().strip().splitlines()[-1]); print("   value %.3g step %.2f sync %.2f e2e %.2f head p50 %.1f (nccl %.1f)"%(d["value"], d["ms_per_step"], d["ms_per_step_unpipelined"], d["e2e"]["ms_per_step"], d["get_head_p50_us"], d.get("get_head_nccl_path_p50_us",0)))
except Exception as e:
    print("   failed", e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
done
