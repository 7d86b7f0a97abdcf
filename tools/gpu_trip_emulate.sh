#!/bin/bash
# 1 GPU: one rank's share of a sharded epoch (bench.py --emulate-world W), sweep of depth / tail form / drain hint / reserved SMs
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
for cfg in "$@"; do
  IFS=: read -r W form depth tl env <<< "$cfg"
  echo "== emulate W=$W tail=$form depth=$depth team_last=$tl ${env:-}"
  out=gpurun_out/emu_w${W}_${form}_d${depth}_t${tl}_${env:-none}.json
  env ${env:-X=1} timeout 300 python bench.py --emulate-world $W --tail-form $form --depth $depth --team-last $tl --steps 20 --warmup 5 --no-extra-configs --no-cpu-baseline > $out 2> ${out%.json}.err
  python - "$out" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open(f).read().strip().splitlines()[-1]); print("   step %.2f sync %.2f e2e %.2f  agg_share %.2f fav_share %.2f"%(d["ms_per_step"], d["ms_per_step_unpipelined"], d["e2e"]["ms_per_step"], d["stage_ms"]["bls_aggregate_rank_share"], d["stage_ms"]["fast_aggregate_verify_rank_share"]))
except Exception as e:
    print("   failed", e); print(open(f.replace(".json",".err")).read()[-1200:])
PY
done
