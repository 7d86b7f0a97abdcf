#!/bin/bash
# final-build ncu evidence: launch list of the bench command + full capture of the fused get_head, K2 and the two-lane hash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra-configs > gpurun_out/bench_under_ncu_final.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_get_head_fused|k_g1_gather_tma|k_hash_to_g2|k_g2_decompress" -c 8 -o /tmp/prof_final -f python tools/profile_small.py > gpurun_out/ncu_final.log 2>&1
ncu -i /tmp/prof_final.ncu-rep --page raw --csv > gpurun_out/r2_final_raw.csv
ncu -i /tmp/prof_final.ncu-rep --page details --csv > gpurun_out/r2_final_details.csv
ls -la gpurun_out | tail -8
