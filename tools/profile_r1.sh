#!/bin/bash
# Run under gpurun (1 GPU).  Produces the launch list and one full capture of the dominant kernel in gpurun_out/.
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_g2_decompress -s 1 -c 1 -o gpurun_out/prof_g2_decompress_r1 -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_miller|k_final_verdict|k_hash_to_g2|k_sig_prepare|k_g1_aggregate|k_ghost" -c 8 -o gpurun_out/prof_verify_r1 -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1
ls -la gpurun_out
