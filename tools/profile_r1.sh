#!/bin/bash
# Run under gpurun (1 GPU).  Launch list + full captures; exports CSV pages on the box and keeps gpurun_out small (<64 MiB).
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1b.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_g2_decompress -s 1 -c 1 -o /tmp/prof_g2_decompress_r1b -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ncu -i /tmp/prof_g2_decompress_r1b.ncu-rep --page raw --csv > gpurun_out/g2_decompress_raw_r1b.csv
ncu -i /tmp/prof_g2_decompress_r1b.ncu-rep --page details --csv > gpurun_out/g2_decompress_details_r1b.csv
ncu -i /tmp/prof_g2_decompress_r1b.ncu-rep --page source --csv > /tmp/src.csv 2>/dev/null; head -c 20000000 /tmp/src.csv > gpurun_out/g2_decompress_source_r1b.csv
cp /tmp/prof_g2_decompress_r1b.ncu-rep gpurun_out/ 2>/dev/null
ncu --set full --clock-control none -k regex:"k_miller_team|k_final_team|k_hash_to_g2|k_g2_finish|k_g1_aggregate|k_ghost|k_g2_segment|k_lmd" -s 24 -c 12 -o /tmp/prof_verify_r1b -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1
ncu -i /tmp/prof_verify_r1b.ncu-rep --page raw --csv > gpurun_out/verify_raw_r1b.csv
ncu -i /tmp/prof_verify_r1b.ncu-rep --page details --csv > gpurun_out/verify_details_r1b.csv
du -sh gpurun_out; ls -la gpurun_out
