#!/bin/bash
# Run under gpurun (1 GPU).  Launch list + full captures; exports CSV pages on the box and keeps gpurun_out small (<64 MiB).
# Numbers printed by bench.py under ncu are never bench values.
set -x
T=${1:-r1c}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$T.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_g2_decompress -s 1 -c 1 -o /tmp/prof_g2_decompress_$T -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ncu -i /tmp/prof_g2_decompress_$T.ncu-rep --page raw --csv > gpurun_out/g2_decompress_raw_$T.csv
ncu -i /tmp/prof_g2_decompress_$T.ncu-rep --page details --csv > gpurun_out/g2_decompress_details_$T.csv
ncu --set full --clock-control none -k regex:"k_miller|k_final|k_hash_to_g2|k_g2_finish|k_g1_aggregate|k_ghost|k_g2_segment|k_lmd" -s 26 -c 14 -o /tmp/prof_verify_$T -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1
ncu -i /tmp/prof_verify_$T.ncu-rep --page raw --csv > gpurun_out/verify_raw_$T.csv
ncu -i /tmp/prof_verify_$T.ncu-rep --page details --csv > gpurun_out/verify_details_$T.csv
du -sh gpurun_out; ls -la gpurun_out | head -30
