#!/bin/bash
# trip 3: full suite on the new build, bench + A/B of the two decompression changes, ncu evidence
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
echo "== full gpu suite"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
for v in "B2_POW_SMEM=0" "B2_SEGSUM_TAIL=0" "B2_K2_TMA=0"; do
  echo "== bench $v"; env $v timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-configs > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, "step %.2f sync %.2f e2e %.2f"%(d["ms_per_step"], d["ms_per_step_unpipelined"], d["e2e"]["ms_per_step"]), d["stage_ms"], "head p50 %.1f"%d["get_head_p50_us"], "gather tma %.3f ldg %.3f"%(d["roofline"]["gather"]["frac"], d["roofline"]["gather"]["plain_ldg_form"]["frac"]))
    except Exception as e:
        print(f, "failed", e)
PY
echo "== ncu launch list"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra-configs > gpurun_out/bench_under_ncu.log 2>&1
echo "== ncu full: decompress"
ncu --set full --clock-control none --import-source on -k regex:k_g2_decompress -s 2 -c 1 -o /tmp/prof_dec -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra-configs > gpurun_out/ncu_full1.log 2>&1
ncu -i /tmp/prof_dec.ncu-rep --page raw --csv > gpurun_out/r2_g2_decompress_raw.csv
ncu -i /tmp/prof_dec.ncu-rep --page details --csv > gpurun_out/r2_g2_decompress_details.csv
echo "== ncu full: gather + fork choice"
ncu --set full --clock-control none --import-source on -k regex:"k_g1_gather|k_ghost" -c 12 -o /tmp/prof_gfc -f python tools/profile_small.py > gpurun_out/ncu_full2.log 2>&1
ncu -i /tmp/prof_gfc.ncu-rep --page raw --csv > gpurun_out/r2_gather_forkchoice_raw.csv
ncu -i /tmp/prof_gfc.ncu-rep --page details --csv > gpurun_out/r2_gather_forkchoice_details.csv
ncu -i /tmp/prof_gfc.ncu-rep --page source --csv -k regex:k_ghost > gpurun_out/r2_ghost_source.csv 2>/dev/null
ls -la gpurun_out | head -40; du -sh gpurun_out
