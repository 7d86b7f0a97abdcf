# diagnostic sweep of the launch-shape knobs of libb200pos.so (B2_TAIL_BLOCK, B2_DEC_BLOCK, B2_PAIRING_FORM) and the pipeline depth; run on the B200
for cfg in "128 128 auto 3" "32 128 auto 3" "64 128 auto 3" "128 128 auto 4" "128 128 auto 3"; do
  set -- $cfg
  B2_TAIL_BLOCK=$1 B2_DEC_BLOCK=$2 B2_PAIRING_FORM=$3 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --depth $4 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
st = d['stage_ms']
print('tail=$1 dec=$2 form=$3 depth=$4', 'pipelined', round(d['ms_per_step'],2), 'unpipelined', round(d['ms_per_step_unpipelined'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'agg', round(st['bls_aggregate_2^20_sigs'],2), 'fav', round(st['fast_aggregate_verify_2048'],2))
"
done
