timeout 200 python -m pytest tests -m gpu -q 2>&1 | tail -1
for k in 128 32 64; do
B2_K2_BLOCK=$k timeout 120 python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('k2_block=$k', 'pipelined', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'unpipelined', round(d['ms_per_step_unpipelined'],2), 'head_p50', round(d['get_head_p50_us'],1))
"
done
