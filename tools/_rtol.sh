for r in 20 200 1000 10000; do
python bench.py --no-configs --picard-rtol-e15 $r --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('rtol_e15', $r, round(d['value'],1), 'it/s', round(d['ms_per_step'],4), 'ms  e2e', round(d['e2e']['value'],1), 'rounds', d['roofline'].get('picard_iterations_last_sweep'), 'parity', d['parity']['per_iteration'])
"
done
