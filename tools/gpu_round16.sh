timeout 600 python -m pytest tests -m gpu -x -q -k "bf16 or bench_shape or one_call or tiles" 2>&1 | tail -3
for i in 1 2; do timeout 300 python bench.py --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms_per_step', round(d['ms_per_step'],4), 'loss', d['final_loss'])
for k,v in d['kernels'].items():
    if 'heads' in k or 'trunk' in k: print('  ', k, v['launches_per_step'], round(v['ms_per_step'],4))"; done
