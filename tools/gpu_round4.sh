#!/bin/bash
# A/B: L2-aware tile order + evict_first activation loads in the backward launches; evict_first saves in the forward chain
mkdir -p gpurun_out
for v in NEFES_L2_ORDER=0 NEFES_L2_ORDER=1 NEFES_L2_ORDER=0 NEFES_L2_ORDER=1; do
  echo "== $v" | tee -a gpurun_out/r4_l2order.log
  env $v timeout 300 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r4_tmp.json 2>/dev/null
  python - <<'PY' | tee -a gpurun_out/r4_l2order.log
import json
d=json.loads(open('gpurun_out/r4_tmp.json').read().strip().splitlines()[-1])
print('ms_per_step', round(d['ms_per_step'],4), 'loss', d['final_loss'])
for k,v in d['kernels'].items():
    if 'bwd' in k or 'wgrad' in k or 'fwd_fine' in k: print('  ', k, round(v['ms_per_step'],4), round(v['GB_per_s']))
PY
done
for v in NEFES_TS2_STG=0 NEFES_TS2_STG=2 NEFES_TS2_STG=0 NEFES_TS2_STG=2; do
  echo "== $v" | tee -a gpurun_out/r4_evict.log
  env $v timeout 300 python tools/prof_fwd.py 2>&1 | grep "saves=on.*chain_fwd" | tee -a gpurun_out/r4_evict.log
done
timeout 900 python -m pytest tests -m gpu -x -q -k "bf16 or bench_shape or one_call or tiles" 2>&1 | tail -3 | tee gpurun_out/r4_tests.log
