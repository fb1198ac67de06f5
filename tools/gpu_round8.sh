#!/bin/bash
# A/B: weight-gradient flush of trunk_bwd as bulk reductions (cp.reduce.async.bulk) instead of per-thread red.v4
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "bf16 or bench_shape or one_call or tiles" 2>&1 | tail -3 | tee gpurun_out/r8_tests.log
for v in NEFES_BULK_FLUSH=0 NEFES_BULK_FLUSH=1 NEFES_BULK_FLUSH=0 NEFES_BULK_FLUSH=1; do
  for w in 1 8; do
    echo "== $v as-world $w" | tee -a gpurun_out/r8_flush.log
    env $v timeout 300 python bench.py --no-extras --no-cpu-baseline --as-world $w 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms_per_step', round(d['ms_per_step'],4), 'loss', d['final_loss'])
for k,v in d['kernels'].items():
    if 'bwd' in k or 'wgrad' in k: print('  ', k, round(v['ms_per_step'],4), round(v['GB_per_s']), round(v['TFLOP_per_s']))" | tee -a gpurun_out/r8_flush.log
  done
done
