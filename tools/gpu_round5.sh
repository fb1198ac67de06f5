#!/bin/bash
# A/B of the forward chain's L2 policies: 2 = evict_first saves, 3 = + weights evict_last / encodings evict_first, 4 = + raw evict_first
mkdir -p gpurun_out
for v in NEFES_TS2_STG=2 NEFES_TS2_STG=3 NEFES_TS2_STG=4 NEFES_TS2_STG=0 NEFES_TS2_STG=2 NEFES_TS2_STG=3 NEFES_TS2_STG=4; do
  echo "== $v" | tee -a gpurun_out/r5_evict.log
  env $v timeout 300 python tools/prof_fwd.py 2>&1 | grep "chain_fwd" | grep -v "coarse saves=off\|fine   saves=off" | tee -a gpurun_out/r5_evict.log
done
for v in NEFES_TS2_STG=2 NEFES_TS2_STG=4; do
  echo "== $v" | tee -a gpurun_out/r5_step.log
  env $v timeout 300 python bench.py --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms_per_step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'loss', d['final_loss'], 'fwd_fine', round(d['kernels']['chain_fwd_fine']['ms_per_step'],4), 'composite_fwd_fine', round(d['kernels']['composite_fwd_fine']['ms_per_step'],4))" | tee -a gpurun_out/r5_step.log
done
timeout 900 python -m pytest tests -m gpu -x -q -k "bf16 or bench_shape or one_call or tiles" 2>&1 | tail -3 | tee gpurun_out/r5_tests.log
