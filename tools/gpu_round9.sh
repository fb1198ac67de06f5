#!/bin/bash
# trunk_bwd as ONE launch walking the four layer pairs (NEFES_TRUNK_PASSES=4, default) against four launches (=1)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "bf16 or bench_shape or one_call or tiles" 2>&1 | tail -3 | tee gpurun_out/r9_tests.log
grep -q passed gpurun_out/r9_tests.log || exit 1
for v in NEFES_TRUNK_PASSES=1 NEFES_TRUNK_PASSES=4 NEFES_TRUNK_PASSES=2 NEFES_TRUNK_PASSES=1 NEFES_TRUNK_PASSES=4; do
  for w in 1 8; do
    echo "== $v as-world $w" | tee -a gpurun_out/r9_passes.log
    env $v timeout 300 python bench.py --no-extras --no-cpu-baseline --as-world $w 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms_per_step', round(d['ms_per_step'],4), 'loss', d['final_loss'], 'trunk_bwd', round(d['kernels']['trunk_bwd']['ms_per_step'],4), round(d['kernels']['trunk_bwd']['GB_per_s']))" | tee -a gpurun_out/r9_passes.log
  done
done
