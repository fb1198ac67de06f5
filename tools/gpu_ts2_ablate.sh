#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/ts2_ablate.log
export NEFES_X=1
for x in 0 1 5 21; do
  echo "== NEFES_CHAIN_X=$x" >> gpurun_out/ts2_ablate.log
  NEFES_UNSAFE_EXPERIMENTS=1 NEFES_CHAIN_X=$x timeout 300 python tools/prof_fwd.py 2>&1 | grep "saves=on.*chain_fwd" >> gpurun_out/ts2_ablate.log
done
cat gpurun_out/ts2_ablate.log
