#!/bin/bash
# TS2 forward chain: stamps and timing ablations (results change under NEFES_CHAIN_X; timing only)
mkdir -p gpurun_out
NEFES_CHAIN_DBG=1 timeout 300 python tools/prof_fwd.py > gpurun_out/ts2_stamps.log 2>&1
for x in 0 1 4 5 16 21 64; do
  echo "== NEFES_CHAIN_X=$x" >> gpurun_out/ts2_ablate.log
  NEFES_UNSAFE_EXPERIMENTS=1 NEFES_CHAIN_X=$x timeout 300 python tools/prof_fwd.py 2>&1 | grep chain_fwd >> gpurun_out/ts2_ablate.log
done
echo "== SS chain (NEFES_FWD_SS=1)" >> gpurun_out/ts2_ablate.log
NEFES_FWD_SS=1 timeout 300 python tools/prof_fwd.py 2>&1 | grep chain_fwd >> gpurun_out/ts2_ablate.log
cat gpurun_out/ts2_ablate.log
