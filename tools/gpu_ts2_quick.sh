#!/bin/bash
# quick A/B of the TS2 forward chain: env assignments in $1 (e.g. "NEFES_TS2_TURNS=0"), timing + stamps + parity subset
mkdir -p gpurun_out
export NEFES_X=1
for v in "$@"; do
  echo "== $v" | tee -a gpurun_out/ts2_quick.log
  env $v timeout 300 python tools/prof_fwd.py 2>&1 | grep "saves=on.*chain_fwd" | tee -a gpurun_out/ts2_quick.log
done
env $1 NEFES_CHAIN_DBG=1 timeout 300 python tools/prof_fwd.py 2>&1 | grep -A29 "chain_ts dbg" | head -30 > gpurun_out/ts2_quick_stamps.log
env $1 timeout 600 python -m pytest tests -m gpu -x -q -k "bf16 or bench_shape or one_call" 2>&1 | tail -3 | tee -a gpurun_out/ts2_quick.log
