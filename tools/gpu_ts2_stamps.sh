#!/bin/bash
mkdir -p gpurun_out
export NEFES_X=1
env "$@" NEFES_CHAIN_DBG=1 timeout 300 python tools/prof_fwd.py 2>&1 | grep -A150 "chain_ts dbg" | head -150 > gpurun_out/ts2_stamps2.log
env "$@" timeout 300 python tools/prof_fwd.py 2>&1 | grep "saves=on.*chain_fwd"
