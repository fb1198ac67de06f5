#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chain_fwd_ts2 -s 4 -c 1 -f -o gpurun_out/ts2 python tools/prof_fwd.py > gpurun_out/ncu_ts2.log 2>&1
tail -3 gpurun_out/ncu_ts2.log; ls -la gpurun_out/*.ncu-rep
