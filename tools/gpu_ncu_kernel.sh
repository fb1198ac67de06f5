#!/bin/bash
# ncu --set full of one launch of kernel regex $1 (skip $2 launches) during tools/time_mlp.py; report -> gpurun_out/$3.ncu-rep
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$1 -s ${2:-2} -c 1 -f -o gpurun_out/$3 python tools/time_mlp.py bf16 > gpurun_out/ncu_$3.log 2>&1
tail -2 gpurun_out/ncu_$3.log
