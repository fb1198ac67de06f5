#!/bin/bash
# ncu --set full of one launch of kernel regex $1 (skipping $2 launches) while running "$4..." -> gpurun_out/$3.ncu-rep
mkdir -p gpurun_out
K=$1; S=$2; O=$3; shift 3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c 1 -f -o gpurun_out/$O "$@" > gpurun_out/ncu_$O.log 2>&1
tail -2 gpurun_out/ncu_$O.log
