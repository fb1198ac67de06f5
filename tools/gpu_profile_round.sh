#!/bin/bash
# Round evidence: launch list of the bench step and one ncu --set full capture of the hot MLP kernels.
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline --no-graph > gpurun_out/launches_bench.log 2>&1
tail -1 gpurun_out/launches.csv | cut -c1-200
# forward coarse+fine chains, one trunk launch, both fused head launches, the leftover wgrad: from a whole bench step
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"chain_fwd_ts2|trunk_bwd|fused_bwd|wgrad_kernel" -s 40 -c 14 -f -o gpurun_out/step_full \
  python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline --no-graph > gpurun_out/step_full.log 2>&1
ls -la gpurun_out/step_full.ncu-rep; tail -2 gpurun_out/step_full.log | cut -c1-200
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
