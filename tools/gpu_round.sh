#!/bin/bash
# One GPU-box visit: gpu tests, MLP timing (default and the TS2 forward chain), bench.  Output under gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python tools/time_mlp.py bf16 > gpurun_out/time_mlp_default.log 2>&1; cat gpurun_out/time_mlp_default.log | tail -3
NEFES_FWD_SS=1 timeout 300 python tools/time_mlp.py bf16 > gpurun_out/time_mlp_ss.log 2>&1; tail -3 gpurun_out/time_mlp_ss.log
NEFES_FWD_SS=1 timeout 600 python -m pytest tests -m gpu -x -q -k "bf16 or bench_shape or one_call" > gpurun_out/pytest_gpu_ss.log 2>&1; tail -3 gpurun_out/pytest_gpu_ss.log
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 3000 gpurun_out/bench_default.json
