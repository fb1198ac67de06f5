#!/bin/bash
# bench.py in every workload at N=1 + the reference arm + new tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_c3.py -x -q > gpurun_out/pytest_c3.log 2>&1; tail -3 gpurun_out/pytest_c3.log
timeout 900 python bench.py > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; tail -c 600 gpurun_out/bench_train.json; tail -3 gpurun_out/bench_train.err
timeout 600 python bench.py --workload c3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; head -c 400 gpurun_out/bench_c3.json; tail -3 gpurun_out/bench_c3.err
timeout 600 python bench.py --workload refine --steps 4 --no-cpu-baseline > gpurun_out/bench_refine.json 2> gpurun_out/bench_refine.err; head -c 400 gpurun_out/bench_refine.json; tail -3 gpurun_out/bench_refine.err
timeout 600 python bench.py --workload sweep --steps 5 > gpurun_out/bench_sweep.json 2> gpurun_out/bench_sweep.err; head -c 300 gpurun_out/bench_sweep.json; tail -3 gpurun_out/bench_sweep.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; head -c 300 gpurun_out/bench_reference.json; tail -3 gpurun_out/bench_reference.err
