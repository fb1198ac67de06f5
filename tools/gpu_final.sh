#!/bin/bash
# state-of-the-round check: full gpu test-suite, smoke(), every bench workload at N=1 with its CPU baseline, the reference arm,
# and the launch list of one training step
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/f_tests.log
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -9 | tee gpurun_out/f_smoke.log
timeout 900 python bench.py > gpurun_out/f_bench_train.json 2> gpurun_out/f_bench_train.err; tail -c 300 gpurun_out/f_bench_train.json; tail -2 gpurun_out/f_bench_train.err
timeout 600 python bench.py --workload c3 > gpurun_out/f_bench_c3.json 2> gpurun_out/f_bench_c3.err; head -c 300 gpurun_out/f_bench_c3.json; tail -2 gpurun_out/f_bench_c3.err
timeout 600 python bench.py --workload c3s3 > gpurun_out/f_bench_c3s3.json 2> gpurun_out/f_bench_c3s3.err; head -c 300 gpurun_out/f_bench_c3s3.json; tail -2 gpurun_out/f_bench_c3s3.err
timeout 600 python bench.py --workload refine --steps 4 > gpurun_out/f_bench_refine.json 2> gpurun_out/f_bench_refine.err; head -c 300 gpurun_out/f_bench_refine.json; tail -2 gpurun_out/f_bench_refine.err
timeout 600 python bench.py --workload sweep --steps 5 > gpurun_out/f_bench_sweep.json 2> gpurun_out/f_bench_sweep.err; head -c 300 gpurun_out/f_bench_sweep.json; tail -2 gpurun_out/f_bench_sweep.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/f_bench_reference.json 2> gpurun_out/f_bench_reference.err; head -c 300 gpurun_out/f_bench_reference.json; tail -2 gpurun_out/f_bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/f_launches.csv python bench.py --no-extras --no-cpu-baseline --no-graph --steps 2 --warmup 3 > gpurun_out/f_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/f_launches.csv | head -30
