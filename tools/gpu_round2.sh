#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; python -c "
import json; d=json.load(open('gpurun_out/bench_train.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['traffic'], d['extras'].get('train_step_tf32_field'), d['extras'].get('train_step_fp32_field'), d['refine_iters_per_s'], d['cpu_baseline']['value'])"
timeout 600 python bench.py --workload c3s3 --no-cpu-baseline > gpurun_out/bench_c3s3.json 2> gpurun_out/bench_c3s3.err; head -c 200 gpurun_out/bench_c3s3.json; echo
timeout 600 python bench.py --workload sweep --steps 5 > gpurun_out/bench_sweep.json 2> gpurun_out/bench_sweep.err; head -c 200 gpurun_out/bench_sweep.json; echo
