#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_refine_c4.py -q -s -k fusion > gpurun_out/pytest_fusion_refine.log 2>&1; grep -a "fusion, graph\|passed\|failed\|Error" gpurun_out/pytest_fusion_refine.log | head
timeout 600 python bench.py --workload c3s3 --no-cpu-baseline > gpurun_out/bench_c3s3.json 2> gpurun_out/bench_c3s3.err; head -c 600 gpurun_out/bench_c3s3.json; echo; tail -3 gpurun_out/bench_c3s3.err
