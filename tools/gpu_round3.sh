#!/bin/bash
# round-2 late check: full gpu test-suite after the TS v1 removal, default bench, HBM write ceiling, evict_first saves,
# and the per-GPU share of an 8-GPU strong-scaling step on one GPU (launch list by kernel)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r3_tests.log
tools/hbm_write 2>&1 | tee gpurun_out/r3_hbm_write.txt
for v in NEFES_TS2_STG=0 NEFES_TS2_STG=2; do
  echo "== $v" | tee -a gpurun_out/r3_evict.log
  env $v timeout 300 python tools/prof_fwd.py 2>&1 | grep "chain_fwd" | tee -a gpurun_out/r3_evict.log
done
timeout 600 python bench.py > gpurun_out/r3_bench_train.json 2> gpurun_out/r3_bench_train.err
tail -c 600 gpurun_out/r3_bench_train.json
timeout 300 python bench.py --as-world 8 --no-extras --no-cpu-baseline > gpurun_out/r3_bench_asworld8.json 2> gpurun_out/r3_bench_asworld8.err
head -c 400 gpurun_out/r3_bench_asworld8.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r3_launches_asworld8.csv \
  python bench.py --as-world 8 --no-extras --no-cpu-baseline --no-graph --steps 2 --warmup 3 > gpurun_out/r3_ncu_asworld8.log 2>&1
echo done
