#!/bin/bash
# multi-GPU bench lines for every workload at N = $1 GPUs -> gpurun_out/scale_n$1_<workload>.json
N=$1
mkdir -p gpurun_out
run() {  # name, extra args
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --no-cpu-baseline $2 \
    > gpurun_out/scale_n${N}_$1.json 2> gpurun_out/scale_n${N}_$1.err
  echo "== $1 rc=$?"; head -c 260 gpurun_out/scale_n${N}_$1.json; echo; grep -a "Error\|error" gpurun_out/scale_n${N}_$1.err | head -3
}
run train_weak "--steps 20 --warmup 3"
run train_strong "--steps 20 --warmup 3 --scaling strong --no-extras"
run c3s3 "--steps 20 --warmup 3 --workload c3s3"
run refine "--workload refine --steps 4"
run sweep "--workload sweep --steps 5"
