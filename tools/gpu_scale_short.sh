#!/bin/bash
# the three multi-GPU lines that changed late in round 2 (train weak / strong, refinement) at N = $1 GPUs -> gpurun_out/scale2_n$1_<workload>.json
N=$1
mkdir -p gpurun_out
run() {  # name, extra args
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --no-cpu-baseline --no-extras $2 \
    > gpurun_out/scale2_n${N}_$1.json 2> gpurun_out/scale2_n${N}_$1.err
  echo "== $1 rc=$?"; head -c 260 gpurun_out/scale2_n${N}_$1.json; echo; grep -a "Error\|error" gpurun_out/scale2_n${N}_$1.err | head -3
}
run train_weak "--steps 20 --warmup 3"
run train_strong "--steps 20 --warmup 3 --scaling strong"
run refine "--workload refine --steps 4"
