#!/bin/bash
# torch-DDP (eager launches) training step only, N GPUs: bash tools/gpu_r02_train_n.sh N
set +e
N=$1
O=gpurun_out
mkdir -p $O
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
$RUN bench.py --gpus $N --config train --train-eager --steps 10 --warmup 3 > $O/r02w_train_${N}gpu.json 2> $O/r02w_train_${N}gpu.err
cat $O/r02w_train_${N}gpu.json | cut -c1-250; grep -o '"ms_per_step_without_allreduce.*' $O/r02w_train_${N}gpu.json
tail -2 $O/r02w_train_${N}gpu.err
