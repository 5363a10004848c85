#!/bin/bash
# graph-mode training step (one CUDA graph + flat-gradient all-reduce) on N GPUs (N=1: plain python): bash tools/gpu_r02_train_graph_n.sh N
set +e
N=$1
O=gpurun_out
if [ "$N" = "1" ]; then RUN="python"; else RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"; fi
timeout 300 $RUN bench.py --gpus $N --config train --steps 10 --warmup 3 > $O/r02z_train_graph_${N}gpu.json 2> $O/r02z_train_graph_${N}gpu.err
cut -c1-250 $O/r02z_train_graph_${N}gpu.json; grep -o '"ms_per_step_without_allreduce.*' $O/r02z_train_graph_${N}gpu.json; grep -v Warning $O/r02z_train_graph_${N}gpu.err | tail -3
