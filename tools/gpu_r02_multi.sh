#!/bin/bash
# Round-2 multi-GPU session (r02w: final code of the round — training path on the tensor-core kernels): bash tools/gpu_r02_multi.sh N   (N = 2, 4 or 8 GPUs of one box)
# stage-1 headline (weak scaling), stage-2 instance-sharded (strong scaling, B_total 32..4096), DDP training step.
set +e
N=$1
O=gpurun_out
mkdir -p $O
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
$RUN bench.py --gpus $N --steps 20 --warmup 5 > $O/r02w_stage1_${N}gpu.json 2> $O/r02w_stage1_${N}gpu.err
head -c 250 $O/r02w_stage1_${N}gpu.json; echo
for T in 32 256 1024 4096; do
  if [ $((T / N)) -ge 1 ]; then
    $RUN bench.py --gpus $N --config stage2 --batch-total $T --steps 5 --warmup 3 >> $O/r02w_stage2_${N}gpu.jsonl 2>> $O/r02w_stage2_${N}gpu.err
  fi
done
cut -c1-200 $O/r02w_stage2_${N}gpu.jsonl
$RUN bench.py --gpus $N --config train --steps 5 --warmup 3 > $O/r02w_train_${N}gpu.json 2> $O/r02w_train_${N}gpu.err
cat $O/r02w_train_${N}gpu.json | head -c 1200; echo
tail -3 $O/r02w_train_${N}gpu.err
