#!/bin/bash
# training-path bring-up: unit tests of train_tail, the training gradient tests, the 1-GPU train line, a pm_gemm regression run
set +e
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_train_tail.py -x -q > $O/r02t_train_tail.log 2>&1
tail -25 $O/r02t_train_tail.log
timeout 600 python -m pytest tests/test_gpu_pose_model.py tests/test_gpu_pm_gemm.py -x -q -k "training or pm_gemm or gemm" > $O/r02t_pose.log 2>&1
tail -8 $O/r02t_pose.log
timeout 300 python bench.py --config train --steps 5 --warmup 3 > $O/r02t_train_1gpu.json 2> $O/r02t_train_1gpu.err
cut -c1-400 $O/r02t_train_1gpu.json; tail -3 $O/r02t_train_1gpu.err
