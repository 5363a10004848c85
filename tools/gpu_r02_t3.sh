#!/bin/bash
set +e
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_train_tail.py tests/test_gpu_pose_model.py tests/test_gpu_fda.py tests/test_gpu_pm_gemm.py tests/test_gpu_engine.py -q > $O/r02t_tests3.log 2>&1
tail -6 $O/r02t_tests3.log
timeout 300 python bench.py --config train --steps 10 --warmup 3 > $O/r02t_train_1gpu.json 2> $O/r02t_train_1gpu.err
cut -c1-330 $O/r02t_train_1gpu.json; tail -2 $O/r02t_train_1gpu.err
timeout 300 python bench.py --config train --train-layers --steps 5 --warmup 3 > $O/r02t_train_1gpu_layers.json 2> $O/r02t_train_1gpu_layers.err
cut -c1-330 $O/r02t_train_1gpu_layers.json
timeout 300 python bench.py --steps 20 --warmup 5 > $O/r02t_bench_stage1.json 2> $O/r02t_bench_stage1.err
cut -c1-330 $O/r02t_bench_stage1.json; tail -2 $O/r02t_bench_stage1.err
