#!/bin/bash
set +e
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_tail.py tests/test_gpu_neighbour_ops.py tests/test_gpu_pm_gemm.py tests/test_gpu_engine.py -q > $O/r02y_tests.log 2>&1
tail -3 $O/r02y_tests.log
timeout 600 python tools/bench_ops.py > $O/r02y_bench_ops.log 2>&1
grep -E "ball_query|knn" $O/r02y_bench_ops.log | cut -c1-200
timeout 300 python tools/prof_train.py --rows 12 > $O/r02y_prof_train.txt 2>&1
cut -c1-150 $O/r02y_prof_train.txt
