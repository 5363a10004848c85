#!/bin/bash
# Round-2 GPU session D: launch list of the tower kernels (filtered), fixed tests.
set +e
O=gpurun_out
mkdir -p $O
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"spb_|sparse_conv|sp_nn|sp_bucket|pm_gemm|fda_" -c 400 --csv --log-file $O/r02d_launches.csv \
  python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > $O/r02d_ncu_l.log 2>&1
tail -3 $O/r02d_ncu_l.log
timeout 600 python -m pytest tests/test_gpu_pm_gemm.py tests/test_gpu_pose_model.py -q -k "pm16_interpolation or refiner_golden" 2>&1 | tail -5
