#!/bin/bash
# Round-2 GPU session E: tower kernels after the gather / rulebook / pool rewrites.
set +e
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_backbone.py -q -x 2>&1 | tail -30 > $O/r02e_t_backbone.log
tail -6 $O/r02e_t_backbone.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"spb_|sparse_conv|sp_nn|sp_bucket|pm_gemm|fda_" -c 400 --csv --log-file $O/r02e_launches.csv \
  python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > $O/r02e_ncu_l.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r02e_bench_points.json 2> $O/r02e_bench_points.err
head -c 300 $O/r02e_bench_points.json; echo; tail -3 $O/r02e_bench_points.err
