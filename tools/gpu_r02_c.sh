#!/bin/bash
# Round-2 GPU session C (1 GPU): device voxelisation + sparse-conv towers — parity tests, then bench from raw points.
set +e
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_backbone.py -q -x 2>&1 | tail -60 > $O/r02c_t_backbone.log
tail -25 $O/r02c_t_backbone.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_backbone.py 2>&1 | tail -15 > $O/r02c_t_rest.log
tail -5 $O/r02c_t_rest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r02c_bench_points.json 2> $O/r02c_bench_points.err
head -c 900 $O/r02c_bench_points.json; echo; tail -5 $O/r02c_bench_points.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --entry pyramids > $O/r02c_bench_pyramids.json 2> $O/r02c_bench_pyramids.err
head -c 400 $O/r02c_bench_pyramids.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02c_launches.csv \
  python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > $O/r02c_ncu_l.log 2>&1
