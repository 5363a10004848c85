#!/bin/bash
# Round-2 GPU session A (1 GPU): parity tests, headline bench, stage-2 / training lines, TS probe, sanitizer logs.
# Everything it writes goes to gpurun_out/ (merged back by gpurun).
set +e
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -40 > $O/r02_t_all.log
python bench.py --steps 20 --warmup 5 > $O/r02_bench_1gpu.json 2> $O/r02_bench_1gpu.err
for T in 32 256 1024 4096; do
  python bench.py --config stage2 --batch-total $T --steps 5 --warmup 3 --no-cpu-baseline >> $O/r02_stage2_1gpu.jsonl 2>> $O/r02_stage2_1gpu.err
done
python bench.py --config train --steps 5 --warmup 3 > $O/r02_train_1gpu.json 2> $O/r02_train_1gpu.err
timeout 120 python tools/probe_umma_ts.py > $O/r02_probe_ts.log 2>&1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest -q -x \
  tests/test_gpu_engine.py::test_pipelined_two_streams_equals_sequential \
  "tests/test_gpu_engine.py::test_engine_padded_capacity_vs_oracle[64]" \
  tests/test_gpu_pose_model.py::test_tail_golden tests/test_gpu_pose_model.py::test_refiner_golden_and_loop \
  tests/test_gpu_pose_model.py::test_nearest_dist_propagates_nan \
  > $O/r02_sanitizer_memcheck_model.log 2>&1
echo "exit=$?" >> $O/r02_sanitizer_memcheck_model.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest -q -x tests/test_gpu_neighbour_ops.py -k "not 16384" \
  > $O/r02_sanitizer_memcheck_neighbour.log 2>&1
echo "exit=$?" >> $O/r02_sanitizer_memcheck_neighbour.log
tail -5 $O/r02_t_all.log
cat $O/r02_bench_1gpu.json | head -c 1500
