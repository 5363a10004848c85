#!/bin/bash
# final evidence pass (1 GPU): ncu counters of one inference pass and one training step, full GPU test suite, smoke,
# default bench + reference arm, compute-sanitizer memcheck over the kernels added this round
set +e
O=gpurun_out
mkdir -p $O
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
timeout 600 ncu --metrics $M --clock-control none --csv --log-file $O/r02z_infer_ncu.csv python bench.py --steps 1 --warmup 2 --no-graph --no-cpu-baseline > $O/r02z_infer_ncu.log 2>&1
timeout 900 ncu --metrics $M --clock-control none --csv --log-file $O/r02z_train_ncu.csv -k regex:"tr_|pm_gemm|fda_|pool_reduce|sp_|svd3" python tools/prof_train.py --ncu --entry backbone > $O/r02z_train_ncu.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q > $O/r02z_t_all.log 2>&1
tail -4 $O/r02z_t_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r02z_smoke.log 2>&1
tail -1 $O/r02z_smoke.log
timeout 600 python bench.py > $O/r02z_bench_default.json 2> $O/r02z_bench_default.err
cut -c1-300 $O/r02z_bench_default.json; tail -2 $O/r02z_bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02z_bench_reference_arm.json 2> $O/r02z_bench_reference_arm.err
cut -c1-300 $O/r02z_bench_reference_arm.json
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest -q -x \
  "tests/test_gpu_train_tail.py::test_mlp_stacks_vs_fp64" \
  "tests/test_gpu_fda.py::test_fda_backward_fused_vs_fp64[2-64-128-128-relu]" \
  "tests/test_gpu_fda.py::test_fda_backward_fused_vs_fp64[2-128-384-128-relu]" \
  tests/test_gpu_neighbour_ops.py -k "mlp_stacks or backward or ball or knn" \
  > $O/r02z_sanitizer_memcheck_round2_kernels.log 2>&1
echo "exit=$?" >> $O/r02z_sanitizer_memcheck_round2_kernels.log
tail -3 $O/r02z_sanitizer_memcheck_round2_kernels.log
