#!/bin/bash
# final evidence pass 1: ncu counters of the neighbourhood ops and of the training kernels, full GPU test suite, smoke
set +e
O=gpurun_out
mkdir -p $O
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__inst_executed.sum,lts__t_bytes.sum
timeout 600 ncu --metrics $M --clock-control none --csv --log-file $O/r02x_ops_ncu.csv -k regex:"group|interp|gather|fps|ball|knn|three_nn|sp_" python tools/prof_bwd_ops.py > $O/r02x_ops_ncu.log 2>&1
timeout 900 ncu --metrics $M,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file $O/r02x_train_ncu.csv -k regex:"tr_|pm_gemm|fda_|pool_reduce" python tools/prof_train.py --ncu > $O/r02x_train_ncu.log 2>&1
tail -2 $O/r02x_train_ncu.log
timeout 1500 python -m pytest tests -m gpu -q > $O/r02x_t_all.log 2>&1
tail -5 $O/r02x_t_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r02x_smoke.log 2>&1
tail -2 $O/r02x_smoke.log
