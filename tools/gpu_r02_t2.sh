#!/bin/bash
set +e
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_fda.py -x -q -k "backward" > $O/r02t_fda_bwd.log 2>&1
tail -5 $O/r02t_fda_bwd.log
timeout 600 python -m pytest tests/test_gpu_train_tail.py tests/test_gpu_pose_model.py -q -s -k "train or training or gradcheck" > $O/r02t_train2.log 2>&1
grep -E "worst|passed|failed|Error" $O/r02t_train2.log | cut -c1-900
timeout 300 python tools/prof_train.py --rows 16 > $O/r02t_prof_train2.txt 2>&1
cut -c1-150 $O/r02t_prof_train2.txt
