#!/bin/bash
set +e
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_tail.py tests/test_gpu_pose_model.py -q -k "train or training" > $O/r02z_tests10.log 2>&1
tail -3 $O/r02z_tests10.log
timeout 300 python bench.py --config train --steps 10 --warmup 3 > $O/r02z_train_graph_1gpu_b.json 2>$O/r02z_train_graph_1gpu_b.err
cut -c1-260 $O/r02z_train_graph_1gpu_b.json; grep -v Warn $O/r02z_train_graph_1gpu_b.err | tail -2
timeout 300 python bench.py --config train --train-entry feats --steps 10 --warmup 3 > $O/r02z_train_graph_feats_1gpu_b.json 2>/dev/null
cut -c1-260 $O/r02z_train_graph_feats_1gpu_b.json
