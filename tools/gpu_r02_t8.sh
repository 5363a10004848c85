#!/bin/bash
set +e
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_pose_model.py -q -k "refiner_backward or training_step" > $O/r02z_so3.log 2>&1
tail -3 $O/r02z_so3.log
timeout 300 python bench.py --config train --train-graph --train-entry feats --steps 10 --warmup 3 > $O/r02z_train_graph_feats_1gpu.json 2>$O/r02z_train_graph_feats_1gpu.err
cut -c1-260 $O/r02z_train_graph_feats_1gpu.json; tail -3 $O/r02z_train_graph_feats_1gpu.err
timeout 300 python bench.py --config train --train-graph --steps 10 --warmup 3 > $O/r02z_train_graph_backbone_1gpu.json 2>$O/r02z_train_graph_backbone_1gpu.err
cut -c1-260 $O/r02z_train_graph_backbone_1gpu.json; tail -3 $O/r02z_train_graph_backbone_1gpu.err
timeout 300 python bench.py --config train --train-entry feats --steps 10 --warmup 3 > $O/r02z_train_eager_feats_1gpu.json 2>/dev/null
cut -c1-260 $O/r02z_train_eager_feats_1gpu.json
