#!/bin/bash
set +e
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > $O/r02z_t_all.log 2>&1
tail -3 $O/r02z_t_all.log
timeout 300 python bench.py --config train --train-entry feats --steps 10 --warmup 3 > $O/r02z_train_feats_1gpu.json 2>$O/r02z_train_feats_1gpu.err
cut -c1-260 $O/r02z_train_feats_1gpu.json
timeout 300 python bench.py --config train --steps 10 --warmup 3 > $O/r02z_train_backbone_1gpu.json 2>$O/r02z_train_backbone_1gpu.err
cut -c1-260 $O/r02z_train_backbone_1gpu.json
timeout 300 python bench.py --steps 20 --warmup 5 > $O/r02z_bench_stage1.json 2> $O/r02z_bench_stage1.err
cut -c1-330 $O/r02z_bench_stage1.json; tail -2 $O/r02z_bench_stage1.err
