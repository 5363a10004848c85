#!/bin/bash
set +e
O=gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sparse_conv3|spb_rulebook|spb_emit|spb_build" -s 39 -c 13 -o $O/r02f_towers \
  python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > $O/r02f_ncu.log 2>&1
tail -2 $O/r02f_ncu.log | head -c 300
ls -la $O/r02f_towers.ncu-rep
