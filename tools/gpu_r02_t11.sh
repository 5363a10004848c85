#!/bin/bash
set +e
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_neighbour_ops.py -q -k "interpolate" > $O/r02z_interp.log 2>&1
tail -2 $O/r02z_interp.log
timeout 600 python tools/bench_ops.py 2>/dev/null | grep -E '"op": "three_interpolate' | cut -c1-220
DCL_INTERP_ROWS=1 timeout 600 python tools/bench_ops.py 2>/dev/null | grep -E '"op": "three_interpolate fwd' | cut -c1-220
