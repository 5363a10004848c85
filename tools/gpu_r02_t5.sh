#!/bin/bash
set +e
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_neighbour_ops.py -q -x > $O/r02v_neigh.log 2>&1
tail -4 $O/r02v_neigh.log
timeout 600 python tools/bench_ops.py > $O/r02v_bench_ops.log 2>&1
python - <<'PY'
import json
for l in open("gpurun_out/r02v_bench_ops.log"):
    if l.startswith("{"):
        d=json.loads(l)
        if "sp." in d["op"]: continue
        print(f'{d["op"]:50s} {d["ms"]*1e3:9.1f} us  ref {0 if not d["ref_kernel_ms"] else d["ref_kernel_ms"]*1e3:9.1f}  hbm {d["hbm_frac"]:.2f}  evals/s {d["evals_per_s"]}  exact {d["bit_exact_vs_ref"]}')
PY
cp gpurun_out/bench_ops.json $O/r02v_bench_ops.json 2>/dev/null
