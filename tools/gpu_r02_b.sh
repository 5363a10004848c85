#!/bin/bash
# Round-2 GPU session B (1 GPU): fp16 operand path — parity tests, A/B bench, ncu launch list and full captures.
set +e
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > $O/r02b_t_all.log
tail -15 $O/r02b_t_all.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r02b_bench_fp16.json 2> $O/r02b_bench_fp16.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --precision fp32-faithful > $O/r02b_bench_fp32f.json 2> $O/r02b_bench_fp32f.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --streams 1 > $O/r02b_bench_fp16_1stream.json 2>> $O/r02b_bench_fp16.err
python bench.py --config stage2 --batch-total 32 --steps 10 --warmup 3 --no-cpu-baseline > $O/r02b_stage2_32.json 2>> $O/r02b_bench_fp16.err
head -c 600 $O/r02b_bench_fp16.json; echo; head -c 300 $O/r02b_bench_fp32f.json; echo; tail -3 $O/r02b_bench_fp16.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02b_launches.csv \
  python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > $O/r02b_ncu_l.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pm_gemm_pair|fda_pair" -s 9 -c 9 -o $O/r02b_gemm_fda \
  python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > $O/r02b_ncu_f.log 2>&1
ls -la $O | grep r02b
