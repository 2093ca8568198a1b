#!/bin/bash
# scripts/gpu_check.sh -- one gpurun call: GPU tests, bench (both arms), ncu launch list and one full capture of the
# dominant kernel.  Everything lands in gpurun_out/.  Usage: gpurun --timeout 1500 -- 'bash scripts/gpu_check.sh [tag]'
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $OUT/smi_$TAG.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu_$TAG.log
tail -3 $OUT/pytest_gpu_$TAG.log
timeout 600 python bench.py --impl reference > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "bench ref rc=$?"
timeout 600 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
cat $OUT/bench_$TAG.json; tail -5 $OUT/bench_$TAG.err
# launch list: same command, short
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu > $OUT/bench_under_ncu_$TAG.log 2>&1; echo "ncu list rc=$?"
# full capture of the Riccati IPM kernel (skip the recording pass: 6 launches, take 2 warm ones)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ipm_kernel -s 8 -c 2 -f -o $OUT/prof_ipm_$TAG \
    python bench.py --steps 3 --warmup 3 --no-cpu > $OUT/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:linearize_kernel -s 8 -c 1 -f -o $OUT/prof_lin_$TAG \
    python bench.py --steps 3 --warmup 3 --no-cpu > $OUT/ncu_full_lin_$TAG.log 2>&1; echo "ncu full lin rc=$?"
ls -la $OUT
