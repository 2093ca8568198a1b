#!/bin/bash
# scripts/gpu_prof.sh [tag] -- launch list + full ncu capture of the IPM kernel (and optionally the lineariser)
TAG=${1:-x}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu > $OUT/bench_under_ncu_$TAG.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ipm_kernel -s 8 -c 1 -f -o $OUT/prof_ipm_$TAG \
    python bench.py --steps 3 --warmup 3 --no-cpu > $OUT/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
if [ "$2" == "lin" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:linearize_kernel -s 8 -c 1 -f -o $OUT/prof_lin_$TAG \
    python bench.py --steps 3 --warmup 3 --no-cpu > $OUT/ncu_full_lin_$TAG.log 2>&1; echo "ncu full lin rc=$?"
fi
