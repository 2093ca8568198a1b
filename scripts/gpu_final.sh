#!/bin/bash
# scripts/gpu_final.sh [tag] -- evidence run for a kernel revision: GPU tests, both bench arms, launch list, full ncu captures of the
# three kernels (IPM, lineariser, EKF), the configuration sweep.  Everything lands in gpurun_out/.
TAG=${1:-x}
OUT=gpurun_out
mkdir -p $OUT
bash scripts/gpu_check.sh $TAG 2>&1 | tail -12
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ekf_kernel -s 4 -c 1 -f -o $OUT/prof_ekf_$TAG \
    python bench.py --workload dob --steps 3 --warmup 3 --no-cpu > $OUT/ncu_full_ekf_$TAG.log 2>&1; echo "ncu full ekf rc=$?"
bash scripts/gpu_sweep.sh $TAG 2>&1 | tail -28
