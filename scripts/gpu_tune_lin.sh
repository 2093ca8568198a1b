#!/bin/bash
# scripts/gpu_tune_lin.sh -- lineariser variants: rebuild with different tiling knobs on the GPU box and bench each
OUT=gpurun_out; mkdir -p $OUT
run() {
  timeout 300 python bench.py --no-cpu --steps 50 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('   value %.3e lin %.4f ipm %.4f iters %.1f bad %d' % (d['value'], d['kernels']['linearize_ms'], d['kernels']['ipm_ms'], d['config']['mean_ipm_iterations'], d['config']['nonzero_status']))"
}
for defs in "$@"; do
  echo "=== $defs"
  BR2_NVCC_DEFS="$defs" python -m bluerov2_b200.build --force > /dev/null && grep -A2 linearize bluerov2_b200/lib/kernels.ptxas.txt | grep -E "Used" | tr '\n' ' '; echo
  run
done
