#!/bin/bash
# scripts/gpu_tune_ipm.sh -- IPM kernel variants: rebuild with different knobs on the GPU box and bench each
run() {
  timeout 300 python bench.py --no-cpu --steps 50 "$@" 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('   value %.3e lin %.4f ipm %.4f iters %.1f bad %d frac %.3f' % (d['value'], d['kernels']['linearize_ms'], d['kernels']['ipm_ms'], d['config']['mean_ipm_iterations'], d['config']['nonzero_status'], d['roofline']['frac']))"
}
for defs in "$@"; do
  echo "=== $defs"
  BR2_NVCC_DEFS="$defs" python -m bluerov2_b200.build --force > /dev/null && grep -A2 ipm_kernel bluerov2_b200/lib/kernels.ptxas.txt | grep -E "Used" | tr '\n' ' '; echo
  run; run --no-fast-path
done
