#!/bin/bash
# scripts/gpu_tune.sh -- rebuild with different knobs on the GPU box and bench each (tuning experiment)
for defs in "-DBR2_NSLOT=2" "-DBR2_NSLOT=2 -DBR2_EXP_EXTRACOPY"; do
  echo "=== $defs"
  BR2_NVCC_DEFS="$defs" python -m bluerov2_b200.build --force > /dev/null && grep -E "Used" bluerov2_b200/lib/kernels.ptxas.txt | tr '\n' ' '; echo
  for extra in "" "--no-fast-path"; do
  timeout 300 python bench.py --no-cpu --steps 50 $extra 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('   value %.3e lin %.3f ipm %.3f iters %.1f bad %d' % (d['value'], d['kernels']['linearize_ms'], d['kernels']['ipm_ms'], d['config']['mean_ipm_iterations'], d['config']['nonzero_status']))"
  done
done
