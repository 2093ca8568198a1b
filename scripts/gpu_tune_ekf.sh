#!/bin/bash
# scripts/gpu_tune_ekf.sh -- EKF kernel variants on the DOB-MPC workload (config 3): ms/step - lin - ipm = EKF time
for defs in "$@"; do
  echo "=== $defs"
  BR2_NVCC_DEFS="$defs" python -m bluerov2_b200.build --force > /dev/null && grep -A2 ekf_kernel bluerov2_b200/lib/ekf.ptxas.txt | grep -E "Used" | tr '\n' ' '; echo
  timeout 300 python bench.py --no-cpu --steps 50 --workload dob 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); k=d['kernels']; print('   value %.3e ms/step %.4f lin %.4f ipm %.4f -> ekf+gaps %.4f bad %d' % (d['value'], d['ms_per_step'], k['linearize_ms'], k['ipm_ms'], d['ms_per_step']-k['linearize_ms']-k['ipm_ms'], d['config']['nonzero_status']))"
done
