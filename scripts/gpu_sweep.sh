#!/bin/bash
# scripts/gpu_sweep.sh [tag] -- the non-headline BASELINE configurations on one GPU: config 3 (DOB-MPC), config 5
# (horizon sweep at batch 8192), the active-bound set, and a batch sweep.  One JSON line each -> gpurun_out/sweep_<tag>.jsonl
TAG=${1:-x}
OUT=gpurun_out/sweep_$TAG.jsonl
mkdir -p gpurun_out; : > $OUT
run() { echo "# $*" >> $OUT; timeout 300 python bench.py --no-cpu --steps 50 --warmup 5 "$@" 2>>gpurun_out/sweep_$TAG.err | tail -1 >> $OUT; }
run --workload dob
for n in 10 20 40 80; do run --batch 8192 --horizon $n; done
run --pos-spread 3.0
run --no-fast-path
for b in 1024 2048 8192 16384 32768; do run --batch $b; done
python - <<PY
import json
for l in open("$OUT"):
    if l.startswith("#"): print(l.strip()); continue
    try:
        d = json.loads(l)
        print("   value %.3e  e2e %.3e  ms/step %.3f  lin %.3f ms  ipm %.3f ms  iters %.2f  frac %.3f  bad %d" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["kernels"]["linearize_ms"], d["kernels"]["ipm_ms"], d["config"]["mean_ipm_iterations"], d["roofline"]["frac"], d["config"]["nonzero_status"]))
    except Exception as e:
        print("   ??", l[:200])
PY
