#!/bin/bash
# round 2, step B: cp.async staging ring + combined stage record.  tests (ipm-related first), probe, bench
T=${1:-r2b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -5 gpurun_out/${T}_pytest.log
timeout 300 python scripts/active_set_probe.py > gpurun_out/${T}_probe.json 2> gpurun_out/${T}_probe.err; tail -3 gpurun_out/${T}_probe.err
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; python - <<PY
import json
for f in ("gpurun_out/${T}_bench.json",):
    try:
        d=json.load(open(f)); print(f, d["value"], d["kernels"], d["roofline"]["frac"], d["e2e"]["value"])
    except Exception as e: print(f, "ERR", e)
PY
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-fast-path > gpurun_out/${T}_bench_ipm.json 2>> gpurun_out/${T}_bench.err; python - <<PY
import json
for f in ("gpurun_out/${T}_bench_ipm.json",):
    try:
        d=json.load(open(f)); print(f, d["value"], d["kernels"], d["roofline"]["frac"], d["config"]["mean_ipm_iterations"])
    except Exception as e: print(f, "ERR", e)
PY
