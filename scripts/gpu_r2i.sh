#!/bin/bash
T=${1:-r2i}
python -m pytest tests/test_bench_gpu.py tests/test_tick.py -x -q 2>&1 | tail -15
python bench.py --no-cpu > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -3 gpurun_out/${T}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench.json"))
print(d["value"], d["ms_per_step"], d["kernels"], d["roofline"]["frac"]); print(d["e2e"]); print(d.get("e2e_explicit_yref"))
s=d["sub_records"]; print("forced_ipm", s["forced_ipm"]["value"], s["forced_ipm"]["roofline"]["frac"], s["forced_ipm"]["kernels"])
print(s["saturated_start"]); print(s["config5_horizon_sweep"]); print(s["config3_dob"])
PY
