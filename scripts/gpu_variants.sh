#!/bin/bash
# bench the occupancy variants built into bluerov2_b200/lib_<tag>/ (BR2_VARIANT): headline + saturated-start probe
for v in "$@"; do
  BR2_VARIANT=$v timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu > gpurun_out/var_${v}_bench.json 2> gpurun_out/var_${v}.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/var_${v}_bench.json")); print("$v", round(d["value"]), d["kernels"], round(d["roofline"]["frac"],4), round(d["e2e"]["value"]))
except Exception as e: print("$v", "ERR", e)
PY
done
