#!/bin/bash
# full GPU suite + quick bench + phase profile after the fused primal test / epilogue trim / EKF rewrite
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2r_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2r_pytest.log
timeout 600 python bench.py --quick > $O/r2r_bench.json 2> $O/r2r_bench.err; echo "bench rc=$?"; tail -2 $O/r2r_bench.err
BR2_VARIANT=prof timeout 300 python scripts/phase_profile.py > $O/r2r_phase.json 2> $O/r2r_phase.err
python - <<PY
import json
d=json.loads(open("$O/r2r_bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), d["ms_per_step"], d["kernels"], "frac", round(d["roofline"]["frac"],4), "e2e", round(d["e2e"]["value"]))
s=d.get("sub_records",{})
for k,v in s.items():
    print(k, {kk: vv for kk, vv in v.items() if kk in ("value","ms_per_step","kernels","tick_ms","qp_ms_per_tick")} if isinstance(v, dict) else v)
p=json.load(open("$O/r2r_phase.json"))
for k in ("fast_path","forced_ipm","saturated_tick0"):
    print(k, p[k]["ipm_ms_mean"], p[k]["cycles_per_instance_and_tick"])
PY
