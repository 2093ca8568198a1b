#!/bin/bash
# evidence run of round 2 (final tree): GPU tests, both bench arms (default command lines), phase profiles, saturated-start probe,
# single-instance latency (config 1), ncu of the EKF kernel.
T=${1:-r02e}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $O/${T}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/${T}_pytest_gpu.log; tail -2 $O/${T}_pytest_gpu.log
timeout 600 python bench.py --impl reference > $O/${T}_bench_reference_arm.json 2> $O/${T}_bench_ref.err; echo "bench ref rc=$?"
timeout 900 python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"; tail -2 $O/${T}_bench.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; tail -1 $O/${T}_smoke.log
BR2_VARIANT=prof timeout 300 python scripts/phase_profile.py > $O/${T}_phase_profile.json 2> $O/${T}_phase.err
BR2_VARIANT=prof timeout 300 python scripts/ekf_phase_profile.py > $O/${T}_ekf_phase_profile.json 2> $O/${T}_ekf_phase.err
timeout 300 python scripts/active_set_probe.py > $O/${T}_active_set_probe.json 2> $O/${T}_probe.err; tail -3 $O/${T}_probe.err
timeout 300 python tests/tools/single_latency.py > $O/${T}_single_instance_latency.json 2> $O/${T}_single.err; cat $O/${T}_single_instance_latency.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ekf_kernel -s 6 -c 1 -f -o $O/prof_ekf_$T python scripts/ekf_phase_profile.py > $O/${T}_ncu_ekf.log 2>&1; echo "ncu ekf rc=$?"
python - <<PY
import json
d=json.loads(open("$O/${T}_bench.json").read().strip().splitlines()[-1]); r=json.loads(open("$O/${T}_bench_reference_arm.json").read().strip().splitlines()[-1])
print("ours", round(d["value"]), d["ms_per_step"], d["kernels"], round(d["roofline"]["frac"],4), "e2e", round(d["e2e"]["value"]), "cpu", round(d["cpu_baseline"]["value"]), d["cpu_baseline"]["cores"])
print("ref arm", round(r["value"]), r["cpu_baseline"]["cores"], "same_config", d["config"]==r["config"])
s=d["sub_records"]; print("forced_ipm", round(s["forced_ipm"]["value"]), round(s["forced_ipm"]["roofline"]["frac"],4)); print("sat", s["saturated_start"]["tick_ms"], s["saturated_start"]["qp_ms_per_tick"])
print("dob", round(s["config3_dob"]["value"]), round(s["config3_dob"]["e2e"]["value"])); print("explicit", d["e2e_explicit_yref"]["value"], d["e2e_explicit_yref"]["fraction_of_windowed_e2e"]); print("like", d["cpu_baseline"]["like_for_like"])
PY
