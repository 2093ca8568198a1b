#!/bin/bash
# ncu evidence of round 2: launch list of one short bench run + full captures of the three hot kernels.  Everything in gpurun_out/.
T=${1:-r02}
O=gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --quick"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches.csv $CMD > $O/${T}_under_ncu.log 2>&1; echo "list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pdas_kernel -s 10 -c 2 -f -o $O/prof_pdas_$T $CMD > $O/${T}_ncu_pdas.log 2>&1; echo "pdas rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:linearize_kernel -s 10 -c 1 -f -o $O/prof_lin_$T $CMD > $O/${T}_ncu_lin.log 2>&1; echo "lin rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ipm_kernel -s 10 -c 1 -f -o $O/prof_ipm_$T $CMD --no-fast-path > $O/${T}_ncu_ipm.log 2>&1; echo "ipm rc=$?"
ls -la $O/*.ncu-rep
