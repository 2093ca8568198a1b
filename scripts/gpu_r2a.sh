#!/bin/bash
# round 2, step A: PDAS default-on + LPT order.  tests, saturated-start probe, headline + forced-IPM bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 300 python scripts/active_set_probe.py > gpurun_out/r2a_probe.json 2> gpurun_out/r2a_probe.err; tail -3 gpurun_out/r2a_probe.err
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; cat gpurun_out/r2a_bench.json | head -c 1500
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-fast-path > gpurun_out/r2a_bench_ipm.json 2>> gpurun_out/r2a_bench.err; cat gpurun_out/r2a_bench_ipm.json | head -c 600
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --pos-spread 3.0 > gpurun_out/r2a_bench_sat.json 2>> gpurun_out/r2a_bench.err; cat gpurun_out/r2a_bench_sat.json | head -c 600
