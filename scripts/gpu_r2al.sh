#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2al_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r2al_pytest.log
timeout 300 python bench.py --no-cpu --quick --no-fast-path --steps 30 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('forced ipm', round(d['value']), d['kernels'], d['mean_qp_iterations'], round(d['roofline']['frac'],4), d['nonzero_status'])"
timeout 300 python bench.py --no-cpu --quick --steps 100 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('default', round(d['value']), d['kernels'], round(d['roofline']['frac'],4))"
timeout 300 python scripts/active_set_probe.py 2>&1 | tail -3
