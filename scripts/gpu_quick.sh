#!/bin/bash
# quick check after a kernel change: the GPU suite + kernel times against the batch size
O=gpurun_out; mkdir -p $O; T=${1:-quick}
timeout 900 python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/${T}_pytest.log
timeout 300 python scripts/batch_scaling_probe.py 148 592 2368 4096 8192 2>&1 | tee $O/${T}_batch_scaling.jsonl | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l); print(d['B'], round(d['tick_ms'],4), round(d['lin_ms'],4), round(d['qp_ms'],4))
    except Exception: print(l.strip()[:200])"
