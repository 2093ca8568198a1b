#!/bin/bash
# batch-scaling probe for library variants: usage gpu_var_scaling.sh tag1 tag2 ... ("" = default lib)
for v in "$@"; do
  echo "== variant '$v'"
  BR2_VARIANT=$v timeout 300 python scripts/batch_scaling_probe.py 148 2368 4096 8192 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l); print(d['B'], round(d['tick_ms'],4), round(d['lin_ms'],4), round(d['qp_ms'],4))
    except Exception: print(l.strip()[:200])"
done
