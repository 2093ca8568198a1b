#!/bin/bash
# N-GPU checks: peer exchange vs NCCL all-gather, then the bench in both exchange modes.  usage: gpurun --gpus N -- 'bash scripts/gpu_multi.sh N tag'
N=${1:-2}; T=${2:-r2m}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR scripts/shard_check.py 2> gpurun_out/${T}_shard_check.err | tail -1 | tee gpurun_out/${T}_shard_check.json
tail -3 gpurun_out/${T}_shard_check.err
for mode in peer nccl; do
  BR2_GATHER=$mode $TR bench.py --gpus $N --steps 200 --warmup 5 > gpurun_out/${T}_bench_${N}gpu_${mode}.json 2> gpurun_out/${T}_bench_${N}gpu_${mode}.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${T}_bench_${N}gpu_${mode}.json").read().strip().splitlines()[-1]); print("$mode", d["n_gpus"], d["config"]["global_batch"], round(d["value"]), d["ms_per_step"], d["kernels"], round(d["e2e"]["value"]))
except Exception as e: print("$mode ERR", e); print(open("gpurun_out/${T}_bench_${N}gpu_${mode}.err").read()[-1500:])
PY
done
python bench.py --quick --no-cpu --steps 200 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('1gpu', round(d['value']), d['ms_per_step'], d['kernels'])"
