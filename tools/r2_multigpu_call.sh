#!/bin/bash
# GPU call (2 GPUs): what the per-step NCCL reduce costs the end-to-end leg with batches in flight
set -u
mkdir -p gpurun_out
for mode in 1 0; do
  WGBS_BENCH_E2E_REDUCE=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$mode bench.py --gpus 2 --steps 40 --warmup 3 --no-extras > gpurun_out/r2q_bench_2gpu_reduce$mode.json 2> gpurun_out/r2q_bench_2gpu_reduce$mode.err; echo "reduce=$mode rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2q_bench_2gpu_reduce$mode.json').read().strip().splitlines()[-1])
print('value',round(d['value']/1e6,1),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']/1e6,1),'ms',round(d['e2e']['ms_per_step'],3),'serial',round(d['e2e']['serial']['ms_per_step'],3))
PY
done
