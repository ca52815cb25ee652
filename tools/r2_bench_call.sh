#!/bin/bash
# GPU call (2 GPUs): the bench line with 8 batches in flight at 1 and at 2 GPUs
set -u
mkdir -p gpurun_out
timeout 400 python bench.py --no-extras > gpurun_out/r2u_bench_1gpu.json 2> gpurun_out/r2u_bench_1gpu.err; echo "bench1 rc=$?"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 --no-extras > gpurun_out/r2u_bench_2gpu.json 2> gpurun_out/r2u_bench_2gpu.err; echo "bench2 rc=$?"
python - <<PY
import json
for f in ('1gpu','2gpu'):
    d=json.loads(open(f'gpurun_out/r2u_bench_{f}.json').read().strip().splitlines()[-1])
    print(f,'value',round(d['value']/1e6,1),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']/1e6,1),'ms',round(d['e2e']['ms_per_step'],3),'S',d['e2e']['batches_in_flight'],'steps',d['e2e']['steps_timed'],'serial',round(d['e2e']['serial']['ms_per_step'],3))
PY
