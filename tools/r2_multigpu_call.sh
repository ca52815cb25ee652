#!/bin/bash
# GPU call (2 GPUs): the four CLIs under torchrun at 1 and 2 GPUs on a 25-chromosome data set, then the bench line at 2 GPUs
set -u
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_zz_bamdev_gpu.py -m gpu -x -q -p no:cacheprovider -k "inflate" > gpurun_out/r2m_tests.log 2>&1; echo "tests rc=$? $(tail -1 gpurun_out/r2m_tests.log)"
timeout 700 python tools/multigpu_cli.py --scale 0.2 --reads 6000000 --pat_records 40000000 --K 20 --gpus 1,2 --out gpurun_out/r2m_multigpu.json > gpurun_out/r2m_multigpu.log 2>&1; echo "mg rc=$?"; tail -12 gpurun_out/r2m_multigpu.log | cut -c1-400
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2m_bench_2gpu.json 2> gpurun_out/r2m_bench_2gpu.err; echo "bench2 rc=$?"; tail -c 400 gpurun_out/r2m_bench_2gpu.json
