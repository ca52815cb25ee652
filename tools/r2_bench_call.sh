#!/bin/bash
# GPU call: the bench line (both arms) with the round's defaults
set -u
mkdir -p gpurun_out
python bench.py > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r2l_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2l_bench_ref.json 2> gpurun_out/r2l_bench_ref.err; echo "ref rc=$?"
