#!/bin/bash
# GPU call: the staged tests (never run on a B200 so far), the whole GPU suite, the bench line
set -u
mkdir -p gpurun_out
run() { local name=$1 limit=$2; shift 2; echo "== $name"; timeout "$limit" "$@" > "gpurun_out/r2r_$name.log" 2>&1; echo "   rc=$? $(tail -1 "gpurun_out/r2r_$name.log" | cut -c1-500)"; }
WGBS_STAGED=1 run staged 300 python -m pytest tests/test_zz_bamdev_gpu.py -m gpu -x -q -p no:cacheprovider -k "device_parts or chromosomes_in_flight"
run suite 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider
python bench.py > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; echo "bench rc=$?"
