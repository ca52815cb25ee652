#!/bin/bash
# GPU call: byte replay with 4 independent steps in flight
set -u
mkdir -p gpurun_out
run() { local name=$1 limit=$2; shift 2; echo "== $name"; timeout "$limit" "$@" > "gpurun_out/r2n_$name.log" 2>&1; echo "   rc=$? $(tail -1 "gpurun_out/r2n_$name.log" | cut -c1-600)"; }
run tests_inflate 300 python -m pytest tests/test_zz_bamdev_gpu.py -m gpu -x -q -p no:cacheprovider -k inflate
run probe_default 120 python tools/inflate_probe.py gpurun_in/bench.bam 8
run e2e 300 python tools/e2e_probe.py 1000000 1,4
