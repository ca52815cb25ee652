#!/bin/bash
# GPU call: probe of the replay with L2 prefetch, then the whole GPU suite
set -u
mkdir -p gpurun_out
run() { local name=$1 limit=$2; shift 2; echo "== $name"; timeout "$limit" "$@" > "gpurun_out/r2o_$name.log" 2>&1; echo "   rc=$? $(tail -1 "gpurun_out/r2o_$name.log" | cut -c1-500)"; }
run probe_default 120 python tools/inflate_probe.py gpurun_in/bench.bam 8
run suite 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider
