#!/bin/bash
# GPU call: the two-phase BGZF decoder -- correctness first (each step under its own limit), then timings, then ncu, then the e2e probe.
set -u
mkdir -p gpurun_out
run() { local name=$1 limit=$2; shift 2; echo "== $name"; timeout "$limit" "$@" > "gpurun_out/r2b_$name.log" 2>&1; echo "   rc=$? $(tail -1 "gpurun_out/r2b_$name.log" | cut -c1-300)"; }
run tests_inflate 300 python -m pytest tests/test_zz_bamdev_gpu.py -m gpu -x -q -p no:cacheprovider
run probe_new 120 python tools/inflate_probe.py gpurun_in/bench.bam 8
NCU="ncu --set full --clock-control none --import-source on"
WGBS_PROBE_CHECK=0 run ncu_new 200 $NCU -k "regex:bgzf_decode_k|bgzf_resolve_k" -s 2 -c 2 -o gpurun_out/r2b_inflate2 python tools/inflate_probe.py gpurun_in/bench.bam 1

run e2e 400 python tools/e2e_probe.py 1000000 1,2,3,4
