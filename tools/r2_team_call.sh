#!/bin/bash
# GPU call: the device BAM suite after the removal of the thread-per-block decoder and the match-per-lane replay
set -u
mkdir -p gpurun_out
run() { local name=$1 limit=$2; shift 2; echo "== $name"; timeout "$limit" "$@" > "gpurun_out/r2v_$name.log" 2>&1; echo "   rc=$? $(tail -1 "gpurun_out/r2v_$name.log" | cut -c1-600)"; }
run tests_bamdev 600 python -m pytest tests/test_zz_bamdev_gpu.py tests/test_zzzz_last_gpu.py tests/test_cli_gpu.py -m gpu -x -q -p no:cacheprovider
run probe_default 120 python tools/inflate_probe.py gpurun_in/bench.bam 8
