#!/bin/bash
# First GPU call of the next round: run everything that was written without GPU access (see DESIGN.md "Staged"), each piece
# under its own time limit so that a hang costs one limit, not the call.  Usage (from the repo root):
#   gpurun --timeout 2400 -- 'bash tools/staged_gpu_run.sh'
# Results land in gpurun_out/staged_*.log; the bench line (with every staged child leg) in gpurun_out/staged_bench.json.
set -u
mkdir -p gpurun_out
export WGBS_STAGED=1
run() { local name=$1 limit=$2; shift 2; echo "== $name"; timeout "$limit" "$@" > "gpurun_out/staged_$name.log" 2>&1; echo "   rc=$? $(tail -1 "gpurun_out/staged_$name.log")"; }
# 1. the verified suite first (must stay green), then each staged test on its own
run suite 1500 python -m pytest tests -m gpu -x -q -k "not staged" -p no:cacheprovider
run inflate_teams 120 python -m pytest tests/test_zz_bamdev_gpu.py -m gpu -q -k "team_kernels"
run pat_tiles 120 python -m pytest tests/test_pat_gpu.py -m gpu -q -k "tile_parser"
run seg_plan 120 python -m pytest tests/test_segment_gpu.py -m gpu -q -k "exact_wave_plan or redux_argmax"
run dev_parts 180 python -m pytest tests/test_zz_bamdev_gpu.py -m gpu -q -k "device_parts"
run cli_streams 180 python -m pytest tests/test_zz_bamdev_gpu.py -m gpu -q -k "chromosomes_in_flight"
# newline-scan variants (the env var is read once per process: whole test files per variant, then a short bench for the kernel time)
for v in batch8 batch16 tma; do
  WGBS_NLSCAN=$v run nlscan_${v}_tests 300 python -m pytest tests/test_pileup_gpu.py -m gpu -x -q
  WGBS_NLSCAN=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-extras > gpurun_out/staged_bench_nlscan_$v.json 2> gpurun_out/staged_bench_nlscan_$v.err; echo "   bench nlscan=$v rc=$?"
done
# the official-format line with 2 and 3 batches in flight on the device-resident leg
for n in 2 3; do
  timeout 300 python bench.py --steps 12 --warmup 3 --no-extras --streams $n > gpurun_out/staged_bench_streams_$n.json 2> gpurun_out/staged_bench_streams_$n.err; echo "   bench streams=$n rc=$?"
done
# 2. the bench with all child legs (direct route, team decoders, batches in flight, segment at scale, pat parsers)
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/staged_bench.json 2> gpurun_out/staged_bench.err; echo "bench rc=$?"
python - <<'P'
import json
try:
    d = json.loads(open("gpurun_out/staged_bench.json").read().strip().splitlines()[-1])
    print("value %.1f M reads/s  e2e %.1f M reads/s" % (d["value"] / 1e6, d["e2e"]["value"] / 1e6))
    for k, v in (d.get("extra") or {}).items():
        if isinstance(v, dict):
            print(" ", k, {x: v[x] for x in ("ms_per_step", "reads_per_sec", "identical_to_sam_text_path", "error") if x in v} or list(v)[:6])
except Exception as e:
    print("no bench line:", e)
P
