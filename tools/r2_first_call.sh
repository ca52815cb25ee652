#!/bin/bash
# Round-2 first GPU call: time the nl_scan variants, then ncu launch list + full captures of the main-path and side-path kernels.
set -u
mkdir -p gpurun_out
for v in default batch8 batch16 tma; do
  if [ $v = default ]; then unset WGBS_NLSCAN; else export WGBS_NLSCAN=$v; fi
  timeout 240 python bench.py --steps 10 --warmup 3 --no-extras > gpurun_out/r2a_bench_nlscan_$v.json 2> gpurun_out/r2a_bench_nlscan_$v.err; echo "bench nlscan=$v rc=$?"
done
unset WGBS_NLSCAN
NCU="ncu --set full --clock-control none --import-source on"
cap() { local name=$1 limit=$2; shift 2; echo "== $name"; timeout "$limit" "$@" > "gpurun_out/r2a_$name.log" 2>&1; echo "   rc=$?"; }
cap launches 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/r2a_launches.csv python bench.py --steps 2 --warmup 3 --no-extras
cap main 420 $NCU -k "regex:nl_scan_k|sam_lines_k|pileup_call_k|pileup_measure_k|pair_resolve_k|pair_insert_k|merge_templates_k|rs_onesweep_k|line_write_k|pat2beta_k" \
    -s 40 -c 12 -o gpurun_out/r2a_main python bench.py --steps 1 --warmup 3 --no-extras
cap extras 600 $NCU -k "regex:pat_lines_k|pat_pack_k|pat2beta_k|homog_k|seg_cost_k|seg_dp_k|seg_trace_k|np_measure_k|np_call_k|nl_count_k|nl_write_k" -c 12 -o gpurun_out/r2a_extras python bench.py --steps 1 --warmup 3
ls -la gpurun_out/r2a_* 2>/dev/null
