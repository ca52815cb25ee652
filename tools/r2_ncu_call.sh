#!/bin/bash
# GPU call: ncu --set full over one pass of every kernel (summarised ON the box: the report itself is too large to bring back), and the
# launch list of one bench run
set -u
mkdir -p gpurun_out
run() { local name=$1 limit=$2; shift 2; echo "== $name"; timeout "$limit" "$@" > "gpurun_out/r2d_$name.log" 2>&1; echo "   rc=$? $(tail -1 "gpurun_out/r2d_$name.log" | cut -c1-300)"; }
run ncu_all 900 ncu --set full --clock-control none --import-source on -o /tmp/r2d_all python tools/kernels_probe.py 200000
python tools/ncu_summarise.py /tmp/r2d_all.ncu-rep gpurun_out/r2d
ncu -i /tmp/r2d_all.ncu-rep --page raw --csv > gpurun_out/r2d_all_raw.csv 2>/dev/null
for k in bgzf_decode_k bgzf_resolve_k pileup_call_k sam_lines_k pat_lines_k pat2beta_k homog_k seg_dp_k seg_cost_k np_call_k; do
  ncu -i /tmp/r2d_all.ncu-rep --page source --csv --kernel-name regex:$k 2>/dev/null | head -4000 > gpurun_out/r2d_src_$k.csv
done
run launches 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 500 --csv --log-file gpurun_out/r2d_launches.csv python bench.py --steps 2 --warmup 3 --no-extras
du -sh gpurun_out; ls -la gpurun_out | head -30
