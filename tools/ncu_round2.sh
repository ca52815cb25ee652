#!/bin/bash
# ncu evidence for the next round in ONE GPU call (1 GPU; every capture under its own time limit):
#   gpurun --timeout 1500 -- 'bash tools/ncu_round2.sh'
# Writes gpurun_out/r2_*.ncu-rep (read here with `ncu -i ... --page raw --csv`) and the launch list of one bench step.
set -u
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
cap() { local name=$1 limit=$2; shift 2; echo "== $name"; timeout "$limit" "$@" > "gpurun_out/r2_$name.log" 2>&1; echo "   rc=$?"; }
# 1. launch list of the bench step (shares, not absolutes)
cap launches 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-extras
# 2. the kernels of the main path, one launch each (warm-up steps skipped by -s)
cap main 600 $NCU -k "regex:nl_scan_k|sam_lines_k|pileup_call_k|pileup_measure_k|pair_resolve_k|pair_insert_k|merge_templates_k|rs_onesweep_k|line_write_k|pat2beta_k" \
    -s 40 -c 12 -o gpurun_out/r2_main python bench.py --steps 1 --warmup 3 --no-extras
# 3. the side paths: pat text parser, pat2beta / homog at 16M records, segment, MM/ML pileup
cap extras 900 $NCU -k "regex:pat_lines_k|pat_pack_k|pat2beta_k|homog_k|seg_cost_k|seg_dp_k|np_measure_k|np_call_k" -c 10 -o gpurun_out/r2_extras python bench.py --steps 1 --warmup 3
# 4. the BGZF decoders on the bench BAM (default and the staged team decoders)
if [ -f gpurun_in/bench.bam ]; then
  for v in 2 g8 g16; do
    WGBS_INFLATE=$v cap inflate_$v 300 $NCU -k regex:bgzf_inflate -s 1 -c 1 -o gpurun_out/r2_inflate_$v python tools/inflate_probe.py gpurun_in/bench.bam 1
  done
fi
ls -la gpurun_out/r2_* 2>/dev/null
