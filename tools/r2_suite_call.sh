#!/bin/bash
# GPU call: the whole GPU suite, then the bench (both arms)
set -u
mkdir -p gpurun_out
run() { local name=$1 limit=$2; shift 2; echo "== $name"; timeout "$limit" "$@" > "gpurun_out/r2c_$name.log" 2>&1; echo "   rc=$? $(tail -1 "gpurun_out/r2c_$name.log" | cut -c1-400)"; }
run suite 1500 python -m pytest tests -m gpu -q -p no:cacheprovider
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2c_bench.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2c_bench_ref.json 2> gpurun_out/r2c_bench_ref.err; echo "ref rc=$?"
python - <<'P'
import json
try:
    d = json.loads(open("gpurun_out/r2c_bench.json").read().strip().splitlines()[-1])
    print("value %.1f M  value_bam %.1f M  e2e %.1f M (serial %.1f M)" % (d["value"] / 1e6, d["value_bam"]["value"] / 1e6, d["e2e"]["value"] / 1e6, d["e2e"]["serial"]["value"] / 1e6))
    print("parity", d.get("parity"), "routes", d.get("routes_identical"))
    print("cpu", {k: v for k, v in (d.get("cpu_baseline") or {}).items() if k != "sample"})
    for k in ("pat2beta", "homog", "segment", "pileup_mm_ml"):
        v = d.get(k) or {}
        print(k, v.get("value"), (v.get("cpu_baseline") or {}).get("value"), v.get("error"))
    r = json.loads(open("gpurun_out/r2c_bench_ref.json").read().strip().splitlines()[-1])
    print("reference arm", r["value"], r["cpu_baseline"]["cores"], r["ms_per_step"])
except Exception as e:
    print("no bench line:", e)
P
