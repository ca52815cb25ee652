#!/bin/bash
# GPU call: compute-sanitizer memcheck over the device BAM front end (team decoder, byte replay, MM/ML tags in place, record table)
set -u
mkdir -p gpurun_out
timeout 800 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_zz_bamdev_gpu.py tests/test_zzzz_last_gpu.py -m gpu -x -q -p no:cacheprovider -k "inflate or deep_codes or mm_ml or tutorial or direct" > gpurun_out/r2s_sanitizer.log 2>&1
echo "sanitizer rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2s_sanitizer.log | tail -5
