#!/usr/bin/env python
"""Time (and, under ncu, profile) the device BGZF inflate on one file:  python tools/inflate_probe.py FILE.bam [reps]
Prints one JSON line: bytes in / out, ms per launch of bgzf_inflate_k (CUDA events around the launch), GB/s of inflated output."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wgbs_tools_b200.api import Context  # noqa: E402

path = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
raw = open(path, "rb").read()
with Context(0) as ctx:
    out = ctx.bgzf_inflate(raw); n = out.nbytes; out.free()          # warm-up (allocator pool, module load)
    ctx.prof(True)
    for _ in range(reps):
        ctx.bgzf_inflate(raw).free()
    rep = ctx.prof_report()
    ctx.prof(False)
    k, (c, ms) = next((k, v) for k, v in rep.items() if k.startswith("bgzf_inflate"))
    print(json.dumps({"file": os.path.basename(path), "variant": os.environ.get("WGBS_INFLATE", "default"), "kernel": k, "compressed_bytes": len(raw),
                      "inflated_bytes": n, "ms_per_launch": ms / c, "inflated_GBps": n / (ms / c / 1e3) / 1e9}))
