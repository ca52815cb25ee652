#!/usr/bin/env python
"""Time (and, under ncu, profile) the device BGZF inflate on one file:  python tools/inflate_probe.py FILE.bam [reps]
Prints one JSON line: bytes in / out, ms per launch of every inflate kernel (CUDA events around each launch), GB/s of inflated
output over the sum, and whether the output equals zlib's (host gunzip of the same blocks)."""
import gzip
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wgbs_tools_b200.api import Context  # noqa: E402

path = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
raw = open(path, "rb").read()
with Context(0) as ctx:
    out = ctx.bgzf_inflate(raw); n = out.nbytes          # warm-up (allocator pool, module load)
    same = None
    if os.environ.get("WGBS_PROBE_CHECK", "1") == "1":
        # (gzip.decompress of a file of thousands of members takes minutes in CPython 3.12: block by block with zlib instead)
        import struct
        import zlib
        from wgbs_tools_b200.csi import bgzf_blocks
        parts = [zlib.decompress(raw[off + 12 + struct.unpack_from("<H", raw, off + 10)[0]: off + bsize - 8], -15) for off, bsize, _ in bgzf_blocks(raw)]
        same = out.to_host().tobytes() == b"".join(parts)
    out.free()
    ctx.prof(True)
    for _ in range(reps):
        ctx.bgzf_inflate(raw).free()
    rep = ctx.prof_report()
    ctx.prof(False)
    ks = {k: ms / c for k, (c, ms) in rep.items() if k.startswith("bgzf_")}
    tot = sum(ks.values())
    print(json.dumps({"file": os.path.basename(path), "variant": os.environ.get("WGBS_INFLATE", "default"), "kernels_ms": ks, "compressed_bytes": len(raw),
                      "inflated_bytes": n, "ms_total": tot, "inflated_GBps": n / (tot / 1e3) / 1e9, "equals_zlib": same}))
