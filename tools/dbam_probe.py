#!/usr/bin/env python
"""One pass of the device-BAM front end + pileup over a .bam of the bench workload (for `ncu` launch lists):
   python tools/dbam_probe.py FILE.bam [passes]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from wgbs_tools_b200 import bamio  # noqa: E402
from wgbs_tools_b200.api import Context  # noqa: E402

raw = open(sys.argv[1], "rb").read()
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
g = bench.genome()
with Context(0) as ctx:
    ix = ctx.load_index(g.loci, 1)
    for _ in range(passes):
        with bamio.DeviceBam.from_bytes(ctx, raw) as db:
            d = db.view_dev(bench.CHR)
            P, st = ctx.pileup_sam(ix, d)
            d.free()
            P.collapse()
            n = len(P)
            P.free()
    print(f"records {st['lines']:,} templates {st['templates']:,} collapsed {n:,} launches {ctx.launches}")
