#!/usr/bin/env python
"""One pass over every kernel of the library on mid-sized inputs (for `ncu --set full`: each kernel gets profiled launches without
the bench's repetitions):  python tools/kernels_probe.py [reads]
text route (tokenizer, pairing, calls, merge, pat2beta, collapse, format), BAM route (two-phase inflate, record table, filters), the pat
parser + pat2beta + homog on 4M records, segment, the MM/ML pileup, cview, beta_to_blocks."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000
    from wgbs_tools_b200 import bamio, synth
    g = bench.genome()
    sam = bench.make_batch(n_reads, 1000)
    bam = bamio.sam_to_bam(sam, [(bench.CHR, bench.CHR_LEN)], procs=8)
    from wgbs_tools_b200.api import Context
    reps = int(os.environ.get("PROBE_REPS", "1"))
    with Context(0) as ctx:
        ix = ctx.load_index(g.loci, 1)
        d_sam = ctx.upload(sam)
        mc = ctx.alloc(g.n_cpg * 8)
        for _ in range(reps):
            P, st = ctx.pileup_sam(ix, d_sam)
            ctx.pat2beta(P, 1, g.n_cpg + 1, meth_cov=mc, zero_first=True)
            P.collapse(); txt = P.to_text(bench.CHR); P.free()
            with bamio.DeviceBam.from_bytes(ctx, bam) as db:
                P, st = db.pileup(ix, bench.CHR)
                P.collapse(); txt2 = P.to_text(bench.CHR); P.free()
            assert txt == txt2
        # pat side
        N, R = bench.N_CPG, 4_000_000
        pt = synth.make_pat_text_fast(3, R, N, chrom=bench.CHR)
        d_pt = ctx.upload(pt)
        blocks = synth.make_blocks(5, 1, N)
        for _ in range(reps):
            Pt = ctx.pats_from_text(d_pt)
            ctx.pat2beta(Pt, 1, N + 1, meth_cov=mc, zero_first=True)
            ctx.homog(Pt, blocks, np.array([0, 0.334, 0.667, 1], np.float32), 3)
            V = Pt.cview(blocks[:2000, 0], blocks[:2000, 1]); V.collapse(mode=2); V.free()
            Pt.free()
        beta = synth.make_betas(9, 1, N)[0]
        ctx.beta_to_blocks(beta, blocks[:, 0], blocks[:, 1], 8)
        # segment: K = 10 x 120 000 sites
        K, S = 10, 120_000
        betas = synth.make_betas(9, K, S)
        for _ in range(reps):
            ctx.segment(betas, g.loci[:S], [(s, 60_000) for s in range(0, S, 60_000)], 1000, 2000, 15)
        # MM/ML pileup
        npsam = synth.make_np_sam(g, 60_000, 7)
        for _ in range(reps):
            Pn, _ = ctx.pileup_sam(ix, npsam); Pn.collapse(); Pn.free()
        print("probe done:", st["lines"], "records,", len(txt), "bytes of pat text")


if __name__ == "__main__":
    main()
