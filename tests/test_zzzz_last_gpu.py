"""GPU tests of paths that have not run on a B200 yet and are not selectable by a default: kept in the LAST file so that, under
`pytest -x`, nothing else depends on them."""
import numpy as np
import pytest

from wgbs_tools_b200 import synth

pytestmark = pytest.mark.gpu


def test_direct_route_hands_mm_ml_batches_to_the_text_route(ctx, oracle, built_lib, monkeypatch):
    """WGBS_DBAM_DIRECT=1 on a BAM whose first passing record carries an MM tag: wgbs_pileup_dbam must notice (bam_records_k) and
    take the text route (MM/ML tags are parsed from text): same templates as with the direct route off, and == the oracle"""
    from wgbs_tools_b200 import bamio
    H = oracle
    g = synth.make_genome(7, "chrT", 400_000)
    sam = synth.make_np_sam(g, 3_000, 9)
    ix = ctx.load_index(g.loci, g.first_idx)
    res = []
    with bamio.DeviceBam.from_bytes(ctx, bamio.sam_to_bam(sam, [("chrT", g.length)])) as db:
        for direct in ("0", "1"):
            monkeypatch.setenv("WGBS_DBAM_DIRECT", direct)
            P, st = db.pileup(ix, "chrT")
            P.collapse()
            res.append((P.to_text("chrT"), {k: st[k] for k in ("lines", "pairs", "empty", "short", "invalid", "paired", "nanopore", "templates")}))
            P.free()
    ix.free()
    assert res[0] == res[1] and res[0][1]["nanopore"] == 1 and res[0][1]["templates"] > 1000
    pout, pst = H.port_patter(sam, g.loci, g.idx(), nanopore=True)
    assert res[1][0] == H.port_collapse(pout)
