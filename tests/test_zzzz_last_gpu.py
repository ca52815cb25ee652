"""GPU tests of paths that have not run on a B200 yet and are not selectable by a default: kept in the LAST file so that, under
`pytest -x`, nothing else depends on them."""
import gzip

import numpy as np
import pytest

from wgbs_tools_b200 import synth

pytestmark = pytest.mark.gpu


def test_pat_files_larger_than_one_call_go_piece_by_piece(ctx, tmp_path, monkeypatch):
    """pat2beta / homog on a pat text cut into pieces at line ends (WGBS_PAT_CHUNK_BYTES; a 30x pat file exceeds the 4 GiB one call
    takes): same .beta / same U-X-M table as in one call, for host text and for text inflated on the device"""
    from wgbs_tools_b200 import homog as hg
    from wgbs_tools_b200 import pat2beta as p2b
    from wgbs_tools_b200.patio import bgzf_compress
    N = 60_000
    txt = synth.make_pat_text_fast(5, 50_000, N, chrom="chr1")
    pg = tmp_path / "x.pat.gz"; pg.write_bytes(bgzf_compress(txt))
    blocks = synth.make_blocks(5, 1, N)
    bed = tmp_path / "b.bed"; bed.write_bytes(synth.blocks_text("chr1", blocks))
    out = {}
    for limit in ("0", "70000"):
        if limit != "0":
            monkeypatch.setenv("WGBS_PAT_CHUNK_BYTES", limit)
        for decode in ("host", "device"):
            d = tmp_path / f"o_{limit}_{decode}"; d.mkdir()
            p2b.pat2beta(ctx, str(pg), str(d), N, decode=decode)
            hg.main([str(pg), "-b", str(bed), "-o", str(d), "--pat_decode", decode])
            out[(limit, decode)] = ((d / "x.beta").read_bytes(), gzip.decompress((d / "x.uxm.bed.gz").read_bytes()))
    assert len(set(out.values())) == 1 and any(out[("0", "host")][0])


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
