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


def test_direct_route_reads_mm_ml_tags_in_place(ctx, oracle, built_lib, monkeypatch):
    """MM/ML batches over the direct route (bam_np_tags_k: the MM:Z string and the ML:B:C array are read where they stand in the
    inflated stream; np.cu reads binary CIGARs, 4-bit bases and uint8 ML values): same templates / counters as the text route
    (view -> SAM text -> tokenizer) and == the oracle (patter --nanopore), for auto-detected and forced --nanopore, several option
    sets, and hand-built odd tags: legacy Mm / Ml names, two MM tags (the last one counts), ML of another subtype (ignored), no ML,
    no tags at all, an empty ML array, tags in front of and behind others, SEQ '*'"""
    from wgbs_tools_b200 import bamio
    H = oracle
    g = synth.make_genome(7, "chrT", 400_000)
    sam = synth.make_np_sam(g, 3_000, 9)
    seq = g.bases[1000:1060].tobytes()
    odd = [b"MM:Z:C+m?,0,1;\tML:B:C,250,3", b"Mm:Z:C+m.,1;\tMl:B:C,200", b"XA:i:5\tMM:Z:C+h?,0;\tXB:Z:abc\tML:B:C,255\tXC:B:s,1,2", b"MM:Z:C+m?,5;\tMM:Z:C+m?,0;\tML:B:C,9",
           b"MM:Z:C+m?,0;\tML:B:c,100", b"MM:Z:C+m?,0;", b"XZ:Z:none", b"MM:Z:C+m?,0;\tML:B:S,300", b"MM:Z:C+m?,0,0;C+h?,0,0;\tML:B:C,1,2,3,4", b"MM:Z:C+C?,0;\tML:B:C,77",
           b"MM:Z:C+m?;\tML:B:C", b"ML:B:C,5\tMM:Z:C+m.,0;"]
    extra = b"".join(b"odd%d\t%d\tchrT\t1001\t60\t60M\t*\t0\t0\t%s\t*\t%s\n" % (k, 16 * (k & 1), seq, t) for k, t in enumerate(odd))
    extra += b"star\t0\tchrT\t1001\t60\t60M\t*\t0\t0\t*\t*\tMM:Z:C+m?,0;\tML:B:C,200\n"
    first = sam[:sam.index(b"\n") + 1]
    for text, force in ((sam, False), (first + extra + sam[len(first):], False), (extra, True)):
        lines = sorted(text.splitlines(keepends=True), key=lambda l: int(l.split(b"\t")[3]))
        text = b"".join(lines)
        ix = ctx.load_index(g.loci, g.first_idx)
        with bamio.DeviceBam.from_bytes(ctx, bamio.sam_to_bam(text, [("chrT", g.length)])) as db:
            for kw in (dict(), dict(np_thresh=0.8, cpc_call="H"), dict(combine_mods=True, clip=3)):
                res = []
                for direct in ("0", "1"):
                    monkeypatch.setenv("WGBS_DBAM_DIRECT", direct)
                    P, st = db.pileup(ix, "chrT", nanopore=force, **kw)
                    P.collapse()
                    res.append((P.to_text("chrT"), {k: st[k] for k in ("lines", "pairs", "empty", "short", "invalid", "paired", "nanopore", "templates")}))
                    P.free()
                assert res[0] == res[1] and res[0][1]["nanopore"] == 1, (force, kw, res[0][1], res[1][1])
                pout, pst = H.port_patter(text, g.loci, g.idx(), nanopore=True, **kw)
                assert res[1][0] == H.port_collapse(pout), (force, kw)
        ix.free()
