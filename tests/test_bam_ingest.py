"""BAM ingest (csrc/bam.cu, host code: runs without a GPU): BGZF inflate + BAM -> the SAM text `samtools view` prints.
samtools is not installed, so parity is a round trip: SAM -> (Python BAM writer) -> BAM -> (native reader) -> SAM."""
import gzip
import os

import numpy as np
import pytest

from wgbs_tools_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def bamio(built_lib):
    from wgbs_tools_b200 import bamio
    return bamio


def test_sam_bam_sam_round_trip_with_all_tag_types(bamio, tmp_path):
    g = synth.make_genome(7, "chrT", 300_000)
    sam = synth.make_sam(g, 5000, 3, paired=True)
    extra = (b"x1\t0\tchrT\t500\t7\t10M2I5M3D20M4S\t*\t0\t0\t" + b"ACGTN" * 8 + b"A\t*\tXA:A:q\tXB:i:-5\tXC:i:300\tXD:i:70000\tXE:f:1.5"
             b"\tXF:Z:hello world\tXG:H:1AE3\tML:B:C,1,2,255\tXH:B:s,-3,400\tXI:B:f,0.25,2\tMM:Z:C+m?,0,1;\n"
             b"x2\t4\t*\t0\t0\t*\t*\t0\t0\t*\t*\n")
    full = extra.splitlines(keepends=True)[0] + sam + extra.splitlines(keepends=True)[1]      # unmapped record last
    p = tmp_path / "t.bam"
    p.write_bytes(bamio.sam_to_bam(full, [("chrM", 16571), ("chrT", g.length)]))
    with bamio.BamFile(str(p), threads=4) as b:
        assert b.refs == ["chrM", "chrT"] and b.nrecords() == full.count(b"\n")
        assert b.view() == full
        assert b.view("chrT") == full[: full.rindex(b"x2\t")]
        assert b.view("chrM") == b""
        f = b.view("chrT", mapq=10, exclude_flags=1796, include_flags=3)                      # samtools view -q 10 -F 1796 -f 3
        exp = b"".join(l for l in sam.splitlines(keepends=True) if int(l.split(b"\t")[1]) & 3 == 3)
        assert f == exp
        # region restriction: records overlapping [20000, 21000]
        r = b.view("chrT", beg=20_000, end=21_000)
        pos = [int(l.split(b"\t")[3]) for l in r.splitlines()]
        assert pos and min(pos) > 20_000 - 200 and max(pos) <= 21_000
        assert "@SQ\tSN:chrT" in b.header


def test_corrupt_and_unsorted_inputs_fail_loudly(bamio, tmp_path):
    from wgbs_tools_b200._lib import WgbsError
    p = tmp_path / "bad.bam"; p.write_bytes(b"not a bam file at all........................")
    with pytest.raises(WgbsError, match="not a BGZF"):
        bamio.BamFile(str(p))
    sam = b"a\t0\tchr2\t5\t60\t4M\t*\t0\t0\tACGT\t*\nb\t0\tchr1\t5\t60\t4M\t*\t0\t0\tACGT\t*\nc\t0\tchr2\t9\t60\t4M\t*\t0\t0\tACGT\t*\n"
    q = tmp_path / "unsorted.bam"; q.write_bytes(bamio.sam_to_bam(sam, [("chr1", 100), ("chr2", 100)]))
    with pytest.raises(WgbsError, match="not sorted"):
        bamio.BamFile(str(q))
    with pytest.raises(WgbsError, match="cannot open"):
        bamio.BamFile(str(tmp_path / "missing.bam"))


def _tutorial():
    """reads + a dictionary placed where the reads themselves show CG dinucleotides (plus random extra loci)"""
    sam = gzip.open(os.path.join(GOLD, "tutorial_reads.sam.gz"), "rb").read()
    loci = set()
    for l in sam.splitlines():
        t = l.split(b"\t")
        if t[5].endswith(b"M") and t[5][:-1].isdigit():            # plain xM alignments: read offset == reference offset
            p0, sq = int(t[3]), t[9]
            loci.update(p0 + i for i in range(len(sq) - 1) if sq[i:i + 2] == b"CG")
    pos = np.array([int(l.split(b"\t")[3]) for l in sam.splitlines()])
    rng = np.random.default_rng(5)
    loci.update(rng.integers(int(pos.min()) - 50, int(pos.max()) + 400, size=300).tolist())
    loci = np.array(sorted(loci))
    loci = loci[np.concatenate([[True], np.diff(loci) >= 2])]
    return sam, loci.astype(np.uint32)


def test_real_reads_fixture_through_the_oracle(oracle):
    """real bisulfite reads (tutorial BAMs, CIGAR ops M/I/D/S/H) against a synthetic dictionary: port == reference"""
    H = oracle
    if not H.have_ref():
        pytest.skip("reference executables not built")
    sam, loci = _tutorial()
    idx = np.arange(1, loci.size + 1)
    d = H.write_tmp(b"".join(b"chr3\t%d\t%d\n" % (l, i) for l, i in zip(loci.tolist(), idx.tolist())), ".CpG.bed")
    out, err = H.ref_patter(sam, d, "chr3", False)
    pout, st = H.port_patter(sam, loci, idx)
    assert out == pout and len(out) > 5000 and max(len(l.split(b"\t")[2]) for l in out.splitlines()) >= 5
    assert {l.split(b"\t")[5][-1:] for l in sam.splitlines()} >= {b"M", b"S", b"H"}


@pytest.mark.gpu
def test_real_reads_fixture_on_gpu(ctx, oracle):
    H = oracle
    sam, loci = _tutorial()
    idx = np.arange(1, loci.size + 1)
    ix = ctx.load_index(loci, 1)
    for kw in (dict(), dict(clip=7, min_cpg=2)):
        P, st = ctx.pileup_sam(ix, sam, **kw)
        P.collapse()
        txt = P.to_text("chr3")
        P.free()
        pout, pst = H.port_patter(sam, loci, idx, **kw)
        assert txt == H.port_collapse(pout)
        assert [st[k] for k in ("lines", "pairs", "empty", "short", "invalid", "paired")] == pst
    ix.free()


@pytest.mark.gpu
def test_bam2pat_cli_reads_bam(ctx, oracle, bamio, tmp_path):
    """bam2pat on a .bam file: same .pat.gz / .beta as the reference pipeline fed by the equivalent SAM text"""
    from wgbs_tools_b200 import bam2pat
    H = oracle
    g = synth.make_genome(41, "chr1", 500_000)
    refdir = tmp_path / "ref"; refdir.mkdir()
    with gzip.open(refdir / "CpG.bed.gz", "wb") as f:
        f.write(g.dict_text())
    (refdir / "CpG.chrome.size").write_text(f"chr1\t{g.n_cpg}\n")
    sam = synth.make_sam(g, 8000, 9, paired=True)
    # sprinkle low-MAPQ and duplicate-flag records: the view filter must drop them exactly like `samtools view -q 10 -F 1796 -f 3`
    lines = sam.splitlines(keepends=True)
    for i in range(0, len(lines), 37):
        t = lines[i].split(b"\t"); t[4] = b"3"; lines[i] = b"\t".join(t)
    for i in range(5, len(lines), 53):
        t = lines[i].split(b"\t"); t[1] = b"%d" % (int(t[1]) | 1024); lines[i] = b"\t".join(t)
    sam = b"".join(lines)
    bam = tmp_path / "s.bam"; bam.write_bytes(bamio.sam_to_bam(sam, [("chr1", g.length)]))
    out = tmp_path / "out"; out.mkdir()
    bam2pat.main([str(bam), "--genome", str(refdir), "-o", str(out)])
    kept = b"".join(l for l in lines if int(l.split(b"\t")[4]) >= 10 and not int(l.split(b"\t")[1]) & 1796 and int(l.split(b"\t")[1]) & 3 == 3)
    pout, _ = H.port_patter(H.port_match_maker(kept), g.loci, g.idx())
    exp = H.port_collapse(pout)
    assert gzip.decompress((out / "s.pat.gz").read_bytes()) == exp
    counts = H.port_pat2beta(exp, 1, g.n_cpg + 1)
    assert (out / "s.beta").read_bytes() == H.ref_trim(counts).tobytes()
