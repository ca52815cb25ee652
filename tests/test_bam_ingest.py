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


def test_view_filters_native_vs_sam_text(bamio, tmp_path):
    """wgbs_bam_view_ex (region, strand awk filters, -r RG, -L whitelist, bedtools -v blacklist, head -N) against the
    independent SAM-text implementation of the same samtools semantics (samfilter.filter_sam)"""
    from wgbs_tools_b200 import samfilter
    g = synth.make_genome(11, "chrT", 200_000)
    sam = synth.make_sam(g, 4000, 5, paired=True)
    rng = np.random.default_rng(3)
    lines = []
    for i, l in enumerate(sam.splitlines()):
        lines.append(l + (b"\tNM:i:1\tRG:Z:grp%d" % (i % 3) if i % 4 else b"\tXS:B:s,1,2"))          # some records without RG
    sam = b"\n".join(lines) + b"\n"
    p = tmp_path / "f.bam"; p.write_bytes(bamio.sam_to_bam(sam, [("chrT", g.length)]))
    starts = np.sort(rng.integers(0, g.length - 3000, 40)); bed = tmp_path / "l.bed"
    bed.write_text("# comment\ntrack name=x\n" + "".join(f"chrT\t{s}\t{s + int(w)}\n" for s, w in zip(starts.tolist(), rng.integers(1, 2500, 40).tolist())) + "chrOther\t5\t9\n")
    iv = samfilter.load_bed_intervals(str(bed))
    assert set(iv) == {"chrT", "chrOther"} and np.all(iv["chrT"][0][1:] > iv["chrT"][1][:-1])                 # merged, disjoint
    cases = [dict(), dict(beg=50_000, end=50_300), dict(flag_eq=(99, 147)), dict(flag_eq=(83, 163), mapq=10, exclude_flags=1796, include_flags=3),
             dict(read_group="grp1"), dict(read_group="grp"), dict(intervals=iv["chrT"]), dict(intervals=iv["chrT"], exclude_intervals=True),
             dict(intervals=iv["chrT"], beg=20_000, end=90_000, read_group="grp2", flag_eq=(99, 147)), dict(max_records=7),
             dict(max_records=5, flag_eq=(163,)), dict(intervals=(np.zeros(0, np.int64), np.zeros(0, np.int64))),
             dict(intervals=(np.zeros(0, np.int64), np.zeros(0, np.int64)), exclude_intervals=True)]
    with bamio.BamFile(str(p), threads=3) as b:
        for kw in cases:
            got = b.view("chrT", **kw)
            exp = samfilter.filter_sam(sam, chrom="chrT", **kw)
            assert got == exp, kw
            if kw is cases[5] or kw is cases[-2]:
                assert got == b""
            elif "max_records" in kw:
                assert got.count(b"\n") == kw["max_records"]
            else:
                assert 0 < got.count(b"\n") <= sam.count(b"\n")
        # a read ending exactly at an interval's start does not overlap; one base earlier does
        l0 = sam.splitlines()[10].split(b"\t"); pos0 = int(l0[3]) - 1; span = samfilter.ref_span(l0[5]); key = b"\t".join(l0[:4]) + b"\t"
        touch = (np.array([pos0 + span], np.int64), np.array([pos0 + span + 1], np.int64))
        assert not any(x.startswith(key) for x in b.view("chrT", intervals=touch).splitlines())
        over = (np.array([pos0 + span - 1], np.int64), np.array([pos0 + span + 1], np.int64))
        assert any(x.startswith(key) for x in b.view("chrT", intervals=over).splitlines())
        before = (np.array([max(0, pos0 - 3)], np.int64), np.array([pos0], np.int64))
        assert not any(x.startswith(key) for x in b.view("chrT", intervals=before).splitlines())


def test_genomic_region_follows_reference_rules(tmp_path):
    """-r / -s resolution (genomic_region.py:76-176): a CpG sitting exactly on the region's end is excluded; sites are
    end-exclusive; single positions; errors"""
    from wgbs_tools_b200.genome import GenomeRef, GenomicRegion, IllegalArgumentError, extend_region
    d = tmp_path / "g"; d.mkdir()
    loci = {"chr1": [100, 200, 300, 400, 500], "chr2": [50, 60], "chrX": [7]}
    i = 1; rows = []
    for c, ls in loci.items():
        for l in ls:
            rows.append(f"{c}\t{l}\t{i}\n"); i += 1
    (d / "CpG.bed.gz").write_bytes(gzip.compress("".join(rows).encode()))
    (d / "CpG.chrome.size").write_text("chr1\t5\nchr2\t2\nchrX\t1\n")
    (d / "chrome.size").write_text("chr1\t1000\nchr2\t100\nchrX\t10\n")
    ref = GenomeRef(str(d))
    G = lambda **kw: GenomicRegion(ref, **kw)
    assert G().is_whole() and G().region_str is None
    r = G(region="chr1:150-400"); assert r.sites == (2, 4) and r.bp_tuple == (150, 400) and r.region_str == "chr1:150-400" and r.nr_sites == 2
    assert G(region="chr1:150-401").sites == (2, 5)
    assert G(region="chr1:1,00-2,01").sites == (1, 3)
    assert G(region="chr1").sites == (1, 6) and G(region="chr1").region_str == "chr1"
    assert G(region="chr2").sites == (6, 8) and G(region="chrX").sites == (8, 9)
    assert G(region="chr1:200").sites == (2, 3) and G(region="chr1:200").region_str == "chr1:200-201"
    s = G(sites="2-4"); assert s.sites == (2, 4) and s.region_str == "chr1:200-301" and s.chrom == "chr1"
    assert G(sites="7").sites == (7, 8) and G(sites="7").region_str == "chr2:60-61"
    assert G(sites="3-3").sites == (3, 4)
    for bad in (dict(region="chr3"), dict(region="chr1:400-300"), dict(region="chr1:900-2000"), dict(region="chr1:101-199"), dict(region="chr1:150-200"),
                dict(region="1:5"), dict(region="chr1:x-y"), dict(sites="5-7"), dict(sites="0-3"), dict(sites="3-100"), dict(sites="a-b")):
        with pytest.raises(IllegalArgumentError):
            G(**bad)
    assert extend_region("chr1:1500-1800") == "chr1:500-2800" and extend_region("chr1:20-30") == "chr1:1-1030" and extend_region("chr1") == "chr1"


@pytest.mark.gpu
def test_bam2pat_cli_region_strand_readgroup_and_lists(ctx, oracle, bamio, tmp_path):
    """bam2pat -r / --top_strand / -rg / -L / --blacklist on a .bam: same pat + beta as the reference executables fed
    with the equivalently filtered SAM text and the dictionary of the region extended by 1000 bp (bam2pat.py:126-204)"""
    from wgbs_tools_b200 import bam2pat, samfilter
    from wgbs_tools_b200.genome import extend_region
    H = oracle
    g = synth.make_genome(43, "chr1", 500_000)
    refdir = tmp_path / "ref"; refdir.mkdir()
    with gzip.open(refdir / "CpG.bed.gz", "wb") as f:
        f.write(g.dict_text())
    (refdir / "CpG.chrome.size").write_text(f"chr1\t{g.n_cpg}\n"); (refdir / "chrome.size").write_text(f"chr1\t{g.length}\n")
    sam = synth.make_sam(g, 9000, 13, paired=True)
    sam = b"".join(l.rstrip(b"\n") + b"\tRG:Z:lib%d\n" % (i // 2 % 2) for i, l in enumerate(sam.splitlines(keepends=True)))   # mates share an RG... mostly
    bam = tmp_path / "s.bam"; bam.write_bytes(bamio.sam_to_bam(sam, [("chr1", g.length)]))
    rng = np.random.default_rng(1)
    st = np.sort(rng.integers(0, g.length - 5000, 60))
    bed = tmp_path / "list.bed"; bed.write_text("".join(f"chr1\t{s}\t{s + 1500}\n" for s in st.tolist()))
    iv = samfilter.load_bed_intervals(str(bed))["chr1"]
    dpath = H.write_tmp(g.dict_text(), ".CpG.bed")

    def expect(region, **kw):
        c, b, e = samfilter.parse_region_str(region)
        kept = samfilter.filter_sam(sam, 10, 1796, 3, chrom=c, beg=b, end=e, **kw)
        if H.have_ref():
            out, _ = H.ref_patter(kept, dpath, extend_region(region), True)
            return H.ref_collapse(out)
        _, xb, xe = samfilter.parse_region_str(extend_region(region))
        m = (g.loci >= xb) & (g.loci <= xe) if xe else np.ones(g.n_cpg, bool)
        out, _ = H.port_patter(H.port_match_maker(kept), g.loci[m], g.idx()[m])
        return H.port_collapse(out)

    runs = [(["-r", "chr1:100,000-180,000"], "chr1:100000-180000", {}, "s"),
            (["-r", "chr1:100000-180000", "--top_strand"], "chr1:100000-180000", dict(flag_eq=(147, 99)), "s"),
            (["--bottom_strand", "-rg", "lib1"], "chr1", dict(flag_eq=(83, 163), read_group="lib1"), "s.lib1"),
            (["-L", str(bed)], "chr1", dict(intervals=iv), "s"),
            (["--blacklist", str(bed), "-r", "chr1:200000-400000"], "chr1:200000-400000", dict(intervals=iv, exclude_intervals=True), "s")]
    for k, (argv, region, kw, name) in enumerate(runs):
        out = tmp_path / f"out{k}"; out.mkdir()
        bam2pat.main([str(bam), "--genome", str(refdir), "-o", str(out)] + argv)
        exp = expect(region, **kw)
        assert len(exp) > 1000
        assert gzip.decompress((out / f"{name}.pat.gz").read_bytes()) == exp, argv
        counts = H.port_pat2beta(exp, 1, g.n_cpg + 1)
        assert (out / f"{name}.beta").read_bytes() == H.ref_trim(counts).tobytes(), argv
    # -s: sites 1000-1200 -> region of their loci
    out = tmp_path / "outs"; out.mkdir()
    bam2pat.main([str(bam), "--genome", str(refdir), "-o", str(out), "-s", "1000-1200", "--no_beta"])
    reg = f"chr1:{g.loci[999]}-{g.loci[1198] + 1}"
    assert gzip.decompress((out / "s.pat.gz").read_bytes()) == expect(reg)
    assert not (out / "s.beta").exists()


def test_template_windows_partition_the_records_and_keep_mates_together(built_lib, tmp_path):
    """wgbs_view_opts.key_beg / key_end (bam2pat piles a large chromosome up in windows): the windows partition the view, no
    QNAME is split over two windows, and the SAM-text filter (filter_sam) selects the same records as the BAM reader"""
    from wgbs_tools_b200 import bamio
    from wgbs_tools_b200.samfilter import filter_sam
    g = synth.make_genome(7, "chrT", 300_000)
    sam = synth.make_sam(g, 20_000, 3, paired=True, single_frac=0.05)
    p = tmp_path / "w.bam"
    p.write_bytes(bamio.sam_to_bam(sam, [("chrT", g.length)]))
    edges = [0, 40_000, 40_300, 41_000, 150_000, 150_001, 299_000, 1 << 40]        # narrow windows: many pairs straddle an edge
    kw = dict(mapq=10, exclude_flags=1796, include_flags=3)
    with bamio.BamFile(str(p), threads=3) as b:
        whole = b.view("chrT", **kw)
        parts = [b.view("chrT", key_window=(lo, hi), **kw) for lo, hi in zip(edges[:-1], edges[1:])]
    assert sorted(whole.splitlines()) == sorted(l for t in parts for l in t.splitlines())
    assert sum(t.count(b"\n") for t in parts) == whole.count(b"\n")
    where = {}
    for k, t in enumerate(parts):
        for l in t.splitlines():
            assert where.setdefault(l.split(b"\t", 1)[0], k) == k                 # every record of a QNAME in ONE window
    assert len([t for t in parts if t]) >= 5
    for (lo, hi), t in zip(zip(edges[:-1], edges[1:]), parts):
        assert filter_sam(sam, chrom="chrT", key_window=(lo, hi), **kw) == t
    # every window is still in coordinate order (what the pairing and the pileup expect of a batch)
    for t in parts:
        pos = [int(l.split(b"\t")[3]) for l in t.splitlines()]
        assert pos == sorted(pos)


@pytest.mark.parametrize("budget", [60_000, 200_000, 1_500_000, 1 << 40])
def test_streamed_parts_give_every_template_once(built_lib, tmp_path, budget):
    """stream_parts (a .bam read as windows of BGZF blocks, records cut off at window ends, templates deferred until both mates
    are in): the union of the yielded (part, chromosome, key window) views == the whole-file views, and no QNAME is split"""
    from wgbs_tools_b200 import bamio
    g1 = synth.make_genome(7, "chrA", 200_000); g2 = synth.make_genome(8, "chrB", 120_000); g3 = synth.make_genome(9, "chrC", 50_000)
    sam = (synth.make_sam(g1, 9_000, 3, paired=True, single_frac=0.05, name_prefix="a") + synth.make_sam(g2, 5_000, 4, paired=True, name_prefix="b")
           + synth.make_sam(g3, 1_500, 5, paired=False, name_prefix="c") + b"u1\t4\t*\t0\t0\t*\t*\t0\t0\tACGT\t*\n" * 3)
    p = tmp_path / "s.bam"
    p.write_bytes(bamio.sam_to_bam(sam, [("chrA", g1.length), ("chrB", g2.length), ("chrC", g3.length), ("chrD", 1000)]))
    kw = dict(mapq=10, exclude_flags=1796)
    with bamio.BamFile(str(p), threads=2) as b:
        whole = {c: b.view(c, **kw) for c in b.refs}
    got = {c: [] for c in whole}; seen_done = set(); items = 0; where = {}
    opener = lambda data, refs, lens, first: bamio.BamPart(data, refs, lens, first, threads=2)
    for part, chrom, win, done in bamio.stream_parts(str(p), opener, lambda c: kw, budget):
        assert chrom not in seen_done
        t = part.view(chrom, key_window=win, **kw)
        for l in t.splitlines():
            q = l.split(b"\t", 1)[0]
            assert where.setdefault(q, items) == items, q                # every record of a QNAME in ONE item
        got[chrom].append(t); items += 1
        if done:
            seen_done.add(chrom)
    for c in whole:
        assert sorted(b"".join(got[c]).splitlines()) == sorted(whole[c].splitlines()), c
    assert seen_done >= {"chrA", "chrB", "chrC"}
    if budget < 1_000_000:
        assert items > 5                                                  # really streamed


def test_chromosome_block_ranges_without_an_index(built_lib, tmp_path):
    """chrom_block_ranges (binary search with wgbs_bam_probe) brackets every chromosome's records, and streaming only a
    chromosome's block range gives exactly that chromosome's whole-file view"""
    from wgbs_tools_b200 import bamio
    gs = [synth.make_genome(7 + i, f"chr{c}", L) for i, (c, L) in enumerate((("A", 150_000), ("B", 90_000), ("C", 30_000), ("E", 60_000)))]
    sam = b"".join(synth.make_sam(g, n, 3 + i, paired=True, name_prefix=f"{g.chrom}_") for i, (g, n) in enumerate(zip(gs, (7000, 4000, 300, 2500))))
    sam += b"u1\t4\t*\t0\t0\t*\t*\t0\t0\tACGT\t*\n" * 5
    refs = [("chrA", 150_000), ("chrB", 90_000), ("chrC", 30_000), ("chrD", 1000), ("chrE", 60_000)]          # chrD: no reads
    p = tmp_path / "s.bam"
    p.write_bytes(bamio.sam_to_bam(sam, refs))
    table = bamio.bgzf_block_table(str(p))
    ranges = bamio.chrom_block_ranges(str(p), table, len(refs))
    assert len(ranges) == len(refs) and ranges[0][0] == 0 and all(lo <= hi for lo, hi in ranges)
    kw = dict(mapq=10, exclude_flags=1796)
    names = [r[0] for r in refs]
    with bamio.BamFile(str(p), threads=2) as b:
        whole = {c: b.view(c, **kw) for c in names}
    opener = lambda data, r, l, f: bamio.BamPart(data, r, l, f, threads=2)
    for ci, c in enumerate(names):
        rng = ranges[ci]
        got = []
        for part, chrom, win, done in bamio.stream_parts(str(p), opener, lambda _c: kw, 150_000, blocks=rng, refs0=names):
            if chrom == c:
                got.append(part.view(chrom, key_window=win, **kw))
        assert sorted(b"".join(got).splitlines()) == sorted(whole[c].splitlines()), c


TUTORIAL = (("Lung_STL002.small.bam", 2793), ("Pancreas_STL002.small.bam", 814))


@pytest.mark.parametrize("name,n_lines", TUTORIAL)
def test_tutorial_bams_known_line_counts(built_lib, name, n_lines):
    """the reference's own tutorial (tutorial/README.md:72-80): `wgbstools bam2pat bams/*.bam -r chr3:119527929-119531943` reports
    "finished 2,793 lines" / "814 lines" -- what `samtools view BAM chr3:119527929-119531943 -q 10 -F 1796` prints (single-end data:
    no -f 3).  The BAMs (written by htslib, not by this repo) are the fixtures; the host reader must print that many lines."""
    from wgbs_tools_b200 import bamio
    bf = bamio.BamFile(os.path.join(GOLD, name), 2)
    txt = bf.view("chr3", mapq=10, exclude_flags=1796, beg=119527929, end=119531943)
    bf.close()
    assert txt.count(b"\n") == n_lines
    fl = {int(l.split(b"\t")[1]) for l in txt.splitlines()}
    assert not any(f & 1796 for f in fl) and all(int(l.split(b"\t")[4]) >= 10 for l in txt.splitlines())
