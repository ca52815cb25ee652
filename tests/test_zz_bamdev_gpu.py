"""BAM ingest on the device (csrc/bamdev.cu): warp-per-block BGZF inflate, record table, `samtools view` filters and SAM
formatting in HBM.  Everything is compared byte for byte with the host reader (csrc/bam.cu, zlib on host threads), which is
itself pinned by SAM -> BAM -> SAM round trips (test_bam_ingest.py); the per-block / per-record logic these kernels run is
additionally pinned without a GPU in test_bamdev_core.py.  (File name: runs after the other GPU suites.)"""
import gzip
import struct
import zlib

import numpy as np
import pytest

from wgbs_tools_b200 import synth

pytestmark = pytest.mark.gpu

EXTRA = (b"x1\t0\tchrT\t500\t7\t10M2I5M3D20M4S\t*\t0\t0\t" + b"ACGTN" * 8 + b"A\t*\tXA:A:q\tXB:i:-5\tXC:i:300\tXD:i:70000\tXE:f:1.5"
         b"\tXF:Z:hello world\tXG:H:1AE3\tML:B:C,1,2,255\tXH:B:s,-3,400\tXI:B:f,0.25,2,1e-07,3.14159\tMM:Z:C+m?,0,1;\tRG:Z:grp1\n")
UNMAPPED = b"x2\t4\t*\t0\t0\t*\t*\t0\t0\t*\t*\n"


@pytest.fixture(scope="module")
def bamio(built_lib):
    from wgbs_tools_b200 import bamio
    return bamio


@pytest.fixture(scope="module")
def sample(bamio, tmp_path_factory):
    g = synth.make_genome(7, "chrT", 300_000)
    lines = synth.make_sam(g, 30_000, 3, paired=True).splitlines()
    for i in range(0, len(lines), 3):
        lines[i] += b"\tRG:Z:grp1" if i % 2 == 0 else b"\tXX:i:5\tRG:Z:other"
    for i in range(0, len(lines), 7):
        t = lines[i].split(b"\t"); t[4] = b"3"; lines[i] = b"\t".join(t)
    full = EXTRA + b"\n".join(lines) + b"\n" + UNMAPPED
    p = tmp_path_factory.mktemp("dbam") / "f.bam"
    p.write_bytes(bamio.sam_to_bam(full, [("chrM", 16571), ("chrT", g.length)]))
    return g, full, str(p)


def _block(data: bytes, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, mem=8) -> bytes:
    co = zlib.compressobj(level, zlib.DEFLATED, -15, mem, strategy)
    comp = co.compress(data) + co.flush()
    return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", len(comp) + 25) + comp
            + struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data)))


def test_bgzf_inflate_kernel_matches_zlib(ctx, sample):
    """the two-phase decoder (bgzf_team_decode_k + bgzf_resolve_k, wgbs_bgzf_inflate): stored / fixed / dynamic blocks, long codes,
    overlapping matches, several deflate blocks per BGZF block, empty blocks, incompressible bytes -- output == zlib's, byte
    for byte; a corrupt block is reported with its number"""
    _inflate_checks(ctx, sample[1])
    from wgbs_tools_b200.patio import bgzf_compress
    bad = bytearray(bgzf_compress(sample[1][:200_000])); bad[18 + 100] ^= 0x55          # inside the first block's deflate payload
    with pytest.raises(Exception, match="inflate failed in BGZF block 0"):
        ctx.bgzf_inflate(bytes(bad)).free()
    bad = bytearray(bgzf_compress(sample[1][:200_000])); bad[-28 - 6] ^= 0x01           # CRC32 field of the last data block
    with pytest.raises(Exception, match="CRC32 mismatch"):
        ctx.bgzf_inflate(bytes(bad)).free()


def test_bgzf_inflate_deep_codes_and_the_fallback_decoder(ctx):
    """hand-written deflate blocks with deep literal codes (test_bamdev_core.deep_code_block): large second-level tables in
    bgzf_team_decode_k, and blocks whose tables exceed the arena (handed to bgzf_warp_inflate_k inside the same call) -- all == zlib"""
    from test_bamdev_core import deep_code_block
    from wgbs_tools_b200.patio import BGZF_EOF
    rng = np.random.default_rng(11)
    parts, plain = [], []
    wide = list(range(1, 8)) + [15] * 244 + [14] * 6               # 8 second-level tables of 32 entries: more than the arena holds
    for k in range(70):
        data = rng.integers(0, 256, int(rng.integers(1, 4000)), dtype=np.uint8).tobytes()
        parts.append(deep_code_block(rng, data, maxbits=int(rng.integers(11, 16)), lens=wide if k % 9 == 4 else None)); plain.append(data)
    out = ctx.bgzf_inflate(b"".join(parts) + BGZF_EOF)
    assert out.to_host().tobytes() == b"".join(plain)
    out.free()


@pytest.mark.parametrize("name,n_lines", (("Lung_STL002.small.bam", 2793), ("Pancreas_STL002.small.bam", 814)))
def test_tutorial_bams_on_the_device(ctx, bamio, name, n_lines):
    """BAMs written by htslib (the reference's tutorial data): inflated and viewed in HBM == the host reader, and the line counts the
    reference's tutorial prints for them (tutorial/README.md:75,80)"""
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name)
    bf = bamio.BamFile(path, 2)
    want = bf.view("chr3", mapq=10, exclude_flags=1796, beg=119527929, end=119531943)
    bf.close()
    with bamio.DeviceBam(ctx, path) as db:
        got = db.view("chr3", mapq=10, exclude_flags=1796, beg=119527929, end=119531943)
    assert got == want and got.count(b"\n") == n_lines


def test_bgzf_inflate_round1_decoder_matches_zlib(ctx, sample, monkeypatch):
    """bgzf_inflate_k (WGBS_INFLATE=2: one warp per block, the round-1 decoder kept as the yardstick of the bench): same checks"""
    monkeypatch.setenv("WGBS_INFLATE", "2")
    _inflate_checks(ctx, sample[1])


def _inflate_checks(ctx, s):
    from wgbs_tools_b200.patio import BGZF_EOF, bgzf_compress
    rng = np.random.default_rng(1)
    datas = [s[:60000], s[60000:125000], b"", b"a", b"ab" * 30000, b"\0" * 65000, rng.integers(0, 256, 50000, dtype=np.uint8).tobytes(),
             rng.integers(0, 4, 65000, dtype=np.uint8).tobytes(), bytes(range(256)) * 200, s[200000:260000]]
    parts, plain = [], []
    for d in datas:
        for lvl in (0, 1, 6, 9):
            if lvl == 0 and len(d) > 65000:
                continue
            parts.append(_block(d, lvl)); plain.append(d)
        for lvl, st, mem in ((6, zlib.Z_FIXED, 8), (6, zlib.Z_HUFFMAN_ONLY, 8), (6, zlib.Z_RLE, 8), (9, zlib.Z_DEFAULT_STRATEGY, 1)):
            parts.append(_block(d, lvl, st, mem)); plain.append(d)
    for k, (blk, d) in enumerate(zip(parts, plain)):                   # one at a time first: a failure names the block kind
        out = ctx.bgzf_inflate(blk)
        got = out.to_host().tobytes() if out.nbytes else b""
        out.free()
        assert got == d, f"block {k}: {len(d)} bytes"
    out = ctx.bgzf_inflate(b"".join(parts) + BGZF_EOF)
    assert out.to_host().tobytes() == b"".join(plain)
    out.free()
    big = s * 3                                                        # a few thousand blocks in one launch
    out = ctx.bgzf_inflate(bgzf_compress(big))
    assert out.nbytes == len(big) and out.to_host().tobytes() == big
    out.free()
    out = ctx.bgzf_inflate(BGZF_EOF)
    assert out.nbytes == 0
    out.free()


def test_pat2beta_cli_device_inflate(ctx, tmp_path):
    """X.pat.gz (BGZF) inflated in HBM and parsed there: same .beta as the host gunzip path"""
    from wgbs_tools_b200 import pat2beta as p2b
    from wgbs_tools_b200.patio import bgzf_compress
    N = 50_000
    idx, pats, cnt = synth.make_pat_records(1, 40_000, N)
    txt = synth.pat_text("chr1", idx, pats, cnt)
    (tmp_path / "a").mkdir(); (tmp_path / "b").mkdir()
    pg = tmp_path / "x.pat.gz"; pg.write_bytes(bgzf_compress(txt))
    p2b.pat2beta(ctx, str(pg), str(tmp_path / "a"), N, decode="host")
    p2b.pat2beta(ctx, str(pg), str(tmp_path / "b"), N, decode="device")
    a = (tmp_path / "a" / "x.beta").read_bytes(); b = (tmp_path / "b" / "x.beta").read_bytes()
    assert len(a) == 2 * N and a == b and any(a)


def test_device_views_equal_host_views(ctx, bamio, sample):
    g, full, path = sample
    rng = np.random.default_rng(5)
    starts = np.sort(rng.integers(0, g.length - 2000, 40)); ends = starts + rng.integers(1, 1500, 40)
    keep = np.concatenate([[True], starts[1:] >= ends[:-1]]); starts, ends = starts[keep], ends[keep]
    cases = [
        dict(chrom=None),
        dict(chrom="chrT"),
        dict(chrom="chrT", mapq=10, exclude_flags=1796, include_flags=3),
        dict(chrom="chrT", beg=20_000, end=21_000),
        dict(chrom="chrT", flag_eq=(99, 147)),
        dict(chrom="chrT", read_group="grp1"),
        dict(chrom="chrT", intervals=(starts, ends)),
        dict(chrom="chrT", intervals=(starts, ends), exclude_intervals=True, mapq=5),
        dict(chrom=None, max_records=1),
        dict(chrom=None, max_records=200),
        dict(chrom="chrT", mapq=10, max_records=77),
        dict(chrom="chrM"),
        dict(chrom=None, read_group="nosuch"),
    ]
    with bamio.BamFile(path, threads=4) as hb, bamio.DeviceBam(ctx, path) as db:
        assert db.refs == hb.refs and db.header == hb.header
        assert db.nrecords() == hb.nrecords() == full.count(b"\n")
        assert [db.nrecords(c) for c in db.refs] == [hb.nrecords(c) for c in hb.refs]
        assert db.view() == full
        for kw in cases:
            assert db.view(**kw) == hb.view(**kw), kw
    # from bytes already in memory (what bench.py times): same object
    with bamio.DeviceBam.from_bytes(ctx, open(path, "rb").read()) as db:
        assert db.view("chrT", mapq=10, exclude_flags=1796, include_flags=3) == b"".join(
            l + b"\n" for l in full.splitlines() if l.split(b"\t")[2] == b"chrT" and int(l.split(b"\t")[4]) >= 10 and int(l.split(b"\t")[1]) & 3 == 3
            and not int(l.split(b"\t")[1]) & 1796)


def test_every_deflate_block_type_and_records_across_blocks(ctx, bamio, tmp_path):
    """the BAM stream cut into BGZF blocks of odd sizes, compressed as stored / fixed / dynamic / RLE / huffman-only blocks:
    records straddle block boundaries everywhere"""
    from wgbs_tools_b200.patio import BGZF_EOF
    g = synth.make_genome(9, "chrT", 200_000)
    sam = synth.make_sam(g, 6000, 5, paired=True)
    raw = b"".join(zlib.decompress(b, 31) for b in _members(bamio.sam_to_bam(sam, [("chrT", g.length)])))
    rng = np.random.default_rng(2)
    parts = []; off = 0; k = 0
    while off < len(raw):
        n = int(rng.integers(1, 60_000)); d = raw[off:off + n]; off += n
        level, strat = [(0, zlib.Z_DEFAULT_STRATEGY), (1, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_FIXED), (9, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_RLE), (6, zlib.Z_HUFFMAN_ONLY)][k % 6]; k += 1
        co = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strat)
        comp = co.compress(d) + co.flush()
        parts.append(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", len(comp) + 25) + comp + struct.pack("<II", zlib.crc32(d), len(d)))
        if k % 5 == 0:
            parts.append(BGZF_EOF)                                    # empty blocks in the middle of the file are legal
    data = b"".join(parts) + BGZF_EOF
    with bamio.DeviceBam.from_bytes(ctx, data) as db:
        assert db.inflated_bytes == len(raw)
        assert db.view() == sam


def _members(bgzf: bytes):
    off = 0
    while off < len(bgzf):
        bs = struct.unpack_from("<H", bgzf, off + 16)[0] + 1
        yield bgzf[off:off + bs]
        off += bs


def test_long_records(ctx, bamio, tmp_path):
    """ONT-sized records: tens of KB each, ML arrays of thousands of values, float tags (printed like printf %g)"""
    rng = np.random.default_rng(11)
    lines = []; pos = 100
    for i in range(60):
        n = int(rng.integers(200, 90_000))
        seq = bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), n))
        qual = bytes(rng.integers(33, 74, n, dtype=np.uint8))
        nm = int(seq.count(b"C") // 3)
        mm = b"MM:Z:C+m?" + b"".join(b",%d" % int(x) for x in rng.integers(0, 3, nm)) + b";"
        ml = b"ML:B:C" + b"".join(b",%d" % int(x) for x in rng.integers(0, 256, nm))
        lines.append(b"read%d\t%d\tchrT\t%d\t60\t%dM\t*\t0\t0\t%s\t%s\t%s\t%s\tqs:f:%s\n" % (i, 16 * (i & 1), pos, n, seq, qual, mm, ml, repr(float(np.float32(rng.random() * 40))).encode()))
        pos += int(rng.integers(1, 3000))
    p = tmp_path / "ont.bam"
    p.write_bytes(bamio.sam_to_bam(b"".join(lines), [("chrT", 10_000_000)]))
    with bamio.BamFile(str(p), threads=2) as hb, bamio.DeviceBam(ctx, str(p)) as db:
        assert db.view() == hb.view()
        assert db.view("chrT", exclude_flags=16) == hb.view("chrT", exclude_flags=16)


def test_corrupt_inputs_fail_loudly(ctx, bamio, sample, tmp_path):
    from wgbs_tools_b200._lib import WgbsError
    _, _, path = sample
    good = open(path, "rb").read()
    with pytest.raises(WgbsError, match="not a BGZF"):
        bamio.DeviceBam.from_bytes(ctx, b"not a bam file at all........................")
    bs0 = struct.unpack_from("<H", good, 16)[0] + 1                    # first BGZF block: wrong ISIZE
    bad = bytearray(good); bad[bs0 - 4:bs0] = struct.pack("<I", struct.unpack_from("<I", good, bs0 - 4)[0] + 1)
    with pytest.raises(WgbsError, match="inflate failed in BGZF block 0"):
        bamio.DeviceBam.from_bytes(ctx, bytes(bad))
    bad = bytearray(good); bad[bs0 + 18] |= 0x07                        # second block: BFINAL=1, BTYPE=3 (reserved)
    with pytest.raises(WgbsError, match="inflate failed in BGZF block 1 .invalid deflate block type"):
        bamio.DeviceBam.from_bytes(ctx, bytes(bad))
    bad = bytearray(good); bad[bs0 - 8] ^= 1                            # first block: CRC32 of the trailer
    with pytest.raises(WgbsError, match="inflate failed in BGZF block 0 .CRC32 mismatch"):
        bamio.DeviceBam.from_bytes(ctx, bytes(bad))
    with pytest.raises(WgbsError, match="trailing bytes|corrupt BGZF"):
        bamio.DeviceBam.from_bytes(ctx, good[:-40])
    sam = b"a\t0\tchr2\t5\t60\t4M\t*\t0\t0\tACGT\t*\nb\t0\tchr1\t5\t60\t4M\t*\t0\t0\tACGT\t*\nc\t0\tchr2\t9\t60\t4M\t*\t0\t0\tACGT\t*\n"
    with pytest.raises(WgbsError, match="not sorted"):
        bamio.DeviceBam.from_bytes(ctx, bamio.sam_to_bam(sam, [("chr1", 100), ("chr2", 100)]))
    with pytest.raises(WgbsError, match="not a BAM"):
        from wgbs_tools_b200.patio import bgzf_compress
        bamio.DeviceBam.from_bytes(ctx, bgzf_compress(b"chr1\t5\tCC\t1\n" * 100))
    with pytest.raises(WgbsError, match="cannot open"):
        bamio.DeviceBam(ctx, str(tmp_path / "missing.bam"))
    # the context is still usable after every failure
    with bamio.DeviceBam(ctx, path) as db:
        assert db.nrecords() > 0


def test_pileup_from_device_text_equals_pileup_from_host_text(ctx, bamio, oracle, tmp_path):
    """compressed BAM -> device inflate -> device view -> pileup (nothing but the compressed bytes crosses PCIe) gives the
    pat text of the oracle pipeline on the same filtered SAM"""
    H = oracle
    g = synth.make_genome(41, "chr1", 500_000)
    sam = synth.make_sam(g, 20_000, 9, paired=True)
    lines = sam.splitlines(keepends=True)
    for i in range(0, len(lines), 37):
        t = lines[i].split(b"\t"); t[4] = b"3"; lines[i] = b"\t".join(t)
    sam = b"".join(lines)
    kept = b"".join(l for l in lines if int(l.split(b"\t")[4]) >= 10 and not int(l.split(b"\t")[1]) & 1796 and int(l.split(b"\t")[1]) & 3 == 3)
    ix = ctx.load_index(g.loci, 1)
    with bamio.DeviceBam.from_bytes(ctx, bamio.sam_to_bam(sam, [("chr1", g.length)])) as db:
        d = db.view_dev("chr1", mapq=10, exclude_flags=1796, include_flags=3)
        assert len(d) == len(kept)
        P, st = ctx.pileup_sam(ix, d)
        d.free()
    P.collapse()
    txt = P.to_text("chr1"); P.free(); ix.free()
    pout, pst = H.port_patter(H.port_match_maker(kept), g.loci, g.idx())
    assert txt == H.port_collapse(pout)
    assert [st[k] for k in ("lines", "pairs", "empty", "short", "invalid", "paired")] == pst


def test_bam2pat_cli_device_decode_equals_host_decode(ctx, bamio, tmp_path):
    from wgbs_tools_b200 import bam2pat
    g1 = synth.make_genome(31, "chr1", 400_000, first_idx=1)
    g2 = synth.make_genome(32, "chr2", 300_000, first_idx=1 + g1.n_cpg)
    refdir = tmp_path / "ref"; refdir.mkdir()
    with gzip.open(refdir / "CpG.bed.gz", "wb") as f:
        f.write(g1.dict_text() + g2.dict_text())
    (refdir / "CpG.chrome.size").write_text(f"chr1\t{g1.n_cpg}\nchr2\t{g2.n_cpg}\n")
    (refdir / "chrome.size").write_text(f"chr1\t{g1.length}\nchr2\t{g2.length}\n")
    sam = synth.make_sam(g1, 9000, 1, paired=True, name_prefix="a") + synth.make_sam(g2, 7000, 2, paired=True, name_prefix="b")
    bam = tmp_path / "s.bam"; bam.write_bytes(bamio.sam_to_bam(sam, [("chr1", g1.length), ("chr2", g2.length)]))
    for argv in ([], ["-r", "chr1:100000-180000", "--top_strand"], ["--long", "--no_beta"], ["--mbias"]):
        outs = []
        for mode in ("host", "device"):
            out = tmp_path / f"out_{mode}_{len(outs)}_{abs(hash(tuple(argv)))}"; out.mkdir()
            bam2pat.main([str(bam), "--genome", str(refdir), "-o", str(out), "--bam_decode", mode] + argv)
            outs.append(out)
        a, b = outs
        assert gzip.decompress((a / "s.pat.gz").read_bytes()) == gzip.decompress((b / "s.pat.gz").read_bytes()), argv
        assert len(gzip.decompress((a / "s.pat.gz").read_bytes())) > 1000
        if "--no_beta" not in argv:
            assert (a / "s.beta").read_bytes() == (b / "s.beta").read_bytes(), argv
        if "--mbias" in argv:
            for x in ("OT", "OB"):
                assert (a / "s.mbias" / f"s.mbias.{x}.txt").read_bytes() == (b / "s.mbias" / f"s.mbias.{x}.txt").read_bytes()


# ---- the direct route: BAM records feed the pileup kernels as they are (no SAM text formatted / tokenized) --------------------------
def _both_routes(ctx, db, ix, chrom, monkeypatch, **kw):
    res = []
    for direct in ("0", "1"):
        monkeypatch.setenv("WGBS_DBAM_DIRECT", direct)
        P, st = db.pileup(ix, chrom, **kw)
        long = bool(kw.get("keep_names"))
        raw = P.to_text(chrom, long=long) if not long else None
        P.collapse(long=long)
        txt = P.to_text(chrom, long=long)
        P.free()
        mb = st.pop("mbias", None)
        res.append((raw, txt, st, None if mb is None else mb.tobytes()))
    assert res[0] == res[1], kw
    return res[1]


def test_direct_route_equals_text_route_and_oracle(ctx, bamio, oracle, monkeypatch):
    H = oracle
    g = synth.make_genome(7, "chrT", 1_000_000)
    ix = ctx.load_index(g.loci, g.first_idx)
    refs = [("chrT", g.length)]
    # 1. paired-end reads with I/D/S events and singletons, plus supplementary copies (3-record QNAME groups: whole-line ordering)
    lines = synth.make_sam(g, 6_000, 5, paired=True, single_frac=0.05).splitlines()
    extra = []
    for l in lines[::97]:
        t = l.split(b"\t"); t[1] = b"%d" % (int(t[1]) | 2048); t[3] = b"%d" % (int(t[3]) + 5000); extra.append(b"\t".join(t))
    sam = b"\n".join(sorted(lines + extra, key=lambda l: int(l.split(b"\t")[3]))) + b"\n"
    with bamio.DeviceBam.from_bytes(ctx, bamio.sam_to_bam(sam, refs)) as db:
        for kw in (dict(), dict(clip=7, min_cpg=2), dict(keep_names=True), dict(mbias=True), dict(view=dict(mapq=10, exclude_flags=1796, include_flags=3)),
                   dict(view=dict(beg=200_000, end=400_000)), dict(view=dict(max_records=1001))):
            raw, txt, st, _ = _both_routes(ctx, db, ix, "chrT", monkeypatch, **kw)
            if not kw:
                pout, pst = H.port_patter(H.port_match_maker(sam), g.loci, g.idx())
                assert txt == H.port_collapse(pout)
                assert [st[k] for k in ("lines", "pairs", "empty", "short", "invalid", "paired")] == pst and st["pairs"] > 2000
    # 2. single-end reads with every CIGAR op, invalid CIGARs, SEQ '*', CIGAR '*', reads beyond the last CpG
    rng = np.random.default_rng(4)
    recs = []; pos = 1000
    for i in range(3000):
        pos += int(rng.integers(1, 300))
        ops = []; qlen = 0
        for _ in range(int(rng.integers(1, 7))):
            op = "MIDSNH=XP"[int(rng.integers(0, 9))]; k = int(rng.integers(1, 40))
            ops.append(f"{k}{op}")
            if op in "MIS=X":
                qlen += k
        seq = bytes(rng.choice(list(b"ACGTN"), size=max(qlen + int(rng.integers(-2, 3)), 1), p=[.2, .3, .25, .2, .05]).tolist())
        cig = "".join(ops).encode() if i % 50 else b"*"
        recs.append(b"r%d\t%d\tchrT\t%d\t60\t%s\t*\t0\t0\t%s\t*" % (i, 16 * int(rng.integers(0, 2)), pos, cig, seq if i % 41 else b"*"))
    p = int(g.loci[-1]) - 5
    recs.append(b"z\t0\tchrT\t%d\t60\t60M\t*\t0\t0\t%s\t*" % (p, g.bases[p:p + 60].tobytes()))
    sam = b"\n".join(recs) + b"\n"
    with bamio.DeviceBam.from_bytes(ctx, bamio.sam_to_bam(sam, refs)) as db:
        raw, txt, st, _ = _both_routes(ctx, db, ix, "chrT", monkeypatch)
        pout, pst = H.port_patter(sam, g.loci, g.idx())
        assert txt == H.port_collapse(pout)
        assert [st[k] for k in ("lines", "pairs", "empty", "short", "invalid", "paired")] == pst and st["invalid"] > 100
        _both_routes(ctx, db, ix, "chrT", monkeypatch, clip=3, mbias=True)
    # 3. MM/ML data takes the text route by itself (first record carries an MM tag): same result under either setting
    sam = synth.make_sam(g, 800, 9, paired=False, np_mode=True) if "np_mode" in synth.make_sam.__code__.co_varnames else None
    if sam:
        with bamio.DeviceBam.from_bytes(ctx, bamio.sam_to_bam(sam, refs)) as db:
            raw, txt, st, _ = _both_routes(ctx, db, ix, "chrT", monkeypatch)
            assert st["nanopore"] == 1 and txt
    ix.free()


def test_bam2pat_cli_direct_route(ctx, bamio, tmp_path, monkeypatch):
    from wgbs_tools_b200 import bam2pat
    g1 = synth.make_genome(31, "chr1", 400_000, first_idx=1)
    refdir = tmp_path / "ref"; refdir.mkdir()
    with gzip.open(refdir / "CpG.bed.gz", "wb") as f:
        f.write(g1.dict_text())
    (refdir / "CpG.chrome.size").write_text(f"chr1\t{g1.n_cpg}\n"); (refdir / "chrome.size").write_text(f"chr1\t{g1.length}\n")
    sam = synth.make_sam(g1, 9000, 1, paired=True)
    bam = tmp_path / "s.bam"; bam.write_bytes(bamio.sam_to_bam(sam, [("chr1", g1.length)]))
    outs = {}
    for mode, direct in (("host", "0"), ("device", "0"), ("device", "1")):
        monkeypatch.setenv("WGBS_DBAM_DIRECT", direct)
        out = tmp_path / f"out_{mode}_{direct}"; out.mkdir()
        bam2pat.main([str(bam), "--genome", str(refdir), "-o", str(out), "--bam_decode", mode, "--mbias"])
        outs[(mode, direct)] = (gzip.decompress((out / "s.pat.gz").read_bytes()), (out / "s.beta").read_bytes(),
                                (out / "s.mbias" / "s.mbias.OT.txt").read_bytes())
    assert outs[("host", "0")] == outs[("device", "0")] == outs[("device", "1")]
    assert len(outs[("host", "0")][0]) > 1000


@pytest.mark.parametrize("decode,direct", [("host", "0"), ("device", "0"), ("device", "1")])
def test_bam2pat_cli_template_windows(ctx, bamio, tmp_path, monkeypatch, decode, direct):
    """a chromosome piled up in template windows (wgbs_view_opts.key_beg / key_end; what bam2pat does when a chromosome is too
    large for one call) == the chromosome in one call: same .pat.gz text, same .beta, for every decoder / route"""
    from wgbs_tools_b200 import bam2pat
    g1 = synth.make_genome(31, "chr1", 400_000, first_idx=1)
    refdir = tmp_path / "ref"; refdir.mkdir()
    with gzip.open(refdir / "CpG.bed.gz", "wb") as f:
        f.write(g1.dict_text())
    (refdir / "CpG.chrome.size").write_text(f"chr1\t{g1.n_cpg}\n"); (refdir / "chrome.size").write_text(f"chr1\t{g1.length}\n")
    sam = synth.make_sam(g1, 12_000, 4, paired=True, single_frac=0.03)
    bam = tmp_path / "s.bam"; bam.write_bytes(bamio.sam_to_bam(sam, [("chr1", g1.length)]))
    monkeypatch.setenv("WGBS_DBAM_DIRECT", direct)
    outs = []
    for tag, limit in (("whole", "0"), ("windows", "1700")):
        monkeypatch.setenv("WGBS_CHUNK_RECORDS", limit)
        out = tmp_path / tag; out.mkdir()
        bam2pat.main([str(bam), "--genome", str(refdir), "-o", str(out), "--bam_decode", decode])
        outs.append((gzip.decompress((out / "s.pat.gz").read_bytes()), (out / "s.beta").read_bytes()))
    assert outs[0] == outs[1] and len(outs[0][0]) > 1000


def test_device_parts_equal_host_parts_and_streamed_cli(ctx, bamio, tmp_path, monkeypatch):
    """wgbs_dbam_open_part / last_record / first_key: a file streamed as parts decoded on the device gives the same items (same
    views, same tails) as the host part reader, and `bam2pat --bam_decode stream` with WGBS_STREAM_BACKEND=device the same files
    as decoding the file as a whole"""
    from wgbs_tools_b200 import bam2pat
    g1 = synth.make_genome(31, "chr1", 300_000, first_idx=1); g2 = synth.make_genome(32, "chr2", 150_000, first_idx=1 + g1.n_cpg)
    sam = synth.make_sam(g1, 9_000, 4, paired=True, single_frac=0.03, name_prefix="a") + synth.make_sam(g2, 4_000, 5, paired=True, name_prefix="b")
    bam = tmp_path / "s.bam"; bam.write_bytes(bamio.sam_to_bam(sam, [("chr1", g1.length), ("chr2", g2.length)]))
    kw = dict(mapq=10, exclude_flags=1796, include_flags=3)
    items = {}
    for name, opener in (("host", lambda d, r, l, f: bamio.BamPart(d, r, l, f, threads=2)), ("device", lambda d, r, l, f: bamio.DeviceBamPart(ctx, d, r, l, f))):
        items[name] = [(chrom, win, done, part.tail, part.view(chrom, key_window=win, **kw)) for part, chrom, win, done in bamio.stream_parts(str(bam), opener, lambda c: kw, 200_000)]
    assert items["host"] == items["device"] and len(items["host"]) > 6
    refdir = tmp_path / "ref"; refdir.mkdir()
    with gzip.open(refdir / "CpG.bed.gz", "wb") as f:
        f.write(g1.dict_text() + g2.dict_text())
    (refdir / "CpG.chrome.size").write_text(f"chr1\t{g1.n_cpg}\nchr2\t{g2.n_cpg}\n"); (refdir / "chrome.size").write_text(f"chr1\t{g1.length}\nchr2\t{g2.length}\n")
    outs = []
    for tag, decode, backend in (("whole", "host", "host"), ("stream_host", "stream", "host"), ("stream_dev", "stream", "device")):
        monkeypatch.setenv("WGBS_STREAM_BACKEND", backend); monkeypatch.setenv("WGBS_STREAM_BYTES", "200000")
        out = tmp_path / tag; out.mkdir()
        bam2pat.main([str(bam), "--genome", str(refdir), "-o", str(out), "--bam_decode", decode])
        outs.append((gzip.decompress((out / "s.pat.gz").read_bytes()), (out / "s.beta").read_bytes()))
    assert outs[0] == outs[1] == outs[2] and len(outs[0][0]) > 1000


@pytest.mark.parametrize("decode", ["host", "device"])
def test_bam2pat_cli_chromosomes_in_flight(ctx, bamio, tmp_path, monkeypatch, decode):
    """--gpu_streams 2: two chromosomes at a time, each on its own Context (stream) and host thread, all adding into the one
    device beta array; a device-decoded file is shared by the Contexts.  Same files as one chromosome after the other."""
    from wgbs_tools_b200 import bam2pat
    gs = [synth.make_genome(31 + k, f"chr{k + 1}", 250_000, first_idx=1) for k in range(3)]
    first = 1
    for g in gs:
        g.first_idx = first; first += g.n_cpg
    refdir = tmp_path / "ref"; refdir.mkdir()
    with gzip.open(refdir / "CpG.bed.gz", "wb") as f:
        f.write(b"".join(g.dict_text() for g in gs))
    (refdir / "CpG.chrome.size").write_text("".join(f"{g.chrom}\t{g.n_cpg}\n" for g in gs)); (refdir / "chrome.size").write_text("".join(f"{g.chrom}\t{g.length}\n" for g in gs))
    sam = b"".join(synth.make_sam(g, 6_000, 4 + k, paired=True, name_prefix=f"c{k}_") for k, g in enumerate(gs))
    bam = tmp_path / "s.bam"; bam.write_bytes(bamio.sam_to_bam(sam, [(g.chrom, g.length) for g in gs]))
    outs = []
    for tag, extra in (("one", []), ("two", ["--gpu_streams", "2"]), ("three", ["--gpu_streams", "3", "--mbias"])):
        out = tmp_path / tag; out.mkdir()
        bam2pat.main([str(bam), "--genome", str(refdir), "-o", str(out), "--bam_decode", decode] + extra)
        outs.append((gzip.decompress((out / "s.pat.gz").read_bytes()), (out / "s.beta").read_bytes()))
    assert outs[0] == outs[1] == outs[2] and len(outs[0][0]) > 3000

