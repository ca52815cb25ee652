"""GPU parity of the bam->pat pileup (tokenizer, pairing, CIGAR/CpG calls, mate merge, collapse, text) through the C ABI
against the reference pipeline  `[match_maker |] patter | sort -k2,2n -k3,3 | uniq -c | awk`  (oracle/_ref) and the C
restatement."""
import numpy as np
import pytest

from wgbs_tools_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def genome():
    return synth.make_genome(7, "chrT", 1_000_000)


def _oracle_pat(H, g, sam, paired, **kw):
    """(raw patter lines sorted, collapsed text, stats) from the reference executables, else from the C port."""
    if H.have_ref():
        d = H.write_tmp(g.dict_text(), ".CpG.bed")
        out, err = H.ref_patter(sam, d, g.chrom, paired, **kw)
        return out, H.ref_collapse(out)
    mm = H.port_match_maker(sam) if paired else sam
    out, _ = H.port_patter(mm, g.loci, g.idx(), **kw)
    return out, H.port_collapse(out)


def _gpu_pat(ctx, g, sam, **kw):
    ix = ctx.load_index(g.loci, g.first_idx)
    P, st = ctx.pileup_sam(ix, sam, **kw)
    raw = P.to_text(g.chrom)
    P.collapse()
    txt = P.to_text(g.chrom)
    P.free(); ix.free()
    return raw, txt, st


def test_sort_pairs_is_a_stable_radix_sort(ctx):
    rng = np.random.default_rng(0)
    for n, hi in [(1, 10), (1000, 4), (70_001, 2**32), (300_000, 2**20)]:
        k = rng.integers(0, hi, size=n, dtype=np.uint64).astype(np.uint32); v = np.arange(n, dtype=np.uint32)
        ko, vo = ctx.sort_pairs(k, v)
        order = np.argsort(k, kind="stable")
        np.testing.assert_array_equal(ko, k[order]); np.testing.assert_array_equal(vo, v[order])


@pytest.mark.parametrize("paired", [True, False])
@pytest.mark.parametrize("clip,min_cpg", [(0, 1), (5, 2)])
def test_pileup_matches_reference_pipeline(ctx, oracle, genome, paired, clip, min_cpg):
    H = oracle
    sam = synth.make_sam(genome, 20_000, 21 + clip, paired=paired)
    ref_raw, ref_txt = _oracle_pat(H, genome, sam, paired, min_cpg=min_cpg, clip=clip)
    raw, txt, st = _gpu_pat(ctx, genome, sam, min_cpg=min_cpg, clip=clip)
    # templates before the collapse: same multiset of lines (the reference's own order depends on an unstable sort)
    assert sorted(l + b"\t1" for l in ref_raw.splitlines()) == sorted(raw.splitlines())
    assert txt == ref_txt                                   # the uncompressed .pat bytes
    mm = H.port_match_maker(sam) if paired else sam
    _, pst = H.port_patter(mm, genome.loci, genome.idx(), min_cpg=min_cpg, clip=clip)
    assert [st[k] for k in ("lines", "pairs", "empty", "short", "invalid", "paired")] == pst
    assert len(txt) > 10_000


def test_pileup_invalid_and_odd_reads(ctx, oracle, genome):
    H = oracle; g = genome
    p = int(g.loci[100]) - 20
    seq = g.bases[p:p + 60].tobytes()
    lines = [
        b"a\t0\tchrT\t%d\t60\t60M\t*\t0\t0\t%s\t*" % (p, seq),
        b"b\t0\tchrT\t%d\t60\t*\t*\t0\t0\t%s\t*" % (p, seq),
        b"c\t0\tchrT\t%d\t60\t70M\t*\t0\t0\t%s\t*" % (p, seq),
        b"d\t0\tchrT\t%d\t60\t30M2P30M\t*\t0\t0\t%s\t*" % (p, seq),
        b"e\t16\tchrT\t%d\t60\t10S40M10H\t*\t0\t0\t%s\t*" % (p, seq),
        b"f\t0\tchrT\t%d\t60\t20M5N35M5S\t*\t0\t0\t%s\t*" % (p, seq),
        b"g\t0\tchrT\t%d\t60\t20=5X35M\t*\t0\t0\t%s\t*" % (p, seq),
        b"h\t0\tchrT",
        b"i\t0\tchrT\t%d\t60\t60M\t*\t0\t0\t*\t*" % p,
        b"j\t0\tchrT\t%d\t60\t60M\t*\t0\t0\t%s\t*" % (int(g.loci[-1]) - 5, seq),
        b"k\t0\tchrT\t%d\t60\t5M\t*\t0\t0\t%s\t*" % (p, seq),
        b"l\t16\tchrT\t%d\t60\t25M3I10M4D22M\t*\t0\t0\t%s\t*" % (p, seq),
        b"m\t0\tchrT\t%d\t60\tM\t*\t0\t0\t%s\t*" % (p, seq),                   # op without a number
        b"n\t0\tchrT\t%d\t60\t99999999999M\t*\t0\t0\t%s\t*" % (p, seq),        # int overflow
        b"o\tx\tchrT\t%d\t60\t60M\t*\t0\t0\t%s\t*" % (p, seq),                 # non numeric flag
        b"",
        b"q\t16\tchrT\t%d\t60\t60M\t*\t0\t0\t%s\t*\tNM:i:0\tXX:Z:abc" % (p + 1, seq),
    ]
    sam = b"\n".join(lines) + b"\n"
    ref_raw, ref_txt = _oracle_pat(H, g, sam, False)
    raw, txt, st = _gpu_pat(ctx, g, sam)
    assert txt == ref_txt
    _, pst = H.port_patter(sam, g.loci, g.idx())
    assert [st[k] for k in ("lines", "pairs", "empty", "short", "invalid", "paired")] == pst
    assert st["invalid"] == 8


def test_pileup_no_trailing_newline_and_single_line(ctx, oracle, genome):
    H = oracle; g = genome
    sam = synth.make_sam(g, 50, 3, paired=False).rstrip(b"\n")
    ref_raw, ref_txt = _oracle_pat(H, g, sam + b"\n", False)
    raw, txt, st = _gpu_pat(ctx, g, sam)
    assert txt == ref_txt and st["lines"] == 50
    raw, txt, st = _gpu_pat(ctx, g, b"")
    assert txt == b"" and st["lines"] == 0


def test_pairing_groups_of_three_and_far_mates(ctx, oracle, genome):
    """supplementary alignments share a QNAME (FLAGS_FILTER 1796 keeps 0x800, bam2pat.py:27): greedy pairing in
    whole-line order; a far-away mate still pairs; a dropped mate leaves a single."""
    H = oracle; g = genome
    base = synth.make_sam(g, 4_000, 5, paired=True, single_frac=0.05)
    lines = base.splitlines()
    # add supplementary copies (flag |= 2048, other position) of some read1 records, keeping coordinate order
    extra = []
    for l in lines[::97]:
        t = l.split(b"\t")
        t[1] = b"%d" % (int(t[1]) | 2048)
        t[3] = b"%d" % (int(t[3]) + 5000)
        extra.append(b"\t".join(t))
    allr = sorted(lines + extra, key=lambda l: int(l.split(b"\t")[3]))
    sam = b"\n".join(allr) + b"\n"
    ref_raw, ref_txt = _oracle_pat(H, g, sam, True)
    raw, txt, st = _gpu_pat(ctx, g, sam)
    assert txt == ref_txt
    assert st["pairs"] > 1500


def test_long_patterns_multiword_sort(ctx, oracle):
    """dense CpG island: >16 and >32 symbols per read (multi-word pool records, multi-pass collapse)"""
    H = oracle
    L = 60_000
    loci = np.arange(1001, 50_000, 3, dtype=np.int64)             # CpG every 3 bp
    g = synth.Genome("chrD", L, loci, 1, None, np.full(loci.size, 0.5, np.float32))
    bases = np.full(L + 2, ord("A"), np.uint8); bases[loci] = ord("C"); bases[loci + 1] = ord("G"); g.bases = bases
    sam = synth.make_sam(g, 6_000, 9, paired=True)
    ref_raw, ref_txt = _oracle_pat(H, g, sam, True)
    raw, txt, st = _gpu_pat(ctx, g, sam)
    assert max(len(l.split(b"\t")[2]) for l in txt.splitlines()) > 48
    assert txt == ref_txt


def test_tutorial_like_real_cigars(ctx, oracle, genome):
    """CIGAR diversity: many I/D/S/H/N events per read"""
    H = oracle; g = genome
    rng = np.random.default_rng(4)
    lines = []
    pos = 1000
    for i in range(3000):
        pos += int(rng.integers(1, 300))
        ops = []; qlen = 0
        for _ in range(int(rng.integers(1, 7))):
            op = "MIDSNH=X"[int(rng.integers(0, 8))]; k = int(rng.integers(1, 40))
            ops.append(f"{k}{op}")
            if op in "MIS=X":
                qlen += k
        seq = bytes(rng.choice(list(b"ACGTN"), size=max(qlen + int(rng.integers(-2, 3)), 1), p=[.2, .3, .25, .2, .05]).tolist())
        lines.append(b"r%d\t%d\tchrT\t%d\t60\t%s\t*\t0\t0\t%s\t*" % (i, 16 * int(rng.integers(0, 2)), pos, "".join(ops).encode(), seq))
    sam = b"\n".join(lines) + b"\n"
    ref_raw, ref_txt = _oracle_pat(H, g, sam, False)
    raw, txt, st = _gpu_pat(ctx, g, sam)
    assert txt == ref_txt
    _, pst = H.port_patter(sam, g.loci, g.idx())
    assert [st[k] for k in ("lines", "pairs", "empty", "short", "invalid", "paired")] == pst
    assert st["invalid"] > 100


# ---- MM/ML (modification-aware) mode -----------------------------------------------------------------------------------
@pytest.mark.parametrize("kw", [dict(), dict(combine_mods=True), dict(cpc_call="H"), dict(np_thresh=0.8), dict(clip=3), dict(min_cpg=3)])
def test_np_pileup_matches_reference(ctx, oracle, genome, kw):
    H = oracle
    sam = synth.make_np_sam(genome, 6000, 5)
    rkw = dict(kw); rkw.setdefault("np_thresh", 0.67)
    ref_raw, ref_txt = _oracle_pat(H, genome, sam, False, nanopore=True, **rkw)
    raw, txt, st = _gpu_pat(ctx, genome, sam, nanopore=True, **rkw)
    assert txt == ref_txt and len(txt) > 5000
    _, pst = H.port_patter(sam, genome.loci, genome.idx(), nanopore=True, **rkw)
    assert [st[k] for k in ("lines", "pairs", "empty", "short", "invalid", "paired")] == pst
    assert st["nanopore"] == 1


def test_np_autodetect_and_odd_tags(ctx, oracle, genome):
    """first line carries MM -> nanopore mode without the flag (patter.cpp:337-338); malformed / unusual MM, ML"""
    H = oracle; g = genome
    k = 300
    while g.loci[k + 5] - g.loci[k] > 100:
        k += 1
    p = int(g.loci[k]) - 2
    n = int(g.loci[k + 5]) - p + 3
    seq = g.bases[p:p + n].tobytes()
    comp = bytes.maketrans(b"ACGTN", b"TGCAN")

    def rec(name, flag, pos, cigar, s, tags):
        return b"%s\t%d\tchrT\t%d\t60\t%s\t*\t0\t0\t%s\t*\tNM:i:0%s" % (name, flag, pos, cigar, s, tags)
    M = b"%dM" % n
    lines = [
        rec(b"a", 0, p, M, seq, b"\tMM:Z:C+m?,0,0,0;C+h?,0,0,0;\tML:B:C,250,3,128,2,250,5"),
        rec(b"b", 16, p, M, seq, b"\tMM:Z:C+m?,0,0,0;C+h?,0,0,0;\tML:B:C,250,3,128,2,250,5"),
        rec(b"c", 0, p, M, seq, b"\tMM:Z:C+m.,1,0;\tML:B:C,255,0"),                       # dot convention: unlisted C -> T
        rec(b"d", 16, p, M, seq, b"\tMm:Z:C+m.,1,0;"),                                       # no ML -> 255
        rec(b"e", 0, p, M, seq, b"\tMM:Z:C+m?,0,0,0;\tML:B:C,250,3"),                       # ML count not a multiple -> invalid
        rec(b"f", 0, p, M, seq, b"\tMM:Z:C+h?,0,1;C+C?,0,0,0;C+m?,2;\tML:B:C,200,10,255,255,255,90"),   # section order + C+C? (ML slicing quirk)
        rec(b"g", 0, p, M, seq, b"\tMM:Z:C+C?,0,0,0;"),                                      # only C+C?
        rec(b"h", 0, p, M, seq, b"\tMM:Z:G-m?,0;"),                                          # no C+ section, '?' -> empty
        rec(b"i", 0, p, M, seq, b""),                                                        # no tags at all -> empty
        rec(b"j", 16, p, M, seq.replace(b"A", b"R", 1), b"\tMM:Z:C+m?,0;\tML:B:C,255"),     # bottom + non ACGTN -> invalid
        rec(b"k", 0, p, b"5M2D%dM3I5M" % (n - 13), seq, b"\tMM:Z:C+m.,0,0,0,0;\tML:B:C,255,1,255,1"),
        rec(b"l", 16, p + 1, b"%dM" % (n - 1), seq[1:], b"\tMM:Z:C+m?,0,0,0,0,0,0,0,0;\tML:B:C,255,255,255,255,255,255,255,255"),
        rec(b"m", 0, p, M, b"*", b"\tMM:Z:C+m?,0;\tML:B:C,255"),                             # SEQ '*' -> empty
        rec(b"n", 0, p, b"*", seq, b"\tMM:Z:C+m?,0;\tML:B:C,255"),                           # bad CIGAR -> invalid
    ]
    sam = b"\n".join(lines) + b"\n"
    for kw in (dict(), dict(combine_mods=True), dict(cpc_call="H"), dict(cpc_call=".")):
        ref_raw, ref_txt = _oracle_pat(H, g, sam, False, **kw)                               # NB: no nanopore flag
        raw, txt, st = _gpu_pat(ctx, g, sam, **kw)
        assert txt == ref_txt, kw
        _, pst = H.port_patter(sam, g.loci, g.idx(), **kw)
        assert [st[k] for k in ("lines", "pairs", "empty", "short", "invalid", "paired")] == pst, kw
        assert st["nanopore"] == 1 and st["invalid"] == 3


def test_very_long_reads_multi_tile_lines(ctx, oracle):
    """lines far longer than a tokenizer tile (64 KiB) / patterns of hundreds of words: SE long reads with many CIGAR ops"""
    H = oracle
    g = synth.make_genome(11, "chrL", 1_500_000)
    rng = np.random.default_rng(8)
    lines = []
    for i, (p, L) in enumerate([(1000, 200_000), (150_000, 90_000), (150_000, 90_000), (600_000, 300_000), (1_200_000, 70_000)]):
        seq = bytearray(g.bases[p:p + L].tobytes())
        ops = []; q = 0
        while q < L:
            m = int(min(L - q, rng.integers(50, 4000)))
            ops.append(b"%dM" % m); q += m
            if q < L and rng.random() < 0.5:
                d = int(rng.integers(1, 30)); ops.append(b"%dD" % d)          # deletion: later bases are shifted on the reference
        span_shift = sum(int(o[:-1]) for o in ops if o.endswith(b"D"))
        lines.append(b"long%d\t%d\tchrL\t%d\t60\t%s\t*\t0\t0\t%s\t*" % (i, 16 * (i % 2), p, b"".join(ops), bytes(seq)))
    sam = b"\n".join(lines) + b"\n"
    assert max(len(l) for l in lines) > 3 * 65536
    ref_raw, ref_txt = _oracle_pat(H, g, sam, False)
    raw, txt, st = _gpu_pat(ctx, g, sam)
    assert txt == ref_txt
    assert max(len(l.split(b"\t")[2]) for l in txt.splitlines()) > 1000


def test_collapse_sums_counts_of_parsed_pat(ctx):
    """wgbs_collapse on arbitrary records: (idx, pattern) order of `sort -k2,2n -k3,3`, identical records merged, counts summed"""
    idx, pats, cnt = synth.make_pat_records(12, 30_000, 600, mean_len=9, max_len=80)
    rng = np.random.default_rng(1)
    sel = rng.integers(0, idx.size, size=60_000)                   # duplicates on purpose, shuffled
    lines = [b"chr1\t%d\t%s\t%d\n" % (idx[i], pats[i], cnt[i]) for i in sel.tolist()]
    P = ctx.pats_from_text(b"".join(lines))
    P.collapse()
    got = P.to_text("chr1")
    P.free()
    agg = {}
    for i in sel.tolist():
        k = (int(idx[i]), pats[i]); agg[k] = agg.get(k, 0) + int(cnt[i])
    exp = b"".join(b"chr1\t%d\t%s\t%d\n" % (k[0], k[1], v) for k, v in sorted(agg.items()))
    assert got == exp


@pytest.mark.parametrize("seed,np_mode", [(1, False), (2, False), (5, False), (3, True), (4, True), (6, True)])
def test_fuzzed_sam_on_gpu(ctx, oracle, genome, seed, np_mode):
    """hostile SAM text: same pat text and the same empty / invalid counters as the oracle, and no crash"""
    from fuzz_sam import fuzz_sam
    H = oracle
    sam = fuzz_sam(genome, seed, 700, np_mode)
    first = (synth.make_np_sam(genome, 1, 99) if np_mode else synth.make_sam(genome, 1, 99, paired=False))
    sam = first + sam
    kw = dict(nanopore=True, np_thresh=0.67) if np_mode else {}
    raw, txt, st = _gpu_pat(ctx, genome, sam, **kw)
    pout, pst = H.port_patter(sam, genome.loci, genome.idx(), **kw)
    assert txt == H.port_collapse(pout)
    assert [st[k] for k in ("lines", "pairs", "empty", "short", "invalid", "paired")] == pst


@pytest.mark.parametrize("paired", [True, False])
def test_mbias_tables_match_patter(ctx, oracle, genome, paired, tmp_path):
    """`patter --mbias` (patter.cpp:116-165): per-read-position meth / unmeth counts, counted before the clip; reads of
    1000+ reference bases are skipped entirely on the bottom strand and beyond position 999 on the top strand"""
    H = oracle
    if not H.have_ref():
        pytest.skip("needs the reference patter executable")
    g = genome
    sam = synth.make_sam(g, 8000, 77, paired=paired)
    if not paired:
        # two long single-end reads (top and bottom) to exercise the MAX_READ_LEN guard
        for flag, p in ((0, 20_000), (16, 40_000)):
            seq = g.bases[p:p + 1500].tobytes()
            sam += b"long%d\t%d\tchrT\t%d\t60\t1500M\t*\t0\t0\t%s\t*\n" % (flag, flag, p, seq)
    else:
        # improper-pair flags (not 83/163/99/147) take no part in the M-bias tables
        lines = sam.splitlines(keepends=True)
        for i in range(0, len(lines), 41):
            t = lines[i].split(b"\t"); t[1] = b"%d" % (int(t[1]) & ~2); lines[i] = b"\t".join(t)
        sam = b"".join(lines)
    d = H.write_tmp(g.dict_text(), ".CpG.bed")
    pref = str(tmp_path / "mb")
    out, err = H.ref_patter(sam, d, g.chrom, paired, clip=4, mbias=pref)
    exp = H.read_mbias(pref)
    ix = ctx.load_index(g.loci, g.first_idx)
    P, st = ctx.pileup_sam(ix, sam, clip=4, mbias=True)
    P.collapse()
    assert P.to_text(g.chrom) == H.ref_collapse(out)
    P.free(); ix.free()
    np.testing.assert_array_equal(st["mbias"], exp)
    assert exp.sum() > 5000


@pytest.mark.parametrize("paired", [True, False])
def test_long_format_matches_reference(ctx, oracle, genome, paired):
    """--long: `patter --long | sort -k2,2n -k3,3 | awk '{print $1,$2,$3,1,$4}'` (bam2pat.py:102-103): one line per template
    with its read name, ordered by (idx, pattern, then the whole line = the read name)"""
    H = oracle
    if not H.have_ref():
        pytest.skip("needs the reference patter executable")
    # a tiny chromosome: many templates share (idx, pattern), so the order is decided by the read names ("q7" < "q70" < "q8")
    g = synth.make_genome(5, "chrT", 40_000)
    sam = synth.make_sam(g, 6000, 91, paired=paired, name_prefix="q")
    d = H.write_tmp(g.dict_text(), ".CpG.bed")
    out, _ = H.ref_patter(sam, d, g.chrom, paired, long=True)
    exp = H.ref_collapse_long(out)
    ix = ctx.load_index(g.loci, g.first_idx)
    P, st = ctx.pileup_sam(ix, sam, keep_names=True)
    P.collapse(long=True)
    got = P.to_text(g.chrom, long=True)
    # the same object still serves the beta path
    assert got == exp and got.count(b"\n") == st["templates"]
    P.free(); ix.free()


@pytest.mark.parametrize("bits", [4, 12])
def test_pairing_survives_hash_collisions(ctx, oracle, genome, monkeypatch, bits):
    """WGBS_PAIR_HASH_BITS truncates the QNAME hash (test hook): every slot overflows, so all mates are found on the
    crowded-slot path (sort by hash, whole-line order, greedy pairing by name) -- same output as without collisions"""
    H = oracle
    sam = synth.make_sam(genome, 6_000, 33, paired=True, single_frac=0.1)
    ref_raw, ref_txt = _oracle_pat(H, genome, sam, True)
    monkeypatch.setenv("WGBS_PAIR_HASH_BITS", str(bits))
    raw, txt, st = _gpu_pat(ctx, genome, sam)
    assert txt == ref_txt
    _, pst = H.port_patter(H.port_match_maker(sam), genome.loci, genome.idx())
    assert [st[k] for k in ("lines", "pairs", "empty", "short", "invalid", "paired")] == pst


def _big_batch(g, n_reads=160_000, seed=21, far_every=23, supp=True):
    """>= 3 match_maker buffers (it flushes every 50 000 lines, match_maker.cpp:162-170) of PE reads: mates 60 kb apart (thousands of
    lines: many pairs straddle a buffer edge), singles, and supplementary copies (0x800 passes -F 1796) 2 kb from their record.  PNEXT /
    POS stay consistent.  Returns the lines, coordinate-sorted."""
    lines = synth.make_sam(g, n_reads, seed, paired=True, single_frac=0.02).splitlines()
    first = {}
    for k, l in enumerate(lines):
        first.setdefault(l.split(b"\t", 1)[0], k)
    out = []
    for k, l in enumerate(lines):
        t = l.split(b"\t")
        fl, pos, pn, tl = int(t[1]), int(t[3]), int(t[7]), int(t[8])
        far = first[t[0]] % far_every == 0 and max(pos, pn) + 60_000 < g.length - 300
        if far and tl > 0:                                     # left mate: its mate moves 60 kb downstream
            t[7] = b"%d" % (pn + 60_000); t[8] = b"%d" % (tl + 60_000)
        elif far and tl < 0:                                   # the right mate itself
            t[3] = b"%d" % (pos + 60_000); t[8] = b"%d" % (tl - 60_000)
        out.append(b"\t".join(t))
        if supp and k % 97 == 0 and not far:
            u = list(t); u[1] = b"%d" % (fl | 2048); u[3] = b"%d" % (pos + 2_000)
            out.append(b"\t".join(u))
    out.sort(key=lambda l: int(l.split(b"\t")[3]))
    return out


def _groups(lines):
    idx = {}
    for i, l in enumerate(lines):
        idx.setdefault(l.split(b"\t", 1)[0], []).append(i)
    return idx


def test_pairing_across_match_maker_buffers(ctx, oracle, genome):
    """160 000 records = four 50 000-line buffers of the reference's match_maker: hundreds of pairs whose mates sit in different
    buffers (the first one is carried over while its PNEXT lies ahead, match_maker.cpp:77-105), singles whose mate never comes --
    pat text and counters == `_ref/match_maker | _ref/patter`"""
    H = oracle; g = genome
    if not H.have_ref():
        pytest.skip("reference executables not built")
    lines = _big_batch(g, supp=False)
    idx = _groups(lines)
    assert len(lines) > 155_000 and sum(1 for v in idx.values() if len(v) == 1) > 1000
    assert sum(1 for v in idx.values() if len(v) == 2 and v[0] // 50_000 != v[1] // 50_000) > 300
    sam = b"\n".join(lines) + b"\n"
    ref_raw, ref_txt = _oracle_pat(H, g, sam, True)
    raw, txt, st = _gpu_pat(ctx, g, sam)
    assert txt == ref_txt
    assert st["pairs"] > 70_000


def test_known_divergences_from_match_maker_buffers(ctx, oracle, genome):
    """match_maker pairs inside 50 000-line buffers (match_maker.cpp:60-105,162-170); this library pairs by QNAME over the whole
    chromosome.  The two differ in exactly two situations, both pinned here with --long output (read names in column 5):
      * a STALE PNEXT: the reference gives up on an unpaired read as soon as its PNEXT lies before the position of a line of the
        current buffer; when the mate really sits in a later buffer it then emits both mates as singles -- here they are merged;
      * a QNAME group of 3+ records (supplementary alignments): match_maker pairs what is adjacent after sorting one buffer's lines,
        re-sorts pairs and singles by position with an unstable std::sort (equal keys: the pair and its third record), and patter then
        pairs consecutive equal names of THAT order again -- which two of the three records end up merged depends on libstdc++'s sort.
        Here the group is ordered by whole line and paired greedily (what match_maker's first step does).
    Every template that is not one of those is identical.  ("pairing".)"""
    H = oracle; g = genome
    if not H.have_ref():
        pytest.skip("reference executables not built")
    lines = _big_batch(g, 110_000, 23, far_every=41)
    idx = _groups(lines)
    special = {q for q, v in idx.items() if len(v) > 2}
    n_split3 = len(special)
    stale = 0
    for q, v in idx.items():
        if len(v) == 2 and v[0] // 50_000 != v[1] // 50_000 and stale < 40:
            t = lines[v[0]].split(b"\t"); t[7] = b"1"; lines[v[0]] = b"\t".join(t); special.add(q); stale += 1
    assert stale >= 20 and n_split3 >= 1
    sam = b"\n".join(lines) + b"\n"
    d = H.write_tmp(g.dict_text(), ".CpG.bed")
    ref_long, _ = H.ref_patter(sam, d, g.chrom, True, long=True)
    ix = ctx.load_index(g.loci, g.first_idx)
    P, st = ctx.pileup_sam(ix, sam, keep_names=True)
    got_long = P.to_text(g.chrom, long=True)
    P.free(); ix.free()
    # patter --long prints `chr idx pattern qname`; the library's long format is the .pat line of bam2pat --long: `chr idx pattern 1 qname`
    rows = lambda txt: [(t[1], t[2], t[-1]) for t in (l.split(b"\t") for l in txt.splitlines())]
    others = lambda txt: sorted(r for r in rows(txt) if r[2] not in special)
    assert others(got_long) == others(ref_long)                 # every other template: identical
    ours = [r for r in rows(got_long) if r[2] in special]
    theirs = [r for r in rows(ref_long) if r[2] in special]
    assert ours and theirs and sorted(ours) != sorted(theirs)

