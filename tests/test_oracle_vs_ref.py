"""Pins the oracle: the plain-C restatement (oracle/oracle_port.c) against the UNMODIFIED reference executables
compiled from /root/reference/src into oracle/_ref/  -- byte for byte, on seeded inputs and on
the edge cases the reference's code handles explicitly.  Runs on CPU."""
import numpy as np
import pytest

from wgbs_tools_b200 import synth


@pytest.fixture(scope="module")
def H(oracle):
    if not oracle.have_ref():
        pytest.skip("reference executables not built (no /root/reference and no prebuilt oracle/_ref)")
    return oracle


@pytest.fixture(scope="module")
def genome():
    return synth.make_genome(7, "chrT", 1_000_000)


def _dict(H, g):
    return H.write_tmp(g.dict_text(), ".CpG.bed")


@pytest.mark.parametrize("paired", [True, False])
@pytest.mark.parametrize("clip,min_cpg", [(0, 1), (5, 2)])
def test_patter_port_matches_reference(H, genome, paired, clip, min_cpg):
    sam = synth.make_sam(genome, 6000, 11 + clip, paired=paired)
    d = _dict(H, genome)
    out, err = H.ref_patter(sam, d, genome.chrom, paired, min_cpg=min_cpg, clip=clip)
    mm = H.port_match_maker(sam) if paired else sam
    pout, st = H.port_patter(mm, genome.loci, genome.idx(), min_cpg=min_cpg, clip=clip)
    assert sorted(out.splitlines()) == sorted(pout.splitlines())
    assert len(out) > 1000
    # the stats line is a secondary parity check (patter.cpp:298-316)
    msg = err.decode().strip().splitlines()[-1]
    assert f"finished {st[0]:,} lines" in msg and f"{st[2]:,} empty" in msg and f"{st[4]:,} invalid" in msg
    assert H.ref_collapse(out) == H.port_collapse(pout)


def test_patter_port_invalid_and_odd_reads(H, genome):
    g = genome
    p = int(g.loci[100]) - 20
    seq = g.bases[p:p + 60].tobytes()
    lines = [
        b"a\t0\tchrT\t%d\t60\t60M\t*\t0\t0\t%s\t*" % (p, seq),
        b"b\t0\tchrT\t%d\t60\t*\t*\t0\t0\t%s\t*" % (p, seq),                 # CIGAR '*' -> invalid
        b"c\t0\tchrT\t%d\t60\t70M\t*\t0\t0\t%s\t*" % (p, seq),               # M longer than SEQ -> invalid
        b"d\t0\tchrT\t%d\t60\t30M2P30M\t*\t0\t0\t%s\t*" % (p, seq),          # P -> invalid
        b"e\t16\tchrT\t%d\t60\t10S40M10H\t*\t0\t0\t%s\t*" % (p, seq),
        b"f\t0\tchrT\t%d\t60\t20M5N35M5S\t*\t0\t0\t%s\t*" % (p, seq),
        b"g\t0\tchrT\t%d\t60\t20=5X35M\t*\t0\t0\t%s\t*" % (p, seq),
        b"h\t0\tchrT",                                                        # too few fields -> invalid
        b"i\t0\tchrT\t%d\t60\t60M\t*\t0\t0\t*\t*" % p,                        # SEQ '*' -> invalid (M > len)
        b"j\t0\tchrT\t%d\t60\t60M\t*\t0\t0\t%s\t*" % (int(g.loci[-1]) - 5, seq),  # runs past the last CpG
        b"k\t0\tchrT\t%d\t60\t5M\t*\t0\t0\t%s\t*" % (p, seq),                 # CIGAR shorter than SEQ
        b"",
    ]
    sam = b"\n".join(lines) + b"\n"
    out, err = H.ref_patter(sam, _dict(H, g), g.chrom, False)
    pout, st = H.port_patter(sam, g.loci, g.idx())
    assert out == pout
    assert st[4] == 5 and b"5 invalid" in err


@pytest.mark.parametrize("kw", [dict(), dict(combine_mods=True), dict(cpc_call="H"), dict(np_thresh=0.8), dict(clip=3)])
def test_np_port_matches_reference(H, genome, kw):
    sam = synth.make_np_sam(genome, 1500, 5)
    rkw = dict(kw); rkw.setdefault("np_thresh", 0.67)
    out, _ = H.ref_patter(sam, _dict(H, genome), genome.chrom, False, nanopore=True, **rkw)
    pout, _ = H.port_patter(sam, genome.loci, genome.idx(), nanopore=True, **rkw)
    assert out == pout and len(out) > 1000


def test_np_known_answer(H, genome):
    """SURVEY 8c micro-vector: per-CpG (m,h) ML pairs at t=0.67 -> C H T . C ; --combine_mods -> C C T . C"""
    g = genome
    # build a top-strand read covering >= 5 CpGs whose only C's we annotate are the CpG C's
    k = 200
    while g.loci[k + 4] - g.loci[k] > 120:
        k += 1
    p = int(g.loci[k]) - 2
    n = int(g.loci[k + 4]) - p + 3
    seq = g.bases[p:p + n].tobytes()
    cs = [j for j, ch in enumerate(seq) if ch == ord("C")]
    cpg = [i for i, j in enumerate(cs) if seq[j + 1:j + 2] == b"G"][:5]
    deltas = [cpg[0]] + [cpg[i] - cpg[i - 1] - 1 for i in range(1, 5)]
    d = ",".join(map(str, deltas)).encode()
    sam = b"q\t0\tchrT\t%d\t60\t%dM\t*\t0\t0\t%s\t*\tMM:Z:C+m?,%s;C+h?,%s;\tML:B:C,250,3,5,128,200,2,250,5,10,40\n" % (p, n, seq, d, d)
    out, _ = H.ref_patter(sam, _dict(H, g), g.chrom, False, nanopore=True, np_thresh=0.67)
    assert out.split(b"\t")[2].strip() == b"CHT.C"
    out2, _ = H.ref_patter(sam, _dict(H, g), g.chrom, False, nanopore=True, np_thresh=0.67, combine_mods=True)
    assert out2.split(b"\t")[2].strip() == b"CCT.C"
    assert H.port_patter(sam, g.loci, g.idx(), nanopore=True)[0] == out
    assert H.port_patter(sam, g.loci, g.idx(), nanopore=True, combine_mods=True)[0] == out2


def test_pat2beta_trim_port_matches_reference(H):
    idx, pats, cnt = synth.make_pat_records(1, 10_000, 50_000)
    txt = synth.pat_text("chr1", idx, pats, cnt)
    ref = H.ref_stdin2beta(txt, 1, 50_001)
    assert ref.shape == (50_000, 2)
    np.testing.assert_array_equal(ref, H.port_pat2beta(txt, 1, 50_001))
    # sub-range (mult_pat2beta style, pat2beta.py:60-65)
    np.testing.assert_array_equal(H.ref_stdin2beta(txt, 20_000, 30_000), H.port_pat2beta(txt, 20_000, 30_000))
    big = ref * 37
    np.testing.assert_array_equal(H.ref_trim(big), H.port_trim(big))
    np.testing.assert_array_equal(H.ref_trim(big * 40, lbeta=True), H.port_trim(big * 40, 16))
    # SURVEY 8c known answers
    ka = np.array([[100, 510], [255, 256], [7, 1000]])
    np.testing.assert_array_equal(H.port_trim(ka), [[50, 255], [254, 255], [1, 255]])
    np.testing.assert_array_equal(H.ref_trim(ka), [[50, 255], [254, 255], [1, 255]])


def test_pat2beta_bad_line_fails_whole_run(H):
    txt = b"chr1\t5\tCC\t2\nchr1\t7\tT\n"
    assert H.ref_stdin2beta(txt, 1, 20).size == 0
    with pytest.raises(ValueError):
        H.port_pat2beta(txt, 1, 20)


@pytest.mark.parametrize("inclusive", [False, True])
@pytest.mark.parametrize("l,rng", [(3, "0,0.334,0.667,1"), (1, "0,0.25,0.5,0.75,1"), (5, "0.1,0.9,1")])
def test_homog_port_matches_reference(H, inclusive, l, rng):
    if rng.startswith("0.1"):
        rng = "0,0.1,0.9,1"
    N = 20_000
    idx, pats, cnt = synth.make_pat_records(3, 30_000, N, mean_len=6)
    txt = synth.pat_text("chr1", idx, pats, cnt)
    blocks = synth.make_blocks(9, 1, N - 500)              # reads run past the last block: exercises the break
    bp = H.write_tmp(synth.blocks_text("chr1", blocks), ".bed")
    r = np.array([float(x) for x in rng.split(",")], np.float32)
    ref = H.ref_homog(txt, bp, rng, l, inclusive=inclusive)
    np.testing.assert_array_equal(ref, H.port_homog(txt, blocks, r, l, inclusive))
    assert ref.sum() > 1000


def test_homog_overlapping_blocks_and_known_answer(H):
    txt = b"chr1\t2\tCC..TH\t2\nchr1\t4\tTTTT\t1\nchr1\t30\tCCC\t5\nchr1\t41\tCCC\t7\nchr1\t60\tTTT\t1\n"
    blocks = np.array([[2, 8], [3, 50], [5, 6], [40, 45]], np.int32)      # nested / overlapping; last block ends before an earlier one
    bp = H.write_tmp(synth.blocks_text("chr1", blocks), ".bed")
    r = np.array([0, .334, .667, 1], np.float32)
    ref = H.ref_homog(txt, bp, "0,0.334,0.667,1", 1)
    np.testing.assert_array_equal(ref, H.port_homog(txt, blocks, r, 1))
    assert ref[0, 2] == 2                                                  # M += 2


def test_segment_port_matches_reference(H):
    n = 2500
    betas = synth.make_betas(4, 6, n)
    paths = [H.write_tmp(b.tobytes(), f".{i}.beta") for i, b in enumerate(betas)]
    g = synth.make_genome(2, "chr1", 400_000, with_bases=False)
    d = g.loci[:n]
    for max_cpg, max_bp, ps in [(200, 2000, 15), (1000, 5000, 1), (50, 300, 0.5)]:
        ref = H.ref_segmentor(paths, 0, n, max_cpg, max_bp, ps, d)
        np.testing.assert_array_equal(ref, H.port_segment(betas, d, max_cpg, max_bp, ps))
        assert ref[0] == 0 and ref[-1] == n


@pytest.mark.parametrize("seed,np_mode", [(1, False), (2, False), (3, True), (4, True)])
def test_fuzzed_sam_port_matches_reference(H, genome, seed, np_mode):
    """hostile SAM text (random CIGARs, truncated lines, odd FLAG/POS, malformed MM/ML): the port tracks the reference"""
    from fuzz_sam import fuzz_sam
    sam = fuzz_sam(genome, seed, 700, np_mode)
    # the first line decides PE / MM mode and a broken first line is fatal in the reference: keep it valid
    first = (synth.make_np_sam(genome, 1, 99) if np_mode else synth.make_sam(genome, 1, 99, paired=False))
    sam = first + sam
    kw = dict(nanopore=True, np_thresh=0.67) if np_mode else {}
    out, err = H.ref_patter(sam, _dict(H, genome), genome.chrom, False, **kw)
    pout, st = H.port_patter(sam, genome.loci, genome.idx(), **kw)
    assert out == pout
    msg = err.decode().strip().splitlines()[-1]
    assert f"{st[2]:,} empty" in msg and f"{st[4]:,} invalid" in msg, (msg, st)
    assert st[4] > 50
