"""GPU parity: pat text parse, pat2beta (+trim) and homog through the C ABI vs the oracle (reference executables and
C restatement).  BASELINE.json config 1 lives here: 10k-read pat over a 50k-CpG index, bit-exact .beta."""
import numpy as np
import pytest

from wgbs_tools_b200 import synth

pytestmark = pytest.mark.gpu


def _ref_beta(H, txt, start, end):
    # the reference allocates its counters with `new int[n]` and never clears them (stdin2beta.cpp:48-49): that only
    # reads as zero when the allocation is large enough to be fresh mmap'd pages, so small ranges go to the port
    if H.have_ref() and end - start >= 40_000:
        return H.ref_stdin2beta(txt, start, end)
    return H.port_pat2beta(txt, start, end)


def test_config1_pat2beta_bit_exact(ctx, oracle):
    H = oracle
    N = 50_000
    idx, pats, cnt = synth.make_pat_records(1, 10_000, N)
    txt = synth.pat_text("chr1", idx, pats, cnt)
    beta, mc = ctx.pat2beta_text(txt, 1, N + 1, want_counts=True)
    ref = _ref_beta(H, txt, 1, N + 1)
    np.testing.assert_array_equal(mc, ref)
    np.testing.assert_array_equal(mc, H.port_pat2beta(txt, 1, N + 1))
    assert beta.tobytes() == H.ref_trim(ref).tobytes()          # the .beta file bytes
    assert beta.shape == (N, 2) and beta.dtype == np.uint8


def test_pat_parse_roundtrip(ctx):
    idx, pats, cnt = synth.make_pat_records(2, 5_000, 20_000, mean_len=9, max_len=70)
    txt = synth.pat_text("chr7", idx, pats, cnt)
    P = ctx.pats_from_text(txt)
    i2, ln, c2, off, pool = P.download()
    np.testing.assert_array_equal(i2, idx)
    np.testing.assert_array_equal(c2, cnt)
    assert P.patterns() == pats
    assert max(len(p) for p in pats) > 32                           # multi-word records exercised
    P.free()


def test_pat_text_edge_cases(ctx, oracle):
    H = oracle
    # no trailing newline, empty lines, extra columns, '.' and 'H' symbols, record straddling both range ends
    txt = b"chr1\t3\tCT.H\t2\n\nchr1\t1\tTTTTTTTTTTTTTTTTTTTTCCCC\t1\textra\tcols\nchr1\t18\tC.T\t7"
    for s, e in [(1, 30), (5, 20), (19, 21), (25, 40)]:
        beta, mc = ctx.pat2beta_text(txt, s, e, want_counts=True)
        np.testing.assert_array_equal(mc, _ref_beta(H, txt, s, e))
    # negative and zero counts are legal for std::stoi and simply add up
    txt2 = b"chr1\t3\tCT.H\t5\nchr1\t4\tTC\t-2\nchr1\t4\tCC\t0\n"
    beta, mc = ctx.pat2beta_text(txt2, 1, 12, want_counts=True)
    np.testing.assert_array_equal(mc, H.port_pat2beta(txt2, 1, 12))
    # empty input
    beta, mc = ctx.pat2beta_text(b"", 1, 11, want_counts=True)
    assert mc.sum() == 0 and beta.shape == (10, 2)


def test_pat_bad_lines_fail_like_reference(ctx):
    from wgbs_tools_b200._lib import WgbsError
    with pytest.raises(WgbsError, match="too few columns"):
        ctx.pat2beta_text(b"chr1\t5\tCC\t2\nchr1\t7\tT\n", 1, 20)
    with pytest.raises(WgbsError, match="non-numeric"):
        ctx.pat2beta_text(b"chr1\tx\tCC\t2\n", 1, 20)


def test_trim_known_answers_and_lbeta(ctx, oracle):
    H = oracle
    rng = np.random.default_rng(0)
    cov = rng.integers(0, 70_000, size=100_000)
    meth = (cov * rng.random(100_000)).astype(np.int64)
    mc = np.stack([meth, cov], 1).astype(np.int32)
    mc[:3] = [[100, 510], [255, 256], [7, 1000]]
    out8 = ctx.trim(mc, mc.shape[0], 8)
    np.testing.assert_array_equal(out8[:3], [[50, 255], [254, 255], [1, 255]])
    np.testing.assert_array_equal(out8, H.ref_trim(mc))
    np.testing.assert_array_equal(ctx.trim(mc, mc.shape[0], 16), H.ref_trim(mc, lbeta=True))


def test_pat2beta_high_coverage_and_unsorted(ctx, oracle):
    """cover > 255 (trim path), many records per site (shared-memory window), and records far outside the CTA
    window (global fallback)."""
    H = oracle
    N = 3_000
    idx, pats, cnt = synth.make_pat_records(5, 60_000, N, mean_len=8)
    cnt = cnt * 3
    order = np.random.default_rng(1).permutation(idx.size)         # deliberately unsorted
    txt = synth.pat_text("chr1", idx[order], [pats[i] for i in order], cnt[order])
    beta, mc = ctx.pat2beta_text(txt, 1, N + 1, want_counts=True)
    ref = _ref_beta(H, txt, 1, N + 1)
    np.testing.assert_array_equal(mc, ref)
    assert ref[:, 1].max() > 255
    assert beta.tobytes() == H.ref_trim(ref).tobytes()


def test_pat2beta_accumulates_across_batches(ctx, oracle):
    H = oracle
    N = 10_000
    idx, pats, cnt = synth.make_pat_records(6, 20_000, N)
    txt = synth.pat_text("chr1", idx, pats, cnt)
    lines = txt.splitlines(keepends=True)
    a, b = b"".join(lines[: len(lines) // 2]), b"".join(lines[len(lines) // 2:])
    Pa, Pb = ctx.pats_from_text(a), ctx.pats_from_text(b)
    buf = ctx.pat2beta(Pa, 1, N + 1)
    ctx.pat2beta(Pb, 1, N + 1, meth_cov=buf, zero_first=False)
    mc = buf.to_host(np.int32).reshape(-1, 2)
    np.testing.assert_array_equal(mc, _ref_beta(H, txt, 1, N + 1))


@pytest.mark.parametrize("inclusive", [False, True])
@pytest.mark.parametrize("l,rng", [(3, "0,0.334,0.667,1"), (1, "0,0.25,0.5,0.75,1"), (5, "0,0.1,0.9,1")])
def test_homog_matches_oracle(ctx, oracle, inclusive, l, rng):
    H = oracle
    N = 20_000
    idx, pats, cnt = synth.make_pat_records(3, 30_000, N, mean_len=6)
    txt = synth.pat_text("chr1", idx, pats, cnt)
    blocks = synth.make_blocks(9, 1, N - 500)
    r = np.array([float(x) for x in rng.split(",")], np.float32)
    P = ctx.pats_from_text(txt)
    got = ctx.homog(P, blocks, r, l, inclusive)
    np.testing.assert_array_equal(got, H.port_homog(txt, blocks, r, l, inclusive))
    if H.have_ref():
        bp = H.write_tmp(synth.blocks_text("chr1", blocks), ".bed")
        np.testing.assert_array_equal(got, H.ref_homog(txt, bp, rng, l, inclusive=inclusive))
    assert got.sum() > 1000


def test_homog_overlapping_blocks(ctx, oracle):
    H = oracle
    txt = b"chr1\t2\tCC..TH\t2\nchr1\t4\tTTTT\t1\nchr1\t30\tCCC\t5\nchr1\t41\tCCC\t7\nchr1\t60\tTTT\t1\n"
    blocks = np.array([[2, 8], [3, 50], [5, 6], [40, 45]], np.int32)
    r = np.array([0, .334, .667, 1], np.float32)
    P = ctx.pats_from_text(txt)
    got = ctx.homog(P, blocks, r, 1)
    np.testing.assert_array_equal(got, H.port_homog(txt, blocks, r, 1))
    assert got[0, 2] == 2


def test_collapse_of_very_long_tie_runs(ctx, oracle):
    """amplicon-like input: 130 000 templates start at ONE CpG with the same first calls and differ further on (one 32-bit sort key,
    so the whole pile is one run of equal keys): runs longer than 64 are sorted by whole CTAs (fix_long_runs_k), not by one thread's
    insertion sort; plus many medium runs around the threshold.  Text == `sort -k2,2n -k3,3 | uniq -c | awk`."""
    H = oracle
    rng = np.random.default_rng(12)
    sym = np.frombuffer(b"CT.H", np.uint8)
    rows = []
    head = b"CTCTCTCTCTCTCTCT"                                          # 16 symbols: more than the key holds
    for _ in range(130_000):
        tail = sym[rng.integers(0, 4, size=int(rng.integers(1, 14)))].tobytes().rstrip(b".") or b"C"
        rows.append(b"chr1\t5000\t" + head + tail)
    for site in range(6000, 6400):                                      # runs of 1..130 around the serial / CTA threshold
        for _ in range(int(rng.integers(1, 131))):
            tail = sym[rng.integers(0, 4, size=int(rng.integers(1, 9)))].tobytes().rstrip(b".") or b"T"
            rows.append(b"chr1\t%d\t" % site + head + tail)
    order = rng.permutation(len(rows))
    raw3 = b"\n".join(rows[i] for i in order) + b"\n"
    text4 = b"\n".join(rows[i] + b"\t1" for i in order) + b"\n"
    P = ctx.pats_from_text(text4)
    P.collapse()
    got = P.to_text("chr1")
    P.free()
    assert got == H.ref_collapse(raw3) if H.have_ref() else H.port_collapse(raw3)
