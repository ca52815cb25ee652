"""cview / view / beta_to_blocks.

CPU part: the per-record logic the device kernels run (csrc/cview_core.cuh) is compiled as host C++ by
tests/cview_core_check.cpp and compared with the reference `cview` executable; the collapse_pat.pl restatement used
by the GPU tests is pinned against the reference's perl script.  GPU part: the same comparisons through the C ABI."""
import gzip
import os
import subprocess

import numpy as np
import pytest

from wgbs_tools_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
N_CPG = 6000


def _pat(seed=3, n_reads=9000, dots=0.12):
    """sorted pat text whose patterns carry inner AND edge dots (what cview --strip / --no_gaps act on), counts 1..9"""
    idx, pats, cnt = synth.make_pat_records(seed, n_reads, N_CPG, first_idx=1, mean_len=6.0, max_len=40)
    rng = np.random.default_rng(seed)
    out = []
    for i, p, c in zip(idx.tolist(), pats, cnt.tolist()):
        b = bytearray(p)
        for k in range(len(b)):
            if rng.random() < dots:
                b[k] = ord(".")
        if not b.strip(b"."):
            b[0] = ord("C")
        out.append((i, bytes(b), c + int(rng.integers(0, 3))))
    out.sort(key=lambda t: (t[0], t[1]))
    return b"".join(b"chr1\t%d\t%s\t%d\n" % t for t in out)


def _blocks(seed, disjoint=True, n=150):
    rng = np.random.default_rng(seed)
    if disjoint:
        cuts = np.sort(rng.choice(np.arange(2, N_CPG + 30), size=2 * n, replace=False))
        return [(int(a), int(b)) for a, b in zip(cuts[0::2], cuts[1::2])]
    st = np.sort(rng.integers(1, N_CPG, size=n))
    bl = [(int(s), int(s + rng.integers(1, 200))) for s in st]
    return bl


def _blocks_file(tmp_path, blocks, name="b.bed"):
    """a 5-column blocks file; cview reads columns 4-5 and sorts them with `sort -k1,1n`"""
    p = tmp_path / name
    p.write_text("".join(f"chr1\t{s * 10}\t{e * 10}\t{s}\t{e}\n" for s, e in blocks))
    return str(p)


def _sorted_like_cview(blocks):
    """`cut -f4-5 | sort -k1,1n` (cview.cpp:21): numeric on the start, ties by the whole "start\\tend" line as bytes"""
    return sorted(blocks, key=lambda b: (b[0], b"%d\t%d" % b))


FLAG_SETS = [dict(), dict(strict=True), dict(strip=True), dict(strict=True, strip=True), dict(no_gaps=True), dict(min_cpgs=4),
             dict(strict=True, strip=True, no_gaps=True, min_cpgs=3), dict(strict=True, min_cpgs=2), dict(strip=True, min_cpgs=5, no_gaps=True)]


@pytest.fixture(scope="module")
def core_check(tmp_path_factory):
    exe = tmp_path_factory.mktemp("cvc") / "cview_core_check"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", str(exe), os.path.join(HERE, "cview_core_check.cpp")])
    return str(exe)


def _run_core(exe, pat, blocks, tmp_path, pre=None, **kw):
    bf = tmp_path / "core_blocks.tsv"; bf.write_text("".join(f"{s}\t{e}\n" for s, e in _sorted_like_cview(blocks)))
    cmd = [exe, str(bf), str(int(kw.get("strict", False))), str(int(kw.get("strip", False))), str(int(kw.get("no_gaps", False))), str(kw.get("min_cpgs", 1))]
    if pre is not None:
        pf = tmp_path / "core_pre.tsv"; pf.write_text("".join(f"{a}\t{b}\n" for a, b in pre)); cmd.append(str(pf))
    return subprocess.run(cmd, input=pat, stdout=subprocess.PIPE, check=True).stdout


def test_cview_core_logic_matches_reference_executable(oracle, core_check, tmp_path):
    H = oracle
    if not H.have_cview():
        pytest.skip("reference cview not built")
    pat = _pat()
    # (a) --sites: one block
    for sites in ((1, N_CPG + 1), (2500, 2600), (1, 40), (N_CPG - 30, N_CPG + 1), (3000, 3001)):
        for kw in FLAG_SETS:
            exp = H.ref_cview(pat, sites=sites, **kw)
            assert _run_core(core_check, pat, [sites], tmp_path, **kw) == exp, (sites, kw)
    assert len(H.ref_cview(pat, sites=(2500, 2600), strict=True)) > 500
    # (b) --blocks_path with disjoint blocks, every flag combination
    bl = _blocks(1)
    bf = _blocks_file(tmp_path, [bl[i] for i in np.random.default_rng(0).permutation(len(bl))])       # file order is irrelevant: cview sorts
    for kw in FLAG_SETS:
        exp = H.ref_cview(pat, blocks_path=bf, **kw)
        assert len(exp) > 1000
        assert _run_core(core_check, pat, bl, tmp_path, **kw) == exp, kw
    # (c) overlapping / nested / duplicated blocks without --strict (with it the reference aborts): the cursor rule
    bl = _blocks(2, disjoint=False) + [(100, 4000), (100, 150), (100, 150), (5990, 5995)]
    bf = _blocks_file(tmp_path, bl, "ovl.bed")
    for kw in (dict(), dict(strip=True), dict(no_gaps=True, min_cpgs=3)):
        exp = H.ref_cview(pat, blocks_path=bf, **kw)
        assert len(exp) > 1000
        assert _run_core(core_check, pat, bl, tmp_path, **kw) == exp, kw
    # (d) the last block ends early although an earlier one reaches further: everything from its end on is cut off
    bl = [(10, 5000), (20, 30)]
    bf = _blocks_file(tmp_path, bl, "early.bed")
    exp = H.ref_cview(pat, blocks_path=bf)
    assert exp and max(int(l.split(b"\t")[1]) for l in exp.splitlines()) < 30
    assert _run_core(core_check, pat, bl, tmp_path) == exp


def test_collapse_pat_port_matches_perl_script(oracle):
    H = oracle
    pat = _pat(5, 3000)
    lines = pat.splitlines(keepends=True)
    dup = b"".join(l * (1 + i % 3) for i, l in enumerate(lines)) + b"chr1\t7000\tCC\t0\nchr1\t7001\tTT\t2\nchr1\t7001\tTT\t5\n"
    ref = H.ref_collapse_pat(dup)
    if ref is None:
        pytest.skip("reference tree not mounted")
    assert H.port_collapse_pat(dup) == ref and ref.endswith(b"chr1\t7001\tTT\t7\n") and b"\t7000\t" not in ref


# ----------------------------------------------------------------------------------------------------------------------
# GPU
# ----------------------------------------------------------------------------------------------------------------------
def _gpu_view(ctx, pat, blocks, mode, pre=None, **kw):
    P = ctx.pats_from_text(pat)
    bl = _sorted_like_cview(blocks)
    V = P.cview([b[0] for b in bl], [b[1] for b in bl], pre=pre, **kw)
    raw = V.to_text("chr1")
    V.collapse(mode=mode)
    txt = V.to_text("chr1")
    P.free(); V.free()
    return raw, txt


@pytest.mark.gpu
def test_cview_gpu_matches_reference(ctx, oracle, tmp_path):
    H = oracle
    pat = _pat()
    have = H.have_cview()
    cases = [([(2500, 2600)], None), ([(1, N_CPG + 1)], None), (_blocks(1), "b1.bed"), (_blocks(7, n=600), "b7.bed")]
    for blocks, fname in cases:
        bf = _blocks_file(tmp_path, blocks, fname) if fname else None
        for kw in FLAG_SETS:
            if have:
                exp = H.ref_cview(pat, blocks_path=bf, **kw) if bf else H.ref_cview(pat, sites=blocks[0], **kw)
            else:
                pytest.skip("reference cview not built")
            raw, txt = _gpu_view(ctx, pat, blocks, 2, **kw)
            assert raw == exp, (fname, kw)
            assert txt == H.port_collapse_pat(H.sort_pat(exp)), (fname, kw)                 # `| sort -k2,2n -k3,3 | collapse_pat.pl -`
            raw, adj = _gpu_view(ctx, pat, blocks, 3, **kw)
            assert adj == H.port_collapse_pat(exp), (fname, kw)                              # whole-file view: no sort (cview.py:47)
    # overlapping blocks, non-strict; and the tabix pre-selection of start indices
    bl = _blocks(2, disjoint=False) + [(100, 4000), (100, 150), (100, 150)]
    bf = _blocks_file(tmp_path, bl, "ovl.bed")
    raw, _ = _gpu_view(ctx, pat, bl, 2, strip=True)
    assert raw == H.ref_cview(pat, blocks_path=bf, strip=True)
    pre = [(2350, 2599)]
    sub = b"".join(l for l in pat.splitlines(keepends=True) if 2350 <= int(l.split(b"\t")[1]) <= 2599)    # tabix chr1:2350-2599
    raw, _ = _gpu_view(ctx, pat, [(2500, 2600)], 2, pre=(np.array([2350]), np.array([2599])), strict=True)
    assert raw == H.ref_cview(sub, sites=(2500, 2600), strict=True) and raw


@pytest.mark.gpu
def test_cview_rejects_what_the_reference_rejects(ctx):
    from wgbs_tools_b200._lib import WgbsError
    P = ctx.pats_from_text(b"chr1\t5\tCT\t1\n")
    for bs, be, kw, msg in (([5], [5], {}, "endCpG <= startCpG"), ([0], [5], {}, "startCpG < 1"), ([9, 3], [12, 5], {}, "sorted"),
                            ([3, 4], [8, 9], dict(strict=True), "non-overlapping"), ([], [], {}, "0 blocks")):
        with pytest.raises(WgbsError, match=msg):
            P.cview(bs, be, **kw)
    P.free()


@pytest.mark.gpu
def test_beta_to_blocks_matches_numpy_reduceat(ctx, oracle):
    H = oracle
    rng = np.random.default_rng(4)
    n = 200_000
    beta = synth.make_betas(1, 1, n)[0]
    blocks = synth.make_blocks(2, 1, n, mean_len=8.0)
    big = np.array([[1, n + 1], [5, 5], [n - 3, n + 50], [1000, 90_000]], blocks.dtype)           # whole file, empty, past the end, long
    bl = np.concatenate([blocks, big])
    exp = np.array([beta[s - 1:e - 1].sum(axis=0, dtype=np.int64) for s, e in bl.tolist()])
    out, sums = ctx.beta_to_blocks(beta, bl[:, 0], bl[:, 1], 8, want_sums=True)
    assert np.array_equal(sums, exp) and sums[-4, 1] > 2**22
    assert out.tobytes() == H.ref_trim(exp).tobytes()
    # nice blocks: the reference's fast path (np.add.reduceat over the sorted borders, beta_to_blocks.py:101-105)
    bins = np.sort(np.unique(blocks.flatten()))[:-1]
    fast = np.add.reduceat(beta, bins - 1)[np.isin(bins, blocks[:, 0])]
    assert np.array_equal(sums[:blocks.shape[0]], fast)
    # .lbeta in, .lbeta out
    lb = (beta.astype(np.uint16) * 97) % 60000
    out16, s16 = ctx.beta_to_blocks(lb, bl[:, 0], bl[:, 1], 16, want_sums=True)
    exp16 = np.array([lb[s - 1:e - 1].sum(axis=0, dtype=np.int64) for s, e in bl.tolist()])
    assert np.array_equal(s16, exp16) and out16.tobytes() == H.ref_trim(exp16, lbeta=True).tobytes()
