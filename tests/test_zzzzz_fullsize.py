"""BASELINE.json's configurations at FULL size, checked through size-independent properties (the oracle cannot run 1M reads
or 100M pat records in test time): sharding invariance (what the multi-GPU split relies on: counts are sums over records),
conservation (every called symbol lands in exactly one cover counter), round trips through the on-disk text format,
sortedness / idempotence of the collapse -- plus the oracle itself on a bounded sample taken at the far end of the index.

The property code is written against a small `engine` interface and runs twice: on the GPU at full size (`-m gpu`), and on
the CPU with the oracle's C restatement as the engine at a small size, which pins the properties themselves (a property that
does not hold for the reference's algorithm would be a wrong test, not a finding).  (File name: runs last.)"""
import os
import zlib

import numpy as np
import pytest

from wgbs_tools_b200 import synth

STAT_SUM = ("lines", "pairs", "empty", "short", "invalid", "templates")


# ---- engines ------------------------------------------------------------------------------------------------------------------
class GpuEngine:
    def __init__(self, ctx):
        self.ctx = ctx

    def pile(self, g, sam):
        """SAM text -> (collapsed pat text, int32[N,2] counts, stats)"""
        ix = self.ctx.load_index(g.loci, g.first_idx)
        P, st = self.ctx.pileup_sam(ix, sam)
        buf = self.ctx.pat2beta(P, g.first_idx, g.first_idx + g.n_cpg)
        mc = buf.to_host(np.int32).reshape(-1, 2)
        buf.free()
        P.collapse()
        txt = P.to_text(g.chrom)
        P.free(); ix.free()
        return txt, mc, st

    def p2b(self, pat, start, end):
        return self.ctx.pat2beta_text(pat, start, end, want_counts=True)[1]

    def homog(self, pat, blocks, rng, min_cpgs):
        P = self.ctx.pats_from_text(pat)
        out = self.ctx.homog(P, blocks, rng, min_cpgs)
        P.free()
        return out


class PortEngine:
    """the oracle's C restatement behind the same interface (CPU, small inputs)"""

    def __init__(self, H):
        self.H = H

    def pile(self, g, sam):
        out, pst = self.H.port_patter(self.H.port_match_maker(sam), g.loci, g.idx())
        txt = self.H.port_collapse(out)
        st = dict(zip(("lines", "pairs", "empty", "short", "invalid", "paired"), pst))
        st["templates"] = len(out.splitlines())
        return txt, self.H.port_pat2beta(txt, g.first_idx, g.first_idx + g.n_cpg), st

    def p2b(self, pat, start, end):
        return self.H.port_pat2beta(pat, start, end)

    def homog(self, pat, blocks, rng, min_cpgs):
        return self.H.port_homog(pat, blocks, rng, min_cpgs)


# ---- the properties -----------------------------------------------------------------------------------------------------------
def split_by_template(sam: bytes):
    """two coordinate-sorted SAM texts; the records of one QNAME stay together (so every pair is still a pair)"""
    a, b = [], []
    for l in sam.splitlines(keepends=True):
        (a if zlib.crc32(l[:l.index(b"\t")]) & 1 else b).append(l)
    return b"".join(a), b"".join(b)


def pat_counts(txt: bytes) -> dict:
    d = {}
    for l in txt.splitlines():
        k, c = l.rsplit(b"\t", 1)
        d[k] = d.get(k, 0) + int(c)
    return d


def check_pileup_properties(E, g, sam):
    N = g.n_cpg
    txt, mc, st = E.pile(g, sam)
    assert st["lines"] == sam.count(b"\n") and st["paired"] == 1 and st["pairs"] > 0
    lines = txt.splitlines()
    # collapse: sorted by (idx numeric, pattern bytes), unique, counts add up to the templates
    keys = [(int(f[1]), f[2]) for f in (l.split(b"\t") for l in lines)]
    assert all(keys[i] < keys[i + 1] for i in range(len(keys) - 1))
    cnt = np.array([int(l.rsplit(b"\t", 1)[1]) for l in lines], np.int64)
    assert int(cnt.sum()) == st["templates"]
    # conservation: every C/T/H of every template is one unit of cover, every C/H one unit of meth
    pats = [k[1] for k in keys]
    called = np.array([len(p) - p.count(b".") for p in pats], np.int64); meth = np.array([p.count(b"C") + p.count(b"H") for p in pats], np.int64)
    assert int(mc[:, 1].sum()) == int((called * cnt).sum()) and int(mc[:, 0].sum()) == int((meth * cnt).sum())
    assert all(p[:1] != b"." and p[-1:] != b"." for p in pats)
    # round trip through the on-disk text: counts from the pat text == counts from the templates
    np.testing.assert_array_equal(E.p2b(txt, g.first_idx, g.first_idx + N), mc)
    # sharding by template: counts, statistics and the collapsed multiset are sums over the shards
    sa, sb = split_by_template(sam)
    ta, mca, sta = E.pile(g, sa)
    tb, mcb, stb = E.pile(g, sb)
    np.testing.assert_array_equal(mca + mcb, mc)
    assert all(sta[k] + stb[k] == st[k] for k in STAT_SUM), (sta, stb, st)
    merged = pat_counts(ta)
    for k, c in pat_counts(tb).items():
        merged[k] = merged.get(k, 0) + c
    assert merged == pat_counts(txt)
    return mc, st


def check_pat_sharding(E, pat, start, end, blocks, rng, min_cpgs):
    """pat2beta and homog over a batch == the sum over two record shards of it; returns (counts, bins)"""
    cut = pat.index(b"\n", len(pat) // 2) + 1
    mc = E.p2b(pat, start, end)
    np.testing.assert_array_equal(E.p2b(pat[:cut], start, end) + E.p2b(pat[cut:], start, end), mc)
    hb = E.homog(pat, blocks, rng, min_cpgs)
    np.testing.assert_array_equal(E.homog(pat[:cut], blocks, rng, min_cpgs) + E.homog(pat[cut:], blocks, rng, min_cpgs), hb)
    return mc, hb


# ---- CPU: the properties hold for the reference's algorithm (oracle port as the engine) ----------------------------------------------
def test_properties_hold_for_the_oracle(oracle):
    E = PortEngine(oracle)
    g = synth.make_genome(7, "chrT", 400_000)
    check_pileup_properties(E, g, synth.make_sam(g, 6_000, 11, paired=True))
    N = 60_000
    pat = synth.make_pat_text_fast(3, 30_000, N, chrom="chr1")
    blocks = synth.make_blocks(5, 1, N)
    mc, hb = check_pat_sharding(E, pat, 1, N + 1, blocks, np.array([0, 0.334, 0.667, 1], np.float32), 3)
    assert mc[:, 1].sum() > 0 and hb.sum() > 0


# ---- GPU, full size ---------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_config2_bam2pat_1m_reads_chr19_index(ctx, oracle):
    """configs[1]: 1M records of 150 bp PE reads over a chr19-sized index (58.6 Mbp, 1.1M CpGs) -- the bench workload"""
    g = synth.make_genome(19, "chr19", 58_617_616, n_cpg=1_100_000)
    sam = synth.make_sam(g, 1_000_000, 1000, paired=True)
    mc, st = check_pileup_properties(GpuEngine(ctx), g, sam)
    assert st["lines"] > 990_000 and st["templates"] > 400_000
    # the .beta bytes of the full-size counts: numpy float64 trim of the reference (utils_wgbs.py:277-290) via the oracle
    dev = ctx.upload(mc)
    beta = ctx.trim(dev, g.n_cpg)
    dev.free()
    assert beta.tobytes() == oracle.port_trim(mc).tobytes()
    # the oracle on a bounded sample from the far end of the chromosome (large POS, large CpG indices)
    tail = sam[sam.rindex(b"\n", 0, len(sam) - 6_000_000) + 1:]
    ta, _ = split_by_template(tail)
    out, pst = oracle.port_patter(oracle.port_match_maker(ta), g.loci, g.idx())
    txt, _, st2 = GpuEngine(ctx).pile(g, ta)
    assert txt == oracle.port_collapse(out)
    assert [st2[k] for k in ("lines", "pairs", "empty", "short", "invalid", "paired")] == pst


@pytest.mark.gpu
def test_config4_segment_200_betas_chunk(ctx, oracle):
    """configs[3]: K = 200 betas, the reference's 60 000-site chunk, max_cpg 5000 (effective min(max_cpg, max_bp // 2), segment.py:65),
    pcount 15: borders identical to segmentor's; a chunk solved inside a many-chunk call == the same chunk solved alone"""
    K, S, chunk = 200, 180_000, 60_000
    betas = synth.make_betas(4, K, S)
    loci = synth.make_genome(2, "chr1", 40_000_000, with_bases=False).loci[:S]
    assert loci.size == S
    max_bp = 2000; max_cpg = min(5000, max_bp // 2)
    dbet = [ctx.upload(b) for b in betas]; dd = ctx.upload(loci)
    chunks = [(s, chunk) for s in range(0, S, chunk)]
    res = ctx.segment(dbet, dd, chunks, max_cpg, max_bp, 15)
    alone = ctx.segment(dbet, dd, [chunks[1]], max_cpg, max_bp, 15)[0]
    for b in dbet + [dd]:
        b.free()
    np.testing.assert_array_equal(res[1], alone)
    for r in res:
        assert r[0] == 0 and r[-1] == chunk and (np.diff(r) > 0).all() and (np.diff(r) <= max_cpg).all()
    H = oracle
    if H.have_ref():
        paths = [H.write_tmp(b.tobytes(), f".{i}.beta") for i, b in enumerate(betas)]
        try:
            ref = H.ref_segmentor(paths, chunk, chunk, max_cpg, max_bp, 15, loci[chunk:2 * chunk])
        finally:
            for p in paths:
                os.remove(p)
    else:
        ref = H.port_segment([b[chunk:2 * chunk] for b in betas], loci[chunk:2 * chunk], max_cpg, max_bp, 15)
    np.testing.assert_array_equal(res[1], ref)


@pytest.mark.gpu
def test_config3_homog_and_pat2beta_hg38_index_100m_records(ctx, oracle):
    """configs[2]: U/X/M homog (and the pat2beta reduction) over a 28.2M-CpG index, 100M pat records, in batches the way a
    chromosome-sharded run feeds them; every batch: sum over two shards == whole, accumulated counts == sum of batches"""
    E = GpuEngine(ctx)
    N = 28_217_448                                              # CpGs of hg38 (CpG.bed.gz of the reference's init_genome)
    total = int(os.environ.get("WGBS_FULLSIZE_RECORDS", 100_000_000))
    per = 12_500_000
    blocks = synth.make_blocks(5, 1, N)                         # ~3.5M blocks tiling the index
    rng = np.array([0, 0.334, 0.667, 1], np.float32)
    acc = ctx.alloc(N * 8)
    acc_host = np.zeros((N, 2), np.int64); bins = np.zeros((blocks.shape[0], 3), np.int64)
    done = 0; first = True; pat = b""
    while done < total:
        n = min(per, total - done)
        pat = synth.make_pat_text_fast(100 + done // per, n, N, chrom="chr1")
        mc, hb = check_pat_sharding(E, pat, 1, N + 1, blocks, rng, 3)
        P = ctx.pats_from_text(pat)
        ctx.pat2beta(P, 1, N + 1, meth_cov=acc, zero_first=first)     # the device accumulator a multi-batch run keeps
        P.free()
        acc_host += mc; bins += hb
        first = False; done += n
    np.testing.assert_array_equal(acc.to_host(np.int32).reshape(N, 2).astype(np.int64), acc_host)
    acc.free()
    assert int(acc_host[:, 1].sum()) > 3 * total and int(bins.sum()) > total // 4
    assert (acc_host[:, 0] <= acc_host[:, 1]).all()
    # the reference executables / port on a bounded sample at the far end of the index (indices ~28M), same blocks there
    lo = N - 400_000
    tail = pat[pat.rindex(b"\n", 0, len(pat) - 4_000_000) + 1:] if len(pat) > 8_000_000 else pat
    t0 = int(tail.split(b"\t", 2)[1])
    sub = blocks[(blocks[:, 0] >= max(lo, t0 + 64))]
    got = E.homog(tail, blocks, rng, 3)[blocks.shape[0] - sub.shape[0]:]
    np.testing.assert_array_equal(got, oracle.port_homog(tail, sub, rng, 3))
    np.testing.assert_array_equal(E.p2b(tail, lo, N + 1), oracle.port_pat2beta(tail, lo, N + 1))
