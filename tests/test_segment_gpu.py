"""GPU parity of the segmenter: bit-exact glibc log2f/log2 ports and border-identical DP vs the reference `segmentor`
executable (oracle/_ref) / its C restatement, run on the same host (same glibc)."""
import ctypes
import os

import numpy as np
import pytest

from wgbs_tools_b200 import synth

pytestmark = pytest.mark.gpu


def _host_logs(H, p):
    L = H.port()
    a = np.empty(p.size, np.float32); b = np.empty(p.size, np.float64)
    L.port_log2f_array(p.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(p.size), a.ctypes.data_as(ctypes.c_void_p))
    L.port_log2_1m_array(p.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(p.size), b.ctypes.data_as(ctypes.c_void_p))
    return a, b


def test_glibc_log2_ports_bit_exact(ctx, oracle):
    """every float p in (0,1) is a possible argument (segmentor.cpp:127-134); default: every exponent, a 1/64 stride of
    mantissas plus the neighbourhoods of 0, 1 and the near-1 branch boundary.  WGBS_EXHAUSTIVE=1 sweeps all 2^30."""
    H = oracle
    stride = 1 if os.environ.get("WGBS_EXHAUSTIVE") else 64
    edge = np.concatenate([np.arange(1, 70000, dtype=np.uint32), np.arange(0x3f800000 - 3_000_000, 0x3f800000, dtype=np.uint32),
                           np.arange(0x00800000 - 100, 0x00800000 + 100, dtype=np.uint32)])
    n_bad = 0
    blocks = [edge] + [np.arange(lo, min(lo + (1 << 26), 0x3f800000), stride, dtype=np.uint32) for lo in range(0x00800000, 0x3f800000, 1 << 26)]
    for bits in blocks:
        p = bits.view(np.float32)
        ga, gb = ctx.glibc_log2_probe(p)
        ha, hb = _host_logs(H, p)
        n_bad += int((ga.view(np.uint32) != ha.view(np.uint32)).sum()) + int((gb.view(np.uint64) != hb.view(np.uint64)).sum())
    assert n_bad == 0


@pytest.mark.parametrize("K,n,max_cpg,max_bp,ps", [(6, 2500, 200, 2000, 15), (3, 4000, 1000, 5000, 1), (10, 1500, 50, 300, 0.5), (1, 800, 800, 10 ** 9, 15)])
def test_segment_matches_reference(ctx, oracle, K, n, max_cpg, max_bp, ps):
    H = oracle
    betas = synth.make_betas(4 + K, K, n)
    g = synth.make_genome(2, "chr1", 600_000, with_bases=False)
    d = g.loci[:n]
    got = ctx.segment(betas, d, [(0, n)], max_cpg, max_bp, ps)[0]
    np.testing.assert_array_equal(got, H.port_segment(betas, d, max_cpg, max_bp, ps))
    if H.have_ref():
        paths = [H.write_tmp(b.tobytes(), f".{i}.beta") for i, b in enumerate(betas)]
        np.testing.assert_array_equal(got, H.ref_segmentor(paths, 0, n, max_cpg, max_bp, ps, d))
        for p in paths:
            os.remove(p)
    assert got[0] == 0 and got[-1] == n and got.size > 10


def test_segment_many_chunks_and_patches_in_one_call(ctx, oracle):
    """chunks (segment.py:129-134) and overlapping stitching patches (segment.py:209-227) batched in one call, with
    zero-coverage stretches and a chunk of a single site"""
    H = oracle
    K, N = 5, 12_000
    betas = synth.make_betas(11, K, N)
    for b in betas:
        b[3000:3300] = 0                                 # uncovered stretch: nt == 0 -> datasets skipped
    g = synth.make_genome(3, "chr1", 2_000_000, with_bases=False)
    d = g.loci[:N]
    chunks = [(0, 5000), (5000, 5000), (10000, 2000), (4950, 100), (9900, 200), (7, 1)]
    got = ctx.segment(betas, d, chunks, 300, 2000, 15)
    for (s, n), b in zip(chunks, got):
        ref = H.port_segment([x[s:s + n] for x in betas], d[s:s + n], 300, 2000, 15)
        np.testing.assert_array_equal(b, ref)


def test_segment_rejects_bad_beta(ctx):
    from wgbs_tools_b200._lib import WgbsError
    b = synth.make_betas(1, 2, 500)
    b[1][100] = (9, 3)                                   # meth > cover
    with pytest.raises(WgbsError, match="meth > cover"):
        ctx.segment(b, np.arange(500, dtype=np.uint32) * 50, [(0, 500)], 100, 2000, 15)


def test_segment_many_chunks_in_one_call(ctx, oracle):
    """waves are packed by the real number of cost cells of every chunk (seg_chunk_cells_k): 40 chunks in one call give what the
    same chunks give one call each, and what the oracle gives"""
    K, n, nch = 4, 3000, 40
    betas = synth.make_betas(3, K, n * nch)
    loci = synth.make_genome(2, "chr1", n * nch * 120, with_bases=False).loci[:n * nch]
    chunks = [(s, n) for s in range(0, n * nch, n)]
    b = ctx.segment(betas, loci, chunks, 1000, 2000, 15)
    assert len(b) == nch
    for c in (0, 7, 39):
        one = ctx.segment(betas, loci, [chunks[c]], 1000, 2000, 15)[0]
        np.testing.assert_array_equal(b[c], one)
    np.testing.assert_array_equal(b[7], oracle.port_segment([x[7 * n:8 * n] for x in betas], loci[7 * n:8 * n], 1000, 2000, 15))


@pytest.mark.parametrize("K,n,max_cpg,max_bp,ps", [(6, 2500, 200, 2000, 15), (3, 4000, 1000, 5000, 1), (1, 800, 800, 10 ** 9, 15)])
def test_segment_windows_wider_than_a_warp(ctx, oracle, K, n, max_cpg, max_bp, ps):
    """borders == oracle, including windows wider than a warp (max_cpg 800 / 1000 with a huge max_bp)"""
    betas = synth.make_betas(11, K, n)
    loci = synth.make_genome(2, "chr1", n * 150, with_bases=False).loci[:n]
    b = ctx.segment(betas, loci, [(0, n)], max_cpg, max_bp, ps)[0]
    np.testing.assert_array_equal(b, oracle.port_segment(betas, loci, max_cpg, max_bp, ps))
