"""Host-side mirrors of the reference's Python (segment.py stitching, homog.py thresholds / scaling) against golden
vectors produced by the reference's own modules (tests/golden/make_golden.py).  CPU: the DP solver is the oracle port;
the `-m gpu` variant runs the same flow on wgbs_segment."""
import json
import os

import numpy as np
import pytest

from wgbs_tools_b200 import homog as hg
from wgbs_tools_b200 import segment as sg
from wgbs_tools_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _cases():
    z = np.load(os.path.join(GOLD, "segment_stitch.npz"))
    n = len({k.split("_")[0] for k in z.files})
    return [{k.split("_", 1)[1]: z[k] for k in z.files if k.startswith(f"c{i}_")} for i in range(n)]


def _inputs(c):
    seed, N, K = int(c["seed"]), int(c["N"]), int(c["K"])
    betas = synth.make_betas(100 + seed, K, N)
    g = synth.make_genome(200 + seed, "chr1", 1_000_000, with_bases=False)
    return betas, g.loci[:N]


@pytest.mark.parametrize("i", range(5))
def test_segment_stitching_matches_reference_golden_cpu(oracle, i):
    c = _cases()[i]
    betas, d = _inputs(c)
    max_cpg, max_bp, ps = int(c["max_cpg"]), int(c["max_bp"]), float(c["ps"])

    def solve(sites):
        return [oracle.port_segment([x[s - 1:e - 1] for x in betas], d[s - 1:e - 1], max_cpg, max_bp, ps) + s for s, e in sites]

    blocks = sg.segment_regions([(1, int(c["N"]) + 1)], solve, int(c["chunk"]))
    merged = np.concatenate([blocks[:, 0], blocks[-1:, 1]])
    np.testing.assert_array_equal(merged, c["merged"])


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(5))
def test_segment_stitching_matches_reference_golden_gpu(ctx, i):
    c = _cases()[i]
    betas, d = _inputs(c)
    solver = sg.GpuSolver(ctx, betas, d, int(c["max_cpg"]), int(c["max_bp"]), float(c["ps"]))
    blocks = sg.segment_regions([(1, int(c["N"]) + 1)], solver, int(c["chunk"]))
    solver.close()
    merged = np.concatenate([blocks[:, 0], blocks[-1:, 1]])
    np.testing.assert_array_equal(merged, c["merged"])


def test_break_to_chunks_and_filters():
    tags, s, e = sg.break_to_chunks([(1, 150_001), (200_000, 200_010)], 60_000)
    assert s == [1, 60_001, 120_001, 200_000] and e == [60_001, 120_001, 150_001, 200_010]
    assert tags == ["1-150001"] * 3 + ["200000-200010"]
    assert sg.effective_max_cpg(1000, 2000) == 1000 and sg.effective_max_cpg(5000, 2000) == 1000
    b = np.array([[1, 2], [2, 5], [5, 6]])
    np.testing.assert_array_equal(sg.filter_min_cpg(b, 2), [[2, 5]])
    assert sg.increase_patch(50, 60) == 60 and sg.increase_patch(60, 60) == 61 and sg.increase_patch(20, 100) == 40


def test_homog_host_logic_matches_reference_golden():
    G = json.load(open(os.path.join(GOLD, "homog_host.json")))
    for l, s in G["rates"].items():
        got, edges = hg.rate_edges(int(l))
        assert got == s
        assert edges.dtype == np.float32 and edges[0] == 0 and edges[-1] == 1
    assert hg.rate_edges(3)[0] == "0,0.334,0.667,1"
    assert hg.rate_edges(4, "0.25,0.75")[0] == "0,0.25,0.75,1"
    d = np.array(G["trim_uxm"]["in"])
    np.testing.assert_array_equal(hg.trim_uxm_to_uint8(d, 8), G["trim_uxm"]["u8"])
    np.testing.assert_array_equal(hg.trim_uxm_to_uint8(d * 40, 16), G["trim_uxm"]["u16"])
    with pytest.raises(hg.IllegalArgumentError):
        hg.parse_range("0,0.5,0.4,1")
    np.testing.assert_array_equal(hg.parse_range("0,0.334,0.667,1"), np.array([0, 0.334, 0.667, 1], np.float32))


def test_homog_block_order_mirrors_sort_and_wrapper():
    lines = [b"chr1\t10\t20\t7\t9\n", b"chr1\t1\t5\t2\t4\n", b"chr1\t10\t30\t7\t8\n", b"chr1\t6\t9\t5\t7\n"]
    order = hg.sort_blocks_order(lines)
    assert order.tolist() == [1, 3, 2, 0]                       # -k4,4n then -k5,5n
    starts = np.array([7, 2, 7, 5])
    counts_sorted = np.arange(12).reshape(4, 3)
    back = hg.restore_block_order(counts_sorted, starts)
    # homog.py:113-118 restores with a stable argsort of startCpG only: ties keep file order (reference behaviour)
    assert back[1].tolist() == [0, 1, 2] and back[3].tolist() == [3, 4, 5]
    assert not hg.blocks_are_sorted(starts, [9, 4, 8, 7]) and hg.blocks_are_sorted([1, 2, 2], [5, 3, 4])


@pytest.mark.gpu
def test_trim_golden_on_gpu(ctx):
    G = json.load(open(os.path.join(GOLD, "homog_host.json")))
    mc = np.array(G["trim_beta"]["in"], np.int64)
    ok = mc.max(axis=1) < 2 ** 31
    np.testing.assert_array_equal(ctx.trim(mc[ok].astype(np.int32), int(ok.sum()), 8), np.array(G["trim_beta"]["u8"])[ok])
    np.testing.assert_array_equal(ctx.trim(mc[ok].astype(np.int32), int(ok.sum()), 16), np.array(G["trim_beta"]["u16"])[ok])


def test_init_genome_matches_reference_definition(tmp_path):
    """CpG dictionary from a FASTA: reference init_genome.py:246-281 (re.finditer('CG') on the upper-cased sequence,
    chromosome_order, is_valid_chrome) restated here as the expectation"""
    import gzip
    import re
    from wgbs_tools_b200 import init_genome as ig
    from wgbs_tools_b200.genome import GenomeRef
    rng = np.random.default_rng(0)
    seqs = {}
    for name, n in [("chr10", 5000), ("chr2", 7001), ("chrX", 3000), ("chrUn_gl0001", 800), ("chrM", 1200), ("chr1", 4000)]:
        s = bytes(rng.choice(list(b"ACGTacgtN"), size=n, p=[.2, .2, .2, .2, .04, .04, .04, .04, .04]).tolist())
        seqs[name] = s
    fa = tmp_path / "g.fa"
    with open(fa, "wb") as f:
        for k, v in seqs.items():
            f.write(b">" + k.encode() + b" some description\n")
            for i in range(0, len(v), 60):
                f.write(v[i:i + 60] + b"\n")
    out = tmp_path / "ref"
    r = ig.init_genome(str(fa), str(out))
    assert r["chroms"] == ["chr1", "chr2", "chr10", "chrX", "chrM"]               # chrUn dropped, reference order
    exp = []; idx = 1
    for c in r["chroms"]:
        for m in re.finditer("CG", seqs[c].decode().upper()):
            exp.append(f"{c}\t{m.start() + 1}\t{idx}\n"); idx += 1
    assert gzip.open(out / "CpG.bed.gz", "rt").read() == "".join(exp)
    assert (out / "chrome.size").read_text().splitlines()[0] == "chr1\t4000"
    ref = GenomeRef(str(out))
    assert ref.nr_sites == idx - 1 and ref.chrom_of_site(1) == "chr1" and ref.chrom_of_site(idx - 1) == "chrM"
    assert ig.chromosome_order("chrY") == 10001 and not ig.is_valid_chrome("chr1_random")


def test_csi_index_round_trip_region_queries(tmp_path):
    """.pat.gz = `cat` of per-chromosome BGZF parts (one EOF block each) -> CSI (min_shift 12, 9 levels, tabix aux) ->
    every region query through the index returns exactly what a full scan returns; the index survives save/load"""
    from wgbs_tools_b200 import csi, synth
    from wgbs_tools_b200.patio import bgzf_compress
    sizes = (("chr1", 300_000), ("chr2", 120_000), ("chrX", 4_000))
    parts, full, first = [], [], 1
    for ci, (c, n) in enumerate(sizes):
        idx, pats, cnt = synth.make_pat_records(ci + 1, n // 4, n, first_idx=first, mean_len=5)
        t = synth.pat_text(c, idx, pats, cnt); full.append(t); parts.append(bgzf_compress(t, 2)); first += n
    p = tmp_path / "x.pat.gz"; p.write_bytes(b"".join(parts))
    out = csi.index_pat(str(p))
    assert out == str(p) + ".csi"
    ix = csi.CsiIndex.load(out)
    assert (ix.min_shift, ix.n_lvls, ix.names, ix.conf) == (12, 9, ["chr1", "chr2", "chrX"], (0, 1, 2, 2, ord("#"), 0))
    assert ix.to_bytes() == open(out, "rb").read()
    for t, (c, n) in enumerate(sizes):
        (beg, end), (nrec, _) = ix.bins[t][ix.meta_bin]["chunks"]
        assert nrec == full[t].count(b"\n") and beg < end
        for bn, rec in ix.bins[t].items():
            if bn != ix.meta_bin:
                assert bn <= ((1 << 30) - 1) // 7 and all(u < v for u, v in rec["chunks"]) and beg <= rec["chunks"][0][0] and rec["chunks"][-1][1] <= end
    recs = [(l.split(b"\t")[0].decode(), int(l.split(b"\t")[1]), l) for l in b"".join(full).splitlines(keepends=True)]
    rng = np.random.default_rng(0)
    base = {"chr1": 1, "chr2": 300_001, "chrX": 420_001}
    for _ in range(200):
        c, n = sizes[rng.integers(0, 3)]
        lo = int(base[c] + rng.integers(0, n)); hi = int(lo + rng.integers(0, (1, 10, 1000, 50_000)[rng.integers(0, 4)]))
        assert csi.read_region(str(p), c, lo, hi, ix) == b"".join(l for cc, i, l in recs if cc == c and lo <= i <= hi), (c, lo, hi)
    assert csi.read_region(str(p), "chr9", 1, 10, ix) == b"" and csi.read_region(str(p), "chr2", 1, 10 ** 9, ix) == full[1]
    # the same file and index through an INDEPENDENT reader written from the BGZF / CSI / tabix specifications (tests/csi_spec_reader.py:
    # no code shared with csi.py; htslib's tabix is not available here): structure validates, and region queries agree with the scan
    import csi_spec_reader as spec
    assert spec.validate_bgzf(p.read_bytes()) > 3
    six = spec.Csi(out)
    assert (six.min_shift, six.depth, six.names) == (12, 9, ["chr1", "chr2", "chrX"])
    assert (six.format, six.col_seq, six.col_beg, six.col_end, six.meta, six.skip) == (0, 1, 2, 2, ord("#"), 0)       # tabix -b 2 -e 2 (generic format)
    for _ in range(120):
        c, n = sizes[rng.integers(0, 3)]
        lo = int(base[c] + rng.integers(0, n)); hi = int(lo + rng.integers(0, (1, 10, 1000, 50_000)[rng.integers(0, 4)]))
        assert spec.query(str(p), six, c, lo, hi) == b"".join(l for cc, i, l in recs if cc == c and lo <= i <= hi), (c, lo, hi)
    assert spec.query(str(p), six, "chrX", 420_001, 424_001) == full[2]
    # hts_reg2bin / reg2bins at this depth: a leaf bin is 4096 sites wide; a range is always inside the bins listed for it
    assert csi.reg2bin(np.array([0]), np.array([1]), 12, 9)[0] == ((1 << 27) - 1) // 7
    for b, e in ((0, 1), (4095, 4097), (5_000_000, 5_000_001), (123, 9_999_999)):
        assert int(csi.reg2bin(np.array([b]), np.array([e]), 12, 9)[0]) in csi.reg2bins(b, e, 12, 9)
    # an empty pat file still gives a loadable index
    q = tmp_path / "e.pat.gz"; q.write_bytes(bgzf_compress(b"", 1))
    assert csi.CsiIndex.load(csi.index_pat(str(q))).names == []


def test_pat_pieces_cut_at_line_ends():
    """patio.pat_pieces (pat2beta / homog on pat files larger than one call): pieces end at line ends, cover the text, fit the limit"""
    from wgbs_tools_b200.patio import pat_pieces
    txt = synth.make_pat_text_fast(3, 20_000, 50_000)
    assert list(pat_pieces(None, txt, 1 << 30)) == [txt]
    for limit in (200, 4096, 100_000):
        ps = [bytes(p) for p in pat_pieces(None, txt, limit)]
        assert b"".join(ps) == txt and all(len(p) <= limit and p.endswith(b"\n") for p in ps) and len(ps) > 1
    ps = [bytes(p) for p in pat_pieces(None, txt[:-1], 4096)]                  # no newline at the end of the text
    assert b"".join(ps) == txt[:-1]
    with pytest.raises(ValueError):
        list(pat_pieces(None, b"chr1\t5\t" + b"C" * 500 + b"\t1\n" * 3, 100))


def test_bench_reference_arm_prints_the_contract_line(oracle):
    """`bench.py --impl reference` (the reference's own executables on the host cores): ONE JSON line with the contract's keys"""
    import subprocess
    import sys
    if not oracle.have_ref():
        pytest.skip("reference executables not built")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--reads", "12000", "--steps", "1", "--warmup", "0"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "bam2pat_reads_per_sec" and d["unit"] == "reads/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1 and d["dtype"] == "u8" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"] and "dictionary" in d["cpu_baseline"]["sample"] and d["cpu_baseline"]["built_with_O2"] is True
    assert d["e2e"] == {"value": d["value"], "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_bench_roofline_object_from_a_profile():
    """bench.build_roofline: kernel names come from the library's profiler possibly WITH template arguments (sam_lines_k<2>); the
    dominant kernel is the largest share of the step among those with defined algorithmic bytes"""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import bench
    w = {"n_rec": 994_890, "text_bytes": 356_666_308, "n_tmpl": 494_264, "seq_end_avg": 181.0, "out_text_bytes": 10_200_000}
    rep = {"nl_scan_k": (3, 0.388), "sam_lines_k<2>": (3, 0.336), "rs_onesweep_k": (12, 0.299), "pileup_call_k": (3, 0.283), "scan_lookback_k<OutT>": (12, 0.172),
           "merge_templates_k": (3, 0.188), "pileup_measure_k": (3, 0.182), "pair_resolve_k": (3, 0.167), "line_write_k": (3, 0.163), "pat2beta_k": (3, 0.1), "tiny_k": (3, 0.01)}
    r = bench.build_roofline(rep, 3, w, 6554.9, "measured")
    assert r["kernel"] == "nl_scan_k" and r["bound"] == "hbm" and r["unit"] == "GB/s"
    assert abs(r["achieved"] - (356_666_308 + 4 * 994_890) / (0.388 / 3 / 1e3) / 1e9) < 1e-6 and abs(r["frac"] - r["achieved"] / 6554.9) < 1e-12
    assert r["traffic"] == json.load(open(os.path.join(root, "profiles", "ncu_traffic.json")))["nl_scan_k"]       # keyed by the bare kernel name
    names = [k["kernel"] for k in r["per_kernel"]]
    assert "sam_lines_k<2>" in names and "merge_templates_k" in names and "scan_lookback_k<OutT>" not in names
    assert abs(sum(k["share_of_step"] for k in r["per_kernel"]) - (sum(v[1] for k, v in rep.items() if k in names) / sum(v[1] for v in rep.values()))) < 1e-9
    # a profile whose top kernel has no byte model: the next one with a model is reported
    r2 = bench.build_roofline({"mystery_k": (3, 9.0), **rep}, 3, w, 6554.9, "measured")
    assert r2["kernel"] == "nl_scan_k" and list(r2["breakdown_ms_per_step"])[0] == "mystery_k"


def _beta_table_cases(tmp_path):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    # the same seeded inputs the golden script fed to the reference (tests/golden/make_golden.py beta_table_inputs), rebuilt here
    N = 20_000
    betas = synth.make_betas(31, 5, N)
    betas[3][:, 1] = np.minimum(betas[3][:, 1], 1)
    betas[3][:, 0] = np.minimum(betas[3][:, 0], betas[3][:, 1])
    paths = []
    for i, b in enumerate(betas):
        p = str(tmp_path / f"s{i}.beta"); b.tofile(p); paths.append(p)
    blocks = synth.make_blocks(7, 1, N, mean_len=6)[:900]
    rows = []
    for k, (a, b) in enumerate(blocks.tolist()):
        if k % 53 == 5:
            rows.append(f"chr1\t{1000 + k}\t{1000 + k + 1}\tNA\tNA\n")
        rows.append(f"chr1\t{10 * a}\t{10 * b}\t{a}\t{b}\n")
    bp = str(tmp_path / "blocks.bed"); open(bp, "w").write("".join(rows))
    gp = str(tmp_path / "groups.csv")
    open(gp, "w").write("name,group,include\ns0,A,True\ns1,B,True\ns2,A,True\ns3,B,True\ns4,A,False\n")
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "beta_to_table.json")))
    return paths, bp, gp, gold


def check_beta_table(ctx, tmp_path):
    from wgbs_tools_b200 import beta_to_table as b2t
    paths, bp, gp, gold = _beta_table_cases(tmp_path)
    for c in gold:
        paths_c = paths if not c["groups"] else paths          # (the groups file leaves s4 out by include == False)
        txt = b2t.table_text(ctx, bp, paths_c, gp if c["groups"] else None, c["min_cov"], c["digits"], c["chunk"])
        assert txt == c["text"], (c["groups"], c["min_cov"], c["digits"])


def test_beta_to_table_host_logic_matches_reference_golden(tmp_path):
    """beta_to_table.py (groups file, NA blocks, min_cov, digits, chunked printing) against the text the REFERENCE's beta_to_table.py
    printed for the same seeded inputs (tests/golden/beta_to_table.json); the per-block sums come from numpy here (no GPU)"""
    class NumpyCtx:                                             # test stand-in for Context.beta_to_blocks: np.add.reduceat like the reference
        def beta_to_blocks(self, data, bs, be, out_bits, want_sums=False):
            c = np.concatenate([np.zeros((1, 2), np.int64), np.cumsum(data.astype(np.int64), axis=0)])
            sums = c[np.asarray(be) - 1] - c[np.asarray(bs) - 1]
            return None, sums
    check_beta_table(NumpyCtx(), tmp_path)


def test_pat_shards_of_a_bgzf_file_partition_its_lines(tmp_path, monkeypatch):
    """patio.read_pat_device_shard (pat2beta / homog under torchrun): every rank inflates only its run of BGZF blocks (+ one neighbour
    on each side) and owns the lines that START inside its run -- the shares must partition the file's lines for any world size,
    also when a block boundary falls exactly on a line end.  The device is stood in for by host memory (no GPU here)."""
    import ctypes
    import gzip

    import numpy as np

    from wgbs_tools_b200 import _lib, patio, synth

    class FakeBuf:
        def __init__(self, data):
            self.a = np.frombuffer(bytearray(data), np.uint8); self.ptr = self.a.ctypes.data; self.nbytes = self.a.size

        def __len__(self):
            return self.nbytes

        def free(self):
            pass

    class FakeCtx:
        h = None

        def bgzf_inflate(self, raw):
            return FakeBuf(gzip.decompress(raw))

    monkeypatch.setattr(_lib.lib, "wgbs_memcpy", lambda h, dst, src, n: ctypes.memmove(dst, src, n) and 0)
    idx, pats, cnt = synth.make_pat_records(3, 4000, 9000)
    text = synth.pat_text("chr1", idx, pats, cnt)
    lines = text.splitlines(keepends=True)
    # blocks of 4096 bytes, and one layout whose block ends coincide with line ends
    layouts = [[text[i:i + 4096] for i in range(0, len(text), 4096)], []]
    acc = b""
    for l in lines:
        acc += l
        if len(acc) > 3000:
            layouts[1].append(acc); acc = b""
    if acc:
        layouts[1].append(acc)
    for k, chunks in enumerate(layouts):
        p = tmp_path / f"t{k}.pat.gz"
        p.write_bytes(b"".join(patio._bgzf_block(c) for c in chunks) + patio.BGZF_EOF)
        for world in (1, 2, 3, 5, 16):
            got = []
            for rank in range(world):
                buf, view = patio.read_pat_device_shard(FakeCtx(), str(p), rank, world)
                got.append(ctypes.string_at(view.ptr, view.nbytes))
            assert b"".join(got) == text, (k, world)
            assert all(g == b"" or g.endswith(b"\n") for g in got)
    q = tmp_path / "plain.pat.gz"; q.write_bytes(gzip.compress(text))
    assert patio.read_pat_device_shard(FakeCtx(), str(q), 0, 2) is None          # plain gzip: the caller shards the host text
