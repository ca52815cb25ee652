"""End-to-end through the CLI mirrors (bam2pat, pat2beta, homog, segment) on a small 2-chromosome synthetic genome:
file-level outputs against the reference executables / oracle."""
import gzip
import os

import numpy as np
import pytest

from wgbs_tools_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def world(tmp_path_factory, oracle):
    H = oracle
    d = tmp_path_factory.mktemp("ref")
    g1 = synth.make_genome(31, "chr1", 600_000, first_idx=1)
    g2 = synth.make_genome(32, "chr2", 400_000, first_idx=1 + g1.n_cpg)
    refdir = d / "synth"
    refdir.mkdir()
    with gzip.open(refdir / "CpG.bed.gz", "wb") as f:
        f.write(g1.dict_text() + g2.dict_text())
    (refdir / "CpG.chrome.size").write_text(f"chr1\t{g1.n_cpg}\nchr2\t{g2.n_cpg}\n")
    sam1 = synth.make_sam(g1, 12_000, 1, paired=True, name_prefix="a")
    sam2 = synth.make_sam(g2, 8_000, 2, paired=True, name_prefix="b")
    samp = d / "sample.sam"
    samp.write_bytes(b"@HD\tVN:1.6\tSO:coordinate\n" + sam1 + sam2)
    # reference pipeline per chromosome (bam2pat.py:144-209 + 88-111), parts concatenated in chromosome order
    parts = []
    for g, s in ((g1, sam1), (g2, sam2)):
        dp = H.write_tmp(g.dict_text(), ".CpG.bed")
        if H.have_ref():
            out, _ = H.ref_patter(s, dp, g.chrom, True)
            parts.append(H.ref_collapse(out))
        else:
            out, _ = H.port_patter(H.port_match_maker(s), g.loci, g.idx())
            parts.append(H.port_collapse(out))
    return dict(dir=d, refdir=str(refdir), sam=str(samp), pat=b"".join(parts), N=g1.n_cpg + g2.n_cpg, g1=g1, g2=g2, H=H)


def test_bam2pat_cli_writes_reference_identical_pat_and_beta(world):
    from wgbs_tools_b200 import bam2pat
    w = world; H = w["H"]
    out = w["dir"] / "out"; out.mkdir()
    bam2pat.main([w["sam"], "--genome", w["refdir"], "-o", str(out), "-q", "10"])
    raw = (out / "sample.pat.gz").read_bytes()
    assert gzip.decompress(raw) == w["pat"]
    # BGZF structure: every member has the BC extra field; one EOF block per chromosome part (cat of parts)
    from wgbs_tools_b200.patio import BGZF_EOF
    assert raw.count(BGZF_EOF) == 2 and raw.endswith(BGZF_EOF) and raw[:4] == b"\x1f\x8b\x08\x04" and raw[12:14] == b"BC"
    beta = np.fromfile(out / "sample.beta", np.uint8).reshape(-1, 2)
    ref_counts = H.ref_stdin2beta(w["pat"], 1, w["N"] + 1) if H.have_ref() else H.port_pat2beta(w["pat"], 1, w["N"] + 1)
    assert beta.tobytes() == H.ref_trim(ref_counts).tobytes()
    # the CSI index written next to it (Indxer, bam2pat.py:415): region reads through it == a scan of the text
    from wgbs_tools_b200 import csi
    ix = csi.CsiIndex.load(str(out / "sample.pat.gz.csi"))
    assert ix.names == ["chr1", "chr2"]
    lo, hi = w["g1"].n_cpg + 500, w["g1"].n_cpg + 900
    exp = b"".join(l for l in w["pat"].splitlines(keepends=True) if l.startswith(b"chr2\t") and lo <= int(l.split(b"\t")[1]) <= hi)
    assert exp and csi.read_region(str(out / "sample.pat.gz"), "chr2", lo, hi, ix) == exp
    w["out"] = out


def test_pat2beta_cli_from_pat_gz(world):
    from wgbs_tools_b200 import pat2beta
    w = world; H = w["H"]
    out = w["dir"] / "out2"; out.mkdir()
    pg = w["dir"] / "x.pat.gz"
    pg.write_bytes(gzip.compress(w["pat"]))
    pat2beta.main([str(pg), "--genome", w["refdir"], "-o", str(out), "-f"])
    pat2beta.main([str(pg), "--genome", w["refdir"], "-o", str(out), "-f", "-l"])
    counts = H.port_pat2beta(w["pat"], 1, w["N"] + 1)
    assert (out / "x.beta").read_bytes() == H.ref_trim(counts).tobytes()
    assert (out / "x.lbeta").read_bytes() == H.ref_trim(counts, lbeta=True).tobytes()


def test_homog_cli_sorted_and_unsorted_blocks(world):
    from wgbs_tools_b200 import homog
    w = world; H = w["H"]
    pg = w["dir"] / "h.pat.gz"; pg.write_bytes(gzip.compress(w["pat"]))
    blocks = synth.make_blocks(5, 1, w["N"], mean_len=10)
    loci = np.concatenate([w["g1"].loci, w["g2"].loci])
    lines = synth.blocks_text("chrX", blocks, loci).splitlines(keepends=True)
    for tag, ls in (("sorted", lines), ("shuffled", [lines[i] for i in np.random.default_rng(0).permutation(len(lines))])):
        bp = w["dir"] / f"blocks_{tag}.bed"; bp.write_bytes(b"".join(ls))
        homog.main([str(pg), "-b", str(bp), "-p", str(w["dir"] / f"hom_{tag}"), "-f", "-l", "3", "--genome", w["refdir"]])     # (<= 5000 blocks: the cview route)
        got = np.array([l.split(b"\t")[5:8] for l in gzip.open(str(w["dir"] / f"hom_{tag}.uxm.bed.gz")).read().splitlines()], dtype=np.int64)
        bl = np.array([[int(l.split(b"\t")[3]), int(l.split(b"\t")[4])] for l in ls])
        order = np.lexsort((bl[:, 1], bl[:, 0]))
        ref_sorted = H.port_homog(w["pat"], bl[order], np.array([0, 0.334, 0.667, 1], np.float32), 3)
        exp = np.empty_like(ref_sorted); exp[order] = ref_sorted
        np.testing.assert_array_equal(got, exp)
        if H.have_ref():
            refc = H.ref_homog(w["pat"], str(bp), "0,0.334,0.667,1", 3, sort_blocks=(tag == "shuffled"))
            np.testing.assert_array_equal(got[order] if tag == "shuffled" else got, refc)
    homog.main([str(pg), "-b", str(w["dir"] / "blocks_sorted.bed"), "-p", str(w["dir"] / "hom_bin"), "-f", "--binary", "--genome", w["refdir"]])
    assert (w["dir"] / "hom_bin.uxm").stat().st_size == 3 * len(lines)


def test_segment_cli_blocks(world):
    from wgbs_tools_b200 import segment as sg
    w = world; H = w["H"]
    N = w["N"]
    betas = synth.make_betas(77, 4, N)
    paths = []
    for i, b in enumerate(betas):
        p = w["dir"] / f"s{i}.beta"; b.tofile(p); paths.append(str(p))
    outp = w["dir"] / "blocks.bed"
    sg.main(["--betas", *paths, "--genome", w["refdir"], "-c", "1500", "--max_bp", "1500", "--min_cpg", "3", "-o", str(outp)])
    rows = [l.split("\t") for l in outp.read_text().splitlines()]
    loci = np.concatenate([w["g1"].loci, w["g2"].loci]).astype(np.int64)
    n1 = w["g1"].n_cpg
    # expected: same host logic with the oracle as the DP solver, chromosome by chromosome
    def solve(sites):
        return [H.port_segment([x[s - 1:e - 1] for x in betas], loci[s - 1:e - 1], 750, 1500, 15) + s for s, e in sites]
    exp = sg.filter_min_cpg(sg.segment_regions([(1, n1 + 1), (n1 + 1, N + 1)], solve, 1500), 3)
    got = np.array([[int(r[3]), int(r[4])] for r in rows])
    np.testing.assert_array_equal(got, exp)
    for r in rows[:50] + rows[-50:]:
        s, e = int(r[3]), int(r[4])
        assert r[0] == ("chr1" if s <= n1 else "chr2") and int(r[1]) == loci[s - 1] and int(r[2]) == loci[e - 2] + 1


def test_view_cli_region_sites_bed_and_whole_file(world, tmp_path):
    """`wgbstools view X.pat.gz [-r|-s|-L] [--strict --strip --no_gaps --min_len]` against
    [tabix |] cview | [sort |] collapse_pat.pl with the reference cview executable (tabix's selection restated)"""
    from wgbs_tools_b200 import view
    w = world; H = w["H"]
    if not H.have_cview():
        pytest.skip("reference cview not built")
    pat = w["pat"]; N = w["N"]; g1, g2 = w["g1"], w["g2"]
    pg = w["dir"] / "v.pat.gz"; pg.write_bytes(gzip.compress(pat))
    lines = pat.splitlines(keepends=True)
    idx = np.array([int(l.split(b"\t")[1]) for l in lines]); chrom = np.array([l.split(b"\t")[0] for l in lines])

    def run(*argv):
        o = tmp_path / "view.out"
        view.main([str(pg), "--genome", w["refdir"], "-o", str(o), *argv])
        return o.read_bytes()

    flagsets = [([], {}), (["--strict"], dict(strict=True)), (["--strict", "--strip", "--min_len", "3"], dict(strict=True, strip=True, min_cpgs=3)),
                (["--no_gaps", "--strip"], dict(no_gaps=True, strip=True))]
    # whole file: gunzip -c | cview --sites "1\tN+1" | collapse_pat.pl (no sort)
    for argv, kw in flagsets:
        assert run(*argv) == H.port_collapse_pat(H.ref_cview(pat, sites=(1, N + 1), **kw)), argv
    # -r on chr2 and -s: tabix pat chr:(s-150)-(e-1) | cview --sites | sort | collapse
    lo, hi = int(g2.loci[2000]), int(g2.loci[2300])
    s = g2.first_idx + 2000; e = g2.first_idx + 2300                     # the CpG sitting on `hi` is excluded
    for sel in (["-r", f"chr2:{lo}-{hi}"], ["-s", f"{s}-{e}"]):
        m = (chrom == b"chr2") & (idx >= max(1, s - 150, g2.first_idx)) & (idx <= e - 1)
        sub = b"".join(l for l, k in zip(lines, m) if k)
        for argv, kw in flagsets:
            exp = H.port_collapse_pat(H.sort_pat(H.ref_cview(sub, sites=(s, e), **kw)))
            assert exp and run(*sel, *argv) == exp, (sel, argv)
        assert run(*sel, "--no_sort", "--strip") == H.port_collapse_pat(H.ref_cview(sub, sites=(s, e), strip=True))
    # the same region read through a CSI index (BGZF pat + `wgbstools index`): only the indexed blocks are inflated
    from wgbs_tools_b200 import index as windex
    from wgbs_tools_b200.patio import bgzf_compress
    pz = tmp_path / "vi.pat.gz"; pz.write_bytes(bgzf_compress(pat, 2)); windex.main([str(pz)])
    o2 = tmp_path / "view2.out"
    view.main([str(pz), "--genome", w["refdir"], "-o", str(o2), "-s", f"{s}-{e}", "--strict"])
    assert o2.read_bytes() == run("-s", f"{s}-{e}", "--strict")
    # a region at the very start of chr2 must not pull chr1 reads (tabix is per chromosome)
    s0 = g2.first_idx; m = (chrom == b"chr2") & (idx <= s0 + 49)
    sub = b"".join(l for l, k in zip(lines, m) if k)
    assert run("-s", f"{s0}-{s0 + 50}") == H.port_collapse_pat(H.sort_pat(H.ref_cview(sub, sites=(s0, s0 + 50))))
    # -L: blocks on both chromosomes; tabix -R over the blocks extended by 100 sites
    rng = np.random.default_rng(8)
    cuts = np.sort(rng.choice(np.arange(2, N), size=400, replace=False)); bl = list(zip(cuts[0::2].tolist(), cuts[1::2].tolist()))
    loci = np.concatenate([g1.loci, g2.loci])
    bed = tmp_path / "view_blocks.bed"
    bed.write_text("#chr\tstart\tend\tstartCpG\tendCpG\n" + "".join(
        f"{'chr1' if a <= g1.n_cpg else 'chr2'}\t{loci[a - 1]}\t{loci[b - 2] + 1}\t{a}\t{b}\n" for a, b in bl) + "chr2\t5\t6\tNA\tNA\n")
    keep = np.zeros(len(lines), bool)
    for a, b in bl:
        keep |= (idx >= max(1, a - 100)) & (idx <= b)
    sub = b"".join(l for l, k in zip(lines, keep) if k)
    for argv, kw in flagsets:
        exp = H.port_collapse_pat(H.sort_pat(H.ref_cview(sub, blocks_path=str(bed), **kw)))
        assert len(exp) > 1000 and run("-L", str(bed), *argv) == exp, argv


def test_view_cli_beta_and_beta_to_blocks_cli(world, tmp_path):
    from wgbs_tools_b200 import beta_to_blocks, view
    w = world; H = w["H"]; N = w["N"]
    beta = synth.make_betas(5, 1, N)[0]
    bp = tmp_path / "t.beta"; beta.tofile(bp)
    o = tmp_path / "b.txt"
    view.main([str(bp), "--genome", w["refdir"], "-s", "100-104", "-o", str(o)])
    loci = np.concatenate([w["g1"].loci, w["g2"].loci])
    assert o.read_bytes() == b"".join(b"chr1\t%d\t%d\t%d\t%d\n" % (loci[i] - 1, loci[i] + 1, beta[i, 0], beta[i, 1]) for i in range(99, 103))
    blocks = synth.make_blocks(3, 1, N, mean_len=12.0)
    bed = tmp_path / "bl.bed"
    bed.write_bytes(b"chr\tstart\tend\tstartCpG\tendCpG\n" + synth.blocks_text("chr1", blocks, loci) + b"chr2\t1\t2\tNA\tNA\n")
    out = tmp_path / "o"; out.mkdir()
    beta_to_blocks.main([str(bp), "-b", str(bed), "-o", str(out), "--bedGraph"])
    beta_to_blocks.main([str(bp), "-b", str(bed), "-o", str(out), "-l"])
    sums = np.array([beta[s - 1:e - 1].sum(axis=0, dtype=np.int64) for s, e in blocks.tolist()] + [[0, 0]])
    assert (out / "t.bin").read_bytes() == H.ref_trim(sums).tobytes()
    assert (out / "t.lbeta").read_bytes() == H.ref_trim(sums, lbeta=True).tobytes()
    bg = (out / "t.bedGraph").read_text().splitlines()
    assert len(bg) == sums.shape[0] and bg[-1].endswith("\t-1\t0")
    t = bg[0].split("\t"); assert t[3] == "%.2f" % (sums[0, 0] / sums[0, 1]) and int(t[4]) == sums[0, 1]


def test_beta_to_table_cli_matches_reference_golden(ctx, tmp_path):
    """`wgbstools beta_to_table` with the per-block sums from wgbs_beta_to_blocks: the text the reference's own beta_to_table.py printed"""
    from test_host_logic import check_beta_table
    check_beta_table(ctx, tmp_path)
