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
        homog.main([str(pg), "-b", str(bp), "-p", str(w["dir"] / f"hom_{tag}"), "-f", "-l", "3"])
        got = np.array([l.split(b"\t")[5:8] for l in gzip.open(str(w["dir"] / f"hom_{tag}.uxm.bed.gz")).read().splitlines()], dtype=np.int64)
        bl = np.array([[int(l.split(b"\t")[3]), int(l.split(b"\t")[4])] for l in ls])
        order = np.lexsort((bl[:, 1], bl[:, 0]))
        ref_sorted = H.port_homog(w["pat"], bl[order], np.array([0, 0.334, 0.667, 1], np.float32), 3)
        exp = np.empty_like(ref_sorted); exp[order] = ref_sorted
        np.testing.assert_array_equal(got, exp)
        if H.have_ref():
            refc = H.ref_homog(w["pat"], str(bp), "0,0.334,0.667,1", 3, sort_blocks=(tag == "shuffled"))
            np.testing.assert_array_equal(got[order] if tag == "shuffled" else got, refc)
    homog.main([str(pg), "-b", str(w["dir"] / "blocks_sorted.bed"), "-p", str(w["dir"] / "hom_bin"), "-f", "--binary"])
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
