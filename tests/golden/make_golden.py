#!/usr/bin/env python3
"""Generate tests/golden/*.npz / *.json by running the REFERENCE's own Python host logic (imported from
/root/reference/src/python) -- only possible in the build container; the fixtures travel, this script documents them.

  segment_stitch.npz : reference segment.py chunk solving + merge_df_list/stitch_2_dfs (segment.py:157-165,199-252) with
                       `segment_process` redirected to the oracle's port of `segmentor` (no tabix here), on seeded betas.
  beta_to_table.json : reference beta_to_table.py (beta2table_generator + dump: get_table :72-107, np.add.reduceat sums, beta2vec, nanmean over
                       groups, %.{digits}f / NA text) on seeded betas, blocks with NA rows, with and without a groups file.
  homog_host.json    : reference homog.py thresholds strings (homog.py:96-104) and trim_uxm_to_uint8 (homog.py:48-58);
                       utils_wgbs.trim_to_uint8 (utils_wgbs.py:277-290) on edge rows.
"""
import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src/python")
from oracle import harness as H          # noqa: E402
from wgbs_tools_b200 import synth        # noqa: E402

import segment as ref_segment            # noqa: E402  (the reference module)
import homog as ref_homog                # noqa: E402
import utils_wgbs as ref_utils           # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def make_segment():
    cases = []
    for seed, N, K, chunk, max_cpg, max_bp, ps in [(1, 3000, 4, 500, 100, 1500, 15), (2, 5000, 3, 700, 300, 3000, 15),
                                                    (3, 1200, 6, 100, 60, 800, 1), (4, 2500, 2, 2500, 1000, 2000, 15),
                                                    (5, 2049, 5, 256, 128, 4000, 15)]:
        betas = synth.make_betas(100 + seed, K, N)
        g = synth.make_genome(200 + seed, "chr1", 1_000_000, with_bases=False)
        assert g.n_cpg >= N
        d = g.loci[:N]
        first = 1                                           # CpG index of array row 0

        def segment_process(params, betas=betas, d=d, max_cpg=max_cpg, max_bp=max_bp, ps=ps):
            start, end = params["sites"]
            assert end - start > 0
            if end - start == 1:
                return np.array([start, end])
            a, b = start - first, end - first
            return H.port_segment([x[a:b] for x in betas], d[a:b], max_cpg, max_bp, ps) + start

        ref_segment.segment_process = segment_process
        starts = list(range(first, first + N, chunk)); ends = starts[1:] + [first + N]
        arr = [segment_process({"sites": (s, e)}) for s, e in zip(starts, ends)]
        pool = types.SimpleNamespace(starmap=lambda f, a: [f(*x) for x in a])
        self = types.SimpleNamespace(param_dict={})
        merged = ref_segment.SegmentByChunks.merge_df_list(self, list(arr), pool)
        cases.append(dict(seed=seed, N=N, K=K, chunk=chunk, max_cpg=max_cpg, max_bp=max_bp, ps=ps, merged=np.asarray(merged)))
    np.savez(os.path.join(OUT, "segment_stitch.npz"), **{f"c{i}_{k}": np.asarray(v) for i, c in enumerate(cases) for k, v in c.items()})
    print("segment_stitch.npz:", [(c["N"], len(c["merged"])) for c in cases])



def beta_table_inputs(tmp):
    """the seeded inputs of the beta_to_table cases, written into directory tmp (the test re-creates them the same way)"""
    N = 20_000
    betas = synth.make_betas(31, 5, N)
    betas[3][:, 1] = np.minimum(betas[3][:, 1], 1)              # a low-coverage sample: many blocks below --min_cov
    betas[3][:, 0] = np.minimum(betas[3][:, 0], betas[3][:, 1])
    paths = []
    for i, b in enumerate(betas):
        p = os.path.join(tmp, f"s{i}.beta"); b.tofile(p); paths.append(p)
    blocks = synth.make_blocks(7, 1, N, mean_len=6)[:900]
    rows = []
    for k, (a, b) in enumerate(blocks.tolist()):
        if k % 53 == 5:
            rows.append(f"chr1\t{1000 + k}\t{1000 + k + 1}\tNA\tNA\n")
        rows.append(f"chr1\t{10 * a}\t{10 * b}\t{a}\t{b}\n")
    bp = os.path.join(tmp, "blocks.bed"); open(bp, "w").write("".join(rows))
    gp = os.path.join(tmp, "groups.csv")
    open(gp, "w").write("name,group,include\ns0,A,True\ns1,B,True\ns2,A,True\ns3,B,True\ns4,A,False\n")
    return paths, bp, gp


def make_beta_table():
    import io
    import tempfile
    import contextlib
    import beta_to_table as ref_b2t
    cases = []
    with tempfile.TemporaryDirectory() as tmp:
        paths, bp, gp = beta_table_inputs(tmp)
        for groups, min_cov, digits, chunk in ((None, 4, 2, 200000), (gp, 4, 2, 300), (gp, 1, 3, 200000), (None, 30, 1, 450)):
            outp = os.path.join(tmp, "t.tsv")
            first = True
            for ch in ref_b2t.beta2table_generator(paths, bp, groups, min_cov, 2, chunk, False):
                ref_b2t.dump(outp, ch, first, digits); first = False
            cases.append(dict(groups=bool(groups), min_cov=min_cov, digits=digits, chunk=chunk, text=open(outp).read()))
    json.dump(cases, open(os.path.join(OUT, "beta_to_table.json"), "w"))
    print("beta_to_table.json:", [len(c["text"]) for c in cases])


def make_homog():
    out = {"rates": {}, "trim_uxm": [], "trim_beta": []}
    for l in range(3, 12):
        th1 = round(1 - (l - 1) / l, 3) + 0.001
        th2 = round((l - 1) / l, 3)
        out["rates"][str(l)] = f"0,{th1},{th2},1"            # homog.py:101-104 verbatim
    rng = np.random.default_rng(0)
    data = rng.integers(0, 3000, size=(200, 3))
    data[:50] = rng.integers(0, 200, size=(50, 3))
    out["trim_uxm"] = {"in": data.tolist(), "u8": ref_homog.trim_uxm_to_uint8(data, 8).tolist(), "u16": ref_homog.trim_uxm_to_uint8(data * 40, 16).tolist()}
    mc = np.array([[100, 510], [255, 256], [7, 1000], [0, 0], [255, 255], [256, 256], [3, 70000], [65535, 65536]], dtype=np.int64)
    out["trim_beta"] = {"in": mc.tolist(), "u8": ref_utils.trim_to_uint8(mc.copy()).tolist(), "u16": ref_utils.trim_to_uint8(mc.copy(), True).tolist()}
    json.dump(out, open(os.path.join(OUT, "homog_host.json"), "w"))
    print("homog_host.json ok")




def make_tutorial_reads():
    """tests/golden/tutorial_reads.sam.gz: ~1000 REAL bisulfite reads (hg19 chr3, SE, CIGAR ops M/I/D/S/H) decoded from the
    reference's tutorial BAMs (tutorial/bams/{Pancreas_STL002,Lung_STL002}.small.bam) with wgbs_tools_b200.bamio -- the
    only real reads available offline; used as a CIGAR-diversity fixture against a synthetic CpG dictionary."""
    import gzip
    from wgbs_tools_b200 import bamio
    out = []
    for name in ("Pancreas_STL002.small", "Lung_STL002.small"):
        lines = bamio.BamFile(f"/root/reference/tutorial/bams/{name}.bam").view().splitlines(keepends=True)
        odd = [l for l in lines if any(c in l.split(b"\t")[5] for c in b"IDSHN")]
        plain = [l for l in lines if l not in odd][:300]
        out += odd[:300] + plain
    out.sort(key=lambda l: (l.split(b"\t")[2], int(l.split(b"\t")[3])))
    gzip.open(os.path.join(OUT, "tutorial_reads.sam.gz"), "wb").write(b"".join(out))


def copy_tutorial_bams():
    """two of the reference's tutorial BAMs (written by htslib): known line counts in tutorial/README.md:75,80"""
    import shutil
    for name in ("Lung_STL002.small.bam", "Pancreas_STL002.small.bam"):
        shutil.copyfile(f"/root/reference/tutorial/bams/{name}", os.path.join(OUT, name))


if __name__ == "__main__":
    make_segment()
    make_homog()
    make_tutorial_reads()
    copy_tutorial_bams()
    make_beta_table()
