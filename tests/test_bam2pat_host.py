"""bam2pat's HOST logic on the CPU: the CLI (regions, filters, template windows for chromosomes too large for one call, merge of
the window outputs, BGZF parts, CSI index, beta) runs here with the oracle's C restatement standing in for the device --
a fake Context with the few methods bam2pat uses.  What the kernels compute is the `-m gpu` suite's business; this pins that
piling a chromosome up in template windows (wgbs_view_opts.key_beg / key_end) changes nothing in the outputs, for SAM input,
for the host BAM reader, with region / strand filters, --long and --mbias-free runs."""
import gzip
import os
import threading

import numpy as np
import pytest

from wgbs_tools_b200 import synth


_LOCK = threading.Lock()


class _Buf:
    def __init__(self, n):
        self.a = np.zeros(n // 4, np.int32)

    def free(self):
        pass


class _PortPats:
    def __init__(self, H, lines, counted, long=False):
        self.H, self.lines, self.counted, self.long_names = H, lines, counted, long

    def with_counts(self) -> bytes:
        if self.counted:
            return b"".join(self.lines)
        return b"".join(b"\t".join(l.rstrip(b"\n").split(b"\t")[:3]) + b"\t1\n" for l in self.lines)

    def collapse(self, long=False, mode=None):
        if long:
            self.lines.sort(key=lambda l: (lambda f: (int(f[1]), f[2], f[3]))(l.rstrip(b"\n").split(b"\t")))
            return self
        d = {}
        for l in self.with_counts().splitlines():
            k, c = l.rsplit(b"\t", 1)
            d[k] = d.get(k, 0) + int(c)
        keys = sorted(d, key=lambda k: (lambda f: (int(f[1]), f[2]))(k.split(b"\t")))
        self.lines = [k + b"\t%d\n" % d[k] for k in keys]; self.counted = True
        return self

    def to_text(self, chrom, long=False):
        if long:                                            # chr idx pattern 1 qname
            return b"".join((lambda f: b"\t".join(f[:3] + [b"1", f[3]]) + b"\n")(l.rstrip(b"\n").split(b"\t")) for l in self.lines)
        return b"".join(self.lines)

    def free(self):
        pass


class PortContext:
    """the slice of api.Context that bam2pat uses, computed by the oracle port"""

    def __init__(self, device=0, stream=None):
        from oracle import harness as H
        self.H = H

    def __enter__(self):
        return self

    def __exit__(self, *a):
        pass

    def sync(self):
        pass

    def close(self):
        pass

    def alloc(self, n):
        return _Buf(n)

    def load_index(self, loci, first):
        class Ix:
            def free(self):
                pass
        ix = Ix(); ix.loci = np.asarray(loci); ix.first = first
        return ix

    def pileup_sam(self, ix, sam, min_cpg=1, clip=0, paired=-1, nanopore=False, np_thresh=0.67, cpc_call="C", combine_mods=False,
                   mbias=False, keep_names=False, nbytes=None):
        H = self.H
        sam = bytes(sam)                                     # (the host BAM reader hands over a uint8 array, not bytes)
        first = next((l for l in sam.splitlines() if l), b"")
        pe = bool(int(first.split(b"\t")[1]) & 1) if paired < 0 else bool(paired)
        # the port's patter prints `chr idx pattern`; --long adds the read name (patter --long): the port has no such switch, so
        # the fake keeps names only as a tie-breaker-free placeholder (tests with --long compare windowed vs whole, both through here)
        out, pst = H.port_patter(H.port_match_maker(sam) if pe else sam, ix.loci, ix.first + np.arange(ix.loci.size), min_cpg=min_cpg, clip=clip,
                                 nanopore=nanopore, np_thresh=np_thresh, cpc_call=cpc_call, combine_mods=combine_mods)
        lines = out.splitlines(keepends=True)
        if keep_names:
            lines = [l.rstrip(b"\n") + b"\tq%06d\n" % (hash(l) % 1000000) for l in lines]
        st = dict(zip(("lines", "pairs", "empty", "short", "invalid", "paired"), pst)); st["nanopore"] = int(nanopore); st["templates"] = len(lines)
        return _PortPats(H, lines, False), st

    def pats_from_text(self, text):
        return _PortPats(self.H, bytes(text).splitlines(keepends=True), True)

    def pat2beta_text(self, text, start, end, nbits=8, want_counts=False):
        mc = self.H.port_pat2beta(bytes(text), start, end)
        beta = self.H.port_trim(mc, nbits)
        return (beta, mc) if want_counts else beta

    def pat2beta(self, P, start, end, meth_cov=None, zero_first=True):
        mc = self.H.port_pat2beta(P.with_counts(), start, end) if P.lines else np.zeros((end - start, 2), np.int32)
        if meth_cov is None:
            meth_cov = _Buf((end - start) * 8)
        with _LOCK:                                         # (the device adds atomically; numpy does not)
            if zero_first:
                meth_cov.a[:] = 0
            meth_cov.a += mc.reshape(-1)
        return meth_cov

    def trim(self, mc, n, nbits=8):
        return self.H.port_trim(mc.a.reshape(-1, 2), nbits)


@pytest.fixture()
def world(tmp_path, oracle, monkeypatch, built_lib):
    from wgbs_tools_b200 import api, bamio
    monkeypatch.setattr(api, "Context", PortContext)
    g1 = synth.make_genome(31, "chr1", 500_000, first_idx=1)
    g2 = synth.make_genome(32, "chr2", 300_000, first_idx=1 + g1.n_cpg)
    refdir = tmp_path / "synth"; refdir.mkdir()
    with gzip.open(refdir / "CpG.bed.gz", "wb") as f:
        f.write(g1.dict_text() + g2.dict_text())
    (refdir / "CpG.chrome.size").write_text(f"chr1\t{g1.n_cpg}\nchr2\t{g2.n_cpg}\n")
    (refdir / "chrome.size").write_text(f"chr1\t{g1.length}\nchr2\t{g2.length}\n")
    sam = synth.make_sam(g1, 9_000, 1, paired=True, name_prefix="a", single_frac=0.04) + synth.make_sam(g2, 5_000, 2, paired=True, name_prefix="b")
    (tmp_path / "s.sam").write_bytes(b"@HD\tVN:1.6\tSO:coordinate\n" + sam)
    (tmp_path / "s.bam").write_bytes(bamio.sam_to_bam(sam, [("chr1", g1.length), ("chr2", g2.length)]))
    return tmp_path, str(refdir)


def _run(tmp, refdir, inp, tag, extra=(), decode="host"):
    from wgbs_tools_b200 import bam2pat
    out = tmp / tag; out.mkdir()
    bam2pat.main([str(tmp / inp), "--genome", refdir, "-o", str(out), "--bam_decode", decode, *extra])
    name = inp.split(".")[0]
    pat = gzip.decompress((out / f"{name}.pat.gz").read_bytes())
    beta = (out / f"{name}.beta").read_bytes() if (out / f"{name}.beta").exists() else None
    return pat, beta


@pytest.mark.parametrize("inp", ["s.sam", "s.bam"])
@pytest.mark.parametrize("extra", [(), ("-r", "chr1:100000-300000"), ("--bottom_strand",), ("--long", "--no_beta"), ("--clip", "5", "--min_cpg", "2")])
def test_windowed_chromosomes_give_the_same_outputs(world, monkeypatch, inp, extra):
    tmp, refdir = world
    monkeypatch.setenv("WGBS_CHUNK_RECORDS", "0"); monkeypatch.setenv("WGBS_CHUNK_BYTES", "0")       # 0: never split
    whole = _run(tmp, refdir, inp, "whole", extra)
    assert len(whole[0]) > 2000
    monkeypatch.setenv("WGBS_CHUNK_RECORDS", "1500"); monkeypatch.setenv("WGBS_CHUNK_BYTES", "500000")   # 4-7 windows per chromosome
    assert _run(tmp, refdir, inp, "windows", extra) == whole


def test_template_windows_cover_everything():
    from wgbs_tools_b200.bam2pat import template_windows
    assert template_windows(100, 0, 0, 1000) is None and template_windows(100, 100, 0, 1000) is None
    w = template_windows(1000, 300, 0, 1000)
    assert w[0][0] == 0 and w[-1][1] == 1 << 40 and all(a[1] == b[0] for a, b in zip(w, w[1:])) and len(w) == 4
    w = template_windows(10**9, 6_000_000, 99_000, 301_000)
    assert w[0][0] == 0 and w[-1][1] == 1 << 40 and all(a[1] == b[0] and a[0] < a[1] for a, b in zip(w, w[1:]))
    assert template_windows(10, 3, 5, 6) in (None, [(0, 1 << 40)])                                       # nothing to cut


@pytest.mark.parametrize("extra", [(), ("-r", "chr2:50000-200000"), ("--top_strand",), ("--long", "--no_beta")])
@pytest.mark.parametrize("budget", ["150000", "700000"])
def test_streamed_bam_gives_the_same_outputs(world, monkeypatch, extra, budget):
    """--bam_decode stream: the file read as parts of WGBS_STREAM_BYTES inflated bytes (records cut off at part ends, templates
    deferred until both mates are in) == the file decoded as a whole"""
    tmp, refdir = world
    monkeypatch.setenv("WGBS_CHUNK_RECORDS", "0")
    whole = _run(tmp, refdir, "s.bam", "whole", extra)
    monkeypatch.setenv("WGBS_STREAM_BYTES", budget)
    assert _run(tmp, refdir, "s.bam", "stream", extra, decode="stream") == whole


def test_pat2beta_cli_on_a_pat_larger_than_one_call(tmp_path, oracle, monkeypatch, built_lib):
    """pat2beta with the text cut into pieces (WGBS_PAT_CHUNK_BYTES) == in one call; gzip input (the host decode path)"""
    from wgbs_tools_b200 import api
    from wgbs_tools_b200 import pat2beta as p2b
    monkeypatch.setattr(api, "Context", PortContext)
    N = 40_000
    txt = synth.make_pat_text_fast(5, 30_000, N, chrom="chr1")
    pg = tmp_path / "x.pat.gz"
    with gzip.open(pg, "wb") as f:
        f.write(txt)
    outs = []
    for limit in (None, "50000", "7000"):
        if limit:
            monkeypatch.setenv("WGBS_PAT_CHUNK_BYTES", limit)
        d = tmp_path / f"o{limit}"; d.mkdir()
        p2b.pat2beta(PortContext(), str(pg), str(d), N, decode="host")
        outs.append((d / "x.beta").read_bytes())
    assert outs[0] == outs[1] == outs[2] == oracle.port_trim(oracle.port_pat2beta(txt, 1, N + 1)).tobytes()


@pytest.mark.parametrize("inp", ["s.sam", "s.bam"])
def test_chromosomes_in_flight_give_the_same_outputs(world, monkeypatch, inp):
    """--gpu_streams 2: two chromosomes at a time, each on its own Context / host thread, counts added into the one beta array"""
    tmp, refdir = world
    monkeypatch.setenv("WGBS_CHUNK_RECORDS", "0"); monkeypatch.setenv("WGBS_CHUNK_BYTES", "0")
    one = _run(tmp, refdir, inp, "one")
    assert _run(tmp, refdir, inp, "two", ("--gpu_streams", "2")) == one
    monkeypatch.setenv("WGBS_CHUNK_RECORDS", "1500"); monkeypatch.setenv("WGBS_CHUNK_BYTES", "500000")
    assert _run(tmp, refdir, inp, "two_windows", ("--gpu_streams", "2", "-v")) == one
