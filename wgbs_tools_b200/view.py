"""view -- `wgbstools view X.pat.gz` / `wgbstools cview` (reference src/python/view.py, cview.py): the records of a pat file
that overlap a region / a sites range / the blocks of a bed file, optionally clipped to them (--strict), stripped of
edge dots (--strip), without gaps (--no_gaps), at least --min_len sites long; then `sort -k2,2n -k3,3 | collapse_pat.pl`.

    [tabix pat chr:s-e | gunzip -c pat] | cview --sites "s\\te" | --blocks_path BED ... | sort -k2,2n -k3,3 | collapse_pat.pl -
becomes wgbs_pats_from_text + wgbs_cview + wgbs_collapse_ex + wgbs_pats_format per chromosome.
Beta / lbeta / bin inputs are printed like the reference's view_beta.sh / np.savetxt (host formatting of a slice)."""
from __future__ import annotations

import argparse
import sys

import numpy as np

from .genome import GenomeRef, GenomicRegion, IllegalArgumentError

MAX_PAT_LEN = 150                  # utils_wgbs.py:38: how far before a region tabix looks for reads reaching into it
COLLAPSE_DOTTED, COLLAPSE_ADJACENT = 2, 3


def load_cview_blocks(path: str) -> list[tuple[int, int]]:
    """columns 4-5 of a blocks file the way cview reads them (cview.cpp:19-85): `cut -f4-5 | sort -k1,1n`, `#` lines,
    a header row and NA rows skipped; invalid blocks raise like the reference's exceptions"""
    import gzip
    op = gzip.open if path.endswith(".gz") else open
    rows = []
    with op(path, "rb") as f:
        for l in f:
            t = l.rstrip(b"\r\n").split(b"\t")
            rows.append(b"\t".join(t[3:5]))

    def num(b: bytes) -> float:                                     # `sort -n`: leading numeric prefix, else 0
        s = b.split(b"\t")[0].strip(); k = 1 if s[:1] == b"-" else 0
        while k < len(s) and s[k:k + 1].isdigit():
            k += 1
        try:
            return float(s[:k])
        except ValueError:
            return 0.0
    rows.sort(key=lambda r: (num(r), r))
    out = []
    for r in rows:
        if not r or r.startswith(b"#"):
            continue
        t = r.split(b"\t")
        if len(t) != 2:
            raise IllegalArgumentError("Invalid block file")
        if not out and not t[0].isdigit():
            continue                                                # header
        if t[0] in (b"NA", b"NaN"):
            continue
        s, e = int(t[0]), int(t[1])
        if e <= s:
            raise IllegalArgumentError("Invalid block: endCpG <= startCpG")
        if s < 1:
            raise IllegalArgumentError("Invalid block: startCpG < 1")
        out.append((s, e))
    if not out:
        raise IllegalArgumentError("Error while loading blocks. 0 blocks found.")
    return out


def extended_ranges(blocks: list[tuple[int, int]], by: int = 100) -> tuple[np.ndarray, np.ndarray]:
    """cview/extend_blocks.sh + `tabix -R`: reads whose start lies in [max(1, startCpG - 100), endCpG] of some block; merged"""
    iv = sorted((max(1, s - by), e) for s, e in blocks)
    lo, hi = [], []
    for a, b in iv:
        if lo and a <= hi[-1]:                                      # bedtools merge: overlapping or book-ended
            hi[-1] = max(hi[-1], b)
        else:
            lo.append(a); hi.append(b)
    return np.array(lo, np.int32), np.array(hi, np.int32)


def split_by_chrom(text: bytes, chroms: list[str]) -> list[tuple[str, bytes]]:
    """a pat file is the `cat` of per-chromosome parts in chromosome order: cut it at the chromosome changes"""
    at = []
    for c in chroms:
        key = c.encode() + b"\t"
        p = 0 if text.startswith(key) else text.find(b"\n" + key)
        if p >= 0:
            at.append((p if p == 0 else p + 1, c))
    at.sort()
    return [(c, text[p:(at[i + 1][0] if i + 1 < len(at) else len(text))]) for i, (p, c) in enumerate(at)]


def view_pat(ctx, ref: GenomeRef, text: bytes, *, gr: GenomicRegion | None = None, blocks=None, prefilter: bool = True, strict=False,
             strip=False, no_gaps=False, min_len: int = 1, no_sort: bool = False, nanopore: bool = False) -> bytes:
    """the whole `wgbstools view` of a pat file on the device; returns the text the reference prints"""
    pre = None
    if blocks is not None:                                          # view_bed (cview.py:79-100): always sorted
        bl = blocks; mode = COLLAPSE_DOTTED
        if prefilter:
            pre = extended_ranges(bl)
    elif gr is None or gr.is_whole():                               # view_gr, whole file (cview.py:29-33): no sort
        bl = [(1, ref.nr_sites + 1)]; mode = COLLAPSE_ADJACENT
    else:
        s, e = gr.sites
        mpl = 100000 if nanopore else MAX_PAT_LEN
        first, _ = ref.chrom_range(gr.chrom)
        pre = (np.array([max(1, s - mpl, first)], np.int32), np.array([e - 1], np.int32))      # tabix pat chrom:ms-(e-1)
        bl = [(s, e)]; mode = COLLAPSE_ADJACENT if no_sort else COLLAPSE_DOTTED
    bs = [b[0] for b in bl]; be = [b[1] for b in bl]
    out = []
    for chrom, part in split_by_chrom(text, ref.chroms):
        if gr is not None and not gr.is_whole() and chrom != gr.chrom:
            continue
        P = ctx.pats_from_text(part)
        V = P.cview(bs, be, strict=strict, strip=strip, no_gaps=no_gaps, min_cpgs=min_len, pre=pre)
        V.collapse(mode=mode)
        out.append(V.to_text(chrom))
        P.free(); V.free()
    return b"".join(out)


def view_beta_text(ref: GenomeRef, path: str, gr: GenomicRegion) -> bytes:
    """view_beta.sh / view_lbeta.sh: `chr  locus-1  locus+1  meth  cov` per site"""
    dt = np.uint16 if path.endswith(".lbeta") else np.uint8
    data = np.fromfile(path, dt).reshape(-1, 2)
    if data.shape[0] != ref.nr_sites:
        raise IllegalArgumentError(f"beta file {path} does not match the genome reference ({ref.name})")
    s, e = (1, ref.nr_sites + 1) if gr.is_whole() else gr.sites
    loci = ref.all_loci()
    out = []
    for c in ref.chroms:
        a, b = ref.chrom_range(c)
        a, b = max(a, s), min(b, e)
        cb = c.encode()
        out += [b"%s\t%d\t%d\t%d\t%d\n" % (cb, l - 1, l + 1, m, v) for l, (m, v) in zip(loci[a - 1:b - 1].tolist(), data[a - 1:b - 1].tolist())]
    return b"".join(out)


def add_view_flags(p):
    p.add_argument("-s", "--sites"); p.add_argument("-r", "--region"); p.add_argument("--genome")
    p.add_argument("-L", "--bed_file", help="bed file with startCpG / endCpG in columns 4-5")
    p.add_argument("--strict", action="store_true", help="Truncate reads that start/end outside the given region")
    p.add_argument("--strip", action="store_true", help="Remove trailing dots (from beginning/end of reads)")
    p.add_argument("--min_len", type=int, default=1, help="Display only reads covering at least MIN_LEN CpG sites [1]")
    p.add_argument("--no_gaps", action="store_true", help="Remove reads with gaps (dots) in them")
    p.add_argument("--no_sort", action="store_true", help="Keep read order, as in the original pat file")
    p.add_argument("--shuffle", action="store_true", help="not supported (random order)")
    p.add_argument("--sub_sample", type=float, help="not supported (random sub-sampling)")
    p.add_argument("-o", "--out_path", help="Output path. [stdout]")
    p.add_argument("-np", "--nanopore", action="store_true", help="pull very long reads starting before the requested region")
    return p


def main(argv=None):
    from .api import Context
    from .patio import read_pat_text
    p = argparse.ArgumentParser(description="View the content of input file (pat/beta) as plain text")
    p.add_argument("input_file")
    a = add_view_flags(p).parse_args(argv)
    if a.shuffle or a.sub_sample is not None:
        raise IllegalArgumentError("--shuffle / --sub_sample draw random numbers: not supported")
    ref = GenomeRef(a.genome)
    gr = GenomicRegion(ref, region=a.region, sites=a.sites)
    f = a.input_file
    if f.endswith((".beta", ".lbeta")):
        txt = view_beta_text(ref, f, gr)
    elif f.endswith(".bin"):                                        # view_other_bin: np.savetxt of the slice
        data = np.fromfile(f, np.uint8).reshape(-1, 2)
        s, e = (1, data.shape[0] + 1) if gr.is_whole() else gr.sites
        txt = b"".join(b"%d\t%d\n" % (m, v) for m, v in data[s - 1:e - 1].tolist())
    elif f.endswith(".pat.gz") or f.endswith(".pat"):
        import os
        if not gr.is_whole() and not a.bed_file and os.path.isfile(f + ".csi"):
            from .csi import read_region                           # tabix pat chrom:ms-(e-1): only the indexed blocks are inflated
            first, _ = ref.chrom_range(gr.chrom)
            text = read_region(f, gr.chrom, max(1, gr.sites[0] - (100000 if a.nanopore else MAX_PAT_LEN), first), gr.sites[1] - 1)
        else:
            text = read_pat_text(f)
        with Context(0) as ctx:
            if a.bed_file:
                bl = load_cview_blocks(a.bed_file)
                first_chr = next((l.split("\t")[0] for l in (open(a.bed_file) if not a.bed_file.endswith(".gz") else __import__("gzip").open(a.bed_file, "rt"))
                                  if l.strip() and not l.startswith("#")), "")
                txt = view_pat(ctx, ref, text, blocks=bl, prefilter=not (len(bl) >= 1e6 and first_chr in ("1", "chr1")),   # cview.py:84-90
                               strict=a.strict, strip=a.strip, no_gaps=a.no_gaps, min_len=a.min_len)
            else:
                txt = view_pat(ctx, ref, text, gr=gr, strict=a.strict, strip=a.strip, no_gaps=a.no_gaps, min_len=a.min_len,
                               no_sort=a.no_sort, nanopore=a.nanopore)
    else:
        raise IllegalArgumentError(f"Unknown input format: {f}")
    out = sys.stdout.buffer if a.out_path is None else open(a.out_path, "wb")
    out.write(txt)
    if a.out_path is not None:
        out.close()


if __name__ == "__main__":
    main()
