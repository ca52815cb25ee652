"""init_genome -- build the CpG dictionary of a FASTA (reference src/python/init_genome.py:151-187, 246-281), without
samtools / tabix / bgzip: SURVEY.md 8f-4.

Outputs in <out_dir> (same names, same text content as the reference):
    CpG.bed.gz       `chr \\t locus \\t idx`: locus = 1-based position of the C of every `CG` in the upper-cased sequence
                     (re.finditer('CG'), init_genome.py:257), idx = 1-based running index over chromosomes in chromosome_order
    CpG.chrome.size  `chr \\t nCpG`
    chrome.size      `chr \\t length`
Only chromosomes matching ^(chr)?(\\d+|[XYM]|MT)$ are kept (is_valid_chrome, :278-281) and they are sorted by
chromosome_order (:263-275) unless --no_sort.  The CG scan is one vectorised numpy comparison per chromosome."""
from __future__ import annotations

import argparse
import gzip
import os
import re

import numpy as np

from .patio import bgzf_compress


def chromosome_order(c: str) -> int:
    if c.startswith("chr"):
        c = c[3:]
    if c.isdigit():
        return int(c)
    return {"X": 10000, "Y": 10001, "M": 10002, "MT": 10002}.get(c, 10003)


def is_valid_chrome(c: str) -> bool:
    return bool(re.match(r"^(chr)?([\d]+|[XYM]|(MT))$", c))


def read_fasta(path: str):
    """yield (name, uint8 upper-cased sequence)"""
    op = gzip.open if path.endswith(".gz") else open
    name, parts = None, []
    with op(path, "rb") as f:
        for line in f:
            if line.startswith(b">"):
                if name is not None:
                    yield name, np.frombuffer(b"".join(parts).upper(), np.uint8)
                name, parts = line[1:].split()[0].decode(), []
            else:
                parts.append(line.strip())
    if name is not None:
        yield name, np.frombuffer(b"".join(parts).upper(), np.uint8)


def cpg_loci(seq: np.ndarray) -> np.ndarray:
    """1-based positions of the C of every non-overlapping... `CG` (CG cannot overlap itself, so finditer == all matches)"""
    return np.flatnonzero((seq[:-1] == ord("C")) & (seq[1:] == ord("G"))).astype(np.int64) + 1 if seq.size > 1 else np.zeros(0, np.int64)


def init_genome(fasta: str, out_dir: str, no_sort: bool = False, threads: int = 8) -> dict:
    chroms = {}
    order = []
    for name, seq in read_fasta(fasta):
        if is_valid_chrome(name):
            chroms[name] = (seq.size, cpg_loci(seq))
            order.append(name)
    if not no_sort:
        order = sorted(order, key=chromosome_order)
    os.makedirs(out_dir, exist_ok=True)
    idx = 1
    parts = []
    with open(os.path.join(out_dir, "CpG.chrome.size"), "w") as fc, open(os.path.join(out_dir, "chrome.size"), "w") as fs:
        for c in order:
            n, loci = chroms[c]
            fs.write(f"{c}\t{n}\n"); fc.write(f"{c}\t{loci.size}\n")
            cb = c.encode()
            parts.append(b"".join(b"%s\t%d\t%d\n" % (cb, l, i) for l, i in zip(loci.tolist(), range(idx, idx + loci.size))))
            idx += loci.size
    with open(os.path.join(out_dir, "CpG.bed.gz"), "wb") as f:
        f.write(bgzf_compress(b"".join(parts), threads))
    return {"chroms": order, "nr_sites": idx - 1}


def main(argv=None):
    p = argparse.ArgumentParser(description="Init genome reference: build the CpG-index dictionary from a FASTA")
    p.add_argument("genome_ref", help="path to a FASTA file (plain or gzip)")
    p.add_argument("-o", "--out_dir", required=True, help="directory to write CpG.bed.gz, CpG.chrome.size, chrome.size into")
    p.add_argument("--no_sort", action="store_true"); p.add_argument("-@", "--threads", type=int, default=8)
    a = p.parse_args(argv)
    r = init_genome(a.genome_ref, a.out_dir, a.no_sort, a.threads)
    print(f"[wt init] {len(r['chroms'])} chromosomes, {r['nr_sites']:,} CpG sites")


if __name__ == "__main__":
    main()
