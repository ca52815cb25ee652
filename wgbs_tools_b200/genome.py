"""Genome reference files, as laid out by the reference's init_genome (src/python/init_genome.py:151-187,246-281):
    <dir>/CpG.bed.gz        chr \\t locus \\t idx   (BGZF text; any gzip reader can stream it)
    <dir>/CpG.chrome.size   chr \\t nCpG          (chromosome order of the index)
No tabix here: the dictionary is read once and kept as numpy arrays (4 B per CpG)."""
from __future__ import annotations

import gzip
import os

import numpy as np


class IllegalArgumentError(ValueError):
    pass


def ref_root() -> str:
    return os.environ.get("WGBS_REF_DIR", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "references"))


class GenomeRef:
    def __init__(self, genome: str | None = None):
        d = genome if genome and os.path.isdir(genome) else os.path.join(ref_root(), genome or "default")
        if not os.path.isdir(d):
            raise IllegalArgumentError(f"Invalid reference name: {genome}")
        self.dir = os.path.realpath(d)
        self.name = os.path.basename(self.dir)
        self.dict_path = os.path.join(self.dir, "CpG.bed.gz")
        self.chroms, sizes = [], []
        with open(os.path.join(self.dir, "CpG.chrome.size")) as f:
            for l in f:
                c, n = l.split()
                self.chroms.append(c); sizes.append(int(n))
        self.sizes = np.array(sizes, np.int64)
        self.first = np.concatenate([[1], 1 + np.cumsum(self.sizes)])      # first CpG index per chromosome (+ sentinel)
        self.nr_sites = int(self.sizes.sum())
        self._loci = None

    def all_loci(self) -> np.ndarray:
        """uint32[nr_sites] locus of CpG i+1 (all chromosomes, index order)"""
        if self._loci is None:
            cache = self.dict_path + ".loci.npy"
            if os.path.isfile(cache) and os.path.getmtime(cache) >= os.path.getmtime(self.dict_path):
                self._loci = np.load(cache)
            else:
                op = gzip.open if open(self.dict_path, "rb").read(2) == b"\x1f\x8b" else open
                with op(self.dict_path, "rb") as f:
                    self._loci = np.array([int(l.split(b"\t", 2)[1]) for l in f], np.uint32)
                try:
                    np.save(cache, self._loci)
                except OSError:
                    pass
            if self._loci.size != self.nr_sites:
                raise IllegalArgumentError("CpG.bed.gz does not match CpG.chrome.size")
        return self._loci

    def chrom_range(self, chrom: str) -> tuple[int, int]:
        """[startCpG, endCpG) of a chromosome"""
        i = self.chroms.index(chrom)
        return int(self.first[i]), int(self.first[i + 1])

    def chrom_loci(self, chrom: str) -> tuple[np.ndarray, int]:
        s, e = self.chrom_range(chrom)
        return self.all_loci()[s - 1:e - 1], s

    def chrom_of_site(self, site: int) -> str:
        return self.chroms[int(np.searchsorted(self.first, site, side="right")) - 1]

    def locus_of_site(self, site: int) -> int:
        return int(self.all_loci()[site - 1])


def parse_region(ref: GenomeRef, args):
    """-s START-END (CpG indices), -r chr:start-end, -L bed with startCpG/endCpG columns, or the whole genome: list of
    (startCpG, endCpG) (reference genomic_region.py / segment.py:94-122)."""
    if getattr(args, "sites", None):
        a, b = args.sites.split("-")
        return [(int(a), int(b))]
    if getattr(args, "region", None):
        r = args.region
        if ":" not in r:
            return [ref.chrom_range(r)]
        c, se = r.split(":"); s, e = (int(x.replace(",", "")) for x in se.split("-"))
        loci, first = ref.chrom_loci(c)
        a = int(np.searchsorted(loci, s, side="left")); b = int(np.searchsorted(loci, e, side="right"))
        if b <= a:
            raise IllegalArgumentError(f"Invalid genomic region: {r}. No CpGs in range")
        return [(first + a, first + b)]
    if getattr(args, "bed_file", None):
        out = []
        for l in open(args.bed_file):
            if l.startswith("#") or not l.strip():
                continue
            t = l.split("\t")
            if t[3].strip().isdigit():
                out.append((int(t[3]), int(t[4])))
        return out
    return [ref.chrom_range(c) for c in ref.chroms]
