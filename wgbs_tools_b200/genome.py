"""Genome reference files, as laid out by the reference's init_genome (src/python/init_genome.py:151-187,246-281):
    <dir>/CpG.bed.gz        chr \\t locus \\t idx   (BGZF text; any gzip reader can stream it)
    <dir>/CpG.chrome.size   chr \\t nCpG          (chromosome order of the index)
No tabix here: the dictionary is read once and kept as numpy arrays (4 B per CpG)."""
from __future__ import annotations

import gzip
import os
import threading
import re

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
        self._lengths = None
        self._lock = threading.Lock()      # the lazy caches below are filled once, whichever thread asks first

    def all_loci(self) -> np.ndarray:
        """uint32[nr_sites] locus of CpG i+1 (all chromosomes, index order)"""
        if self._loci is not None:
            return self._loci
        with self._lock:                                       # one thread parses the ~28M-line dictionary, the others wait for it
            if self._loci is not None:
                return self._loci
            cache = self.dict_path + ".loci.npy"
            if os.path.isfile(cache) and os.path.getmtime(cache) >= os.path.getmtime(self.dict_path):
                loci = np.load(cache)
            else:
                op = gzip.open if open(self.dict_path, "rb").read(2) == b"\x1f\x8b" else open
                with op(self.dict_path, "rb") as f:
                    loci = np.array([int(l.split(b"\t", 2)[1]) for l in f], np.uint32)
                try:
                    np.save(cache, loci)
                except OSError:
                    pass
            if loci.size != self.nr_sites:
                raise IllegalArgumentError("CpG.bed.gz does not match CpG.chrome.size")
            self._loci = loci                                  # published only when complete
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


    def chrom_length(self, chrom: str) -> int:
        """bp length from chrome.size (falls back to the last CpG locus + 1 when the file is absent)"""
        if self._lengths is None:
            with self._lock:                                   # bam2pat --gpu_streams calls this from several worker threads
                if self._lengths is None:
                    lengths = {}
                    path = os.path.join(self.dir, "chrome.size")
                    if os.path.isfile(path):
                        for l in open(path):
                            c, n = l.split()
                            lengths[c] = int(n)
                    self._lengths = lengths                    # published only when complete
        if chrom in self._lengths:
            return self._lengths[chrom]
        loci, _ = self.chrom_loci(chrom)
        return int(loci[-1]) + 1 if loci.size else 0


_CHR = r"(chr)?([\d]+|[XYM]|(MT))"


class GenomicRegion:
    """`-s/--sites START-END` (CpG indices, end exclusive) or `-r/--region chr[:from[-to]]`, resolved against the CpG
    dictionary the way the reference does (genomic_region.py:76-176): `sites` = (s1, s2), `region_str`, `bp_tuple`.
    A site whose locus equals the region's end is NOT part of the region (genomic_region.py:139-144)."""

    def __init__(self, ref: GenomeRef, region: str | None = None, sites: str | None = None):
        self.ref = ref
        self.chrom = None; self.sites = None; self.region_str = None; self.bp_tuple = None
        if sites:
            self._parse_sites(sites)
        elif region:
            self._parse_region(region)
        self.nr_sites = None if self.sites is None else self.sites[1] - self.sites[0]

    def is_whole(self) -> bool:
        return self.sites is None

    def _sites_tuple(self, sites_str: str) -> tuple[int, int]:
        if not sites_str:
            raise IllegalArgumentError(f"Empty sites string: {sites_str}")
        sites_str = sites_str.replace(",", "")
        m = re.match(r"([\d]+)-([\d]+)", sites_str)
        if m:
            s1, s2 = int(m.group(1)), int(m.group(2))
        elif "-" not in sites_str and sites_str.isdigit():
            s1 = int(sites_str); s2 = s1 + 1
        else:
            raise IllegalArgumentError(f'sites must be of format: "start-end" or "site" .\nGot: {sites_str}')
        n = self.ref.nr_sites
        if not n + 1 >= s2 >= s1 >= 1:
            raise IllegalArgumentError(f"sites violate the constraints: {n + 1} >= {s2} > {s1} >= 1")
        if s1 == s2:
            s2 += 1
        return s1, s2

    def _parse_sites(self, sites_str: str):
        s1, s2 = self._sites_tuple(sites_str)
        self.chrom = self.ref.chrom_of_site(s1)
        if self.chrom != self.ref.chrom_of_site(s2 - 1):
            raise IllegalArgumentError("Invalid sites input")               # range crosses chromosomes
        lo, hi = self.ref.locus_of_site(s1), self.ref.locus_of_site(s2 - 1) + 1   # include the whole last site (C and G)
        self.sites = (s1, s2); self.region_str = f"{self.chrom}:{lo}-{hi}"; self.bp_tuple = (lo, hi)

    def _parse_region(self, region: str):
        region = region.replace(",", "")
        if re.match(rf"^{_CHR}$", region):
            if region not in self.ref.chroms:
                raise IllegalArgumentError(f"Unknown chromosome: {region}")
            self.chrom = region
            lo, hi = 1, self.ref.chrom_length(region)
        else:
            m = re.match(rf"^{_CHR}:([\d]+)$", region)
            if m:
                region += f"-{int(m.group(4)) + 1}"
            m = re.match(rf"^({_CHR}):([\d]+)-([\d]+)$", region)
            if not m:
                raise IllegalArgumentError(f"Invalid genomic region: {region}")
            self.chrom = m.group(1)
            if self.chrom not in self.ref.chroms:
                raise IllegalArgumentError(f"Unknown chromosome: {region}")
            lo, hi = int(m.group(5)), int(m.group(6))
        if hi <= lo:
            raise IllegalArgumentError(f"Invalid genomic region: {region}. end before start")
        if hi > self.ref.chrom_length(self.chrom) or lo < 1:
            raise IllegalArgumentError(f"Invalid genomic region: {region}. Out of range")
        loci, first = self.ref.chrom_loci(self.chrom)
        a = int(np.searchsorted(loci, lo, side="left")); b = int(np.searchsorted(loci, hi, side="left"))
        # tabix returns loci in [lo, hi]; the reference then drops a last site sitting exactly on `hi`, and calls the
        # range empty when that leaves a single index (genomic_region.py:139-149)
        b_incl = int(np.searchsorted(loci, hi, side="right"))
        if b_incl <= a or b <= a:
            raise IllegalArgumentError(f"Invalid genomic region: {region}. No CpGs in range")
        self.region_str = region; self.bp_tuple = (lo, hi); self.sites = (first + a, first + b)


def extend_region(region: str, by: int = 1000) -> str:
    """bam2pat.py:31-38: the dictionary patter loads covers the region +- MAX_READ_SIZE"""
    if ":" not in region:
        return region
    chrom, r = region.split(":")
    start, end = map(int, r.split("-"))
    return f"{chrom}:{max(1, start - by)}-{end + by}"


def parse_region(ref: GenomeRef, args):
    """-s START-END (CpG indices), -r chr:start-end, -L bed with startCpG/endCpG columns, or the whole genome: list of
    (startCpG, endCpG) (reference genomic_region.py / segment.py:94-122)."""
    if getattr(args, "sites", None) or getattr(args, "region", None):
        return [GenomicRegion(ref, region=getattr(args, "region", None), sites=getattr(args, "sites", None)).sites]
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
