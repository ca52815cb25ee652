"""The `samtools view` stage of reference bam2pat.py:126-159 on SAM TEXT (inputs given as .sam / stdin; .bam inputs are
filtered natively by wgbs_bam_view_ex), plus the BED interval lists both paths share.

    samtools view BAM region -q Q -F X [-f Y] [| awk '($2 == A || $2 == B)'] [-r RG] [-M -L whitelist]
    [... | bedtools intersect -sorted -v -abam stdin -b blacklist | samtools view]"""
from __future__ import annotations

import gzip
import re

import numpy as np

_REF_OPS = re.compile(rb"(\d+)([MIDNSHP=X])")


def ref_span(cigar: bytes) -> int:
    """reference bases covered (M, D, N, =, X), at least 1 (htslib bam_endpos)"""
    n = sum(int(k) for k, op in _REF_OPS.findall(cigar) if op in b"MDN=X")
    return n if n > 0 else 1


def load_bed_intervals(path: str) -> dict[str, tuple[np.ndarray, np.ndarray]]:
    """BED (plain or gzip; `#`, `track`, `browser` lines skipped) -> per chromosome sorted, merged 0-based half-open
    intervals as (starts, ends) int64 arrays"""
    op = gzip.open if open(path, "rb").read(2) == b"\x1f\x8b" else open
    per: dict[str, list[tuple[int, int]]] = {}
    with op(path, "rt") as f:
        for l in f:
            if not l.strip() or l.startswith(("#", "track", "browser")):
                continue
            t = l.split("\t") if "\t" in l else l.split()
            if len(t) < 3 or not t[1].strip().lstrip("-").isdigit():
                continue                                            # header row
            per.setdefault(t[0], []).append((int(t[1]), int(t[2])))
    out = {}
    for c, iv in per.items():
        iv.sort()
        s, e = [], []
        for a, b in iv:
            if b <= a:
                continue
            if s and a <= e[-1]:                                    # overlapping or book-ended: merge (a read overlapping either overlaps the union)
                e[-1] = max(e[-1], b)
            else:
                s.append(a); e.append(b)
        out[c] = (np.array(s, np.int64), np.array(e, np.int64))
    return out


def overlaps_any(iv: tuple[np.ndarray, np.ndarray], pos0: int, span: int) -> bool:
    k = int(np.searchsorted(iv[1], pos0, side="right"))             # first interval ending after pos0
    return k < iv[1].size and iv[0][k] < pos0 + span


def parse_region_str(region: str | None):
    """'chr' -> (chr, 0, 0); 'chr:beg-end' -> (chr, beg, end) (1-based closed)"""
    if not region:
        return None, 0, 0
    if ":" not in region:
        return region, 0, 0
    c, r = region.split(":")
    b, e = r.replace(",", "").split("-")
    return c, int(b), int(e)


def filter_sam(sam: bytes, mapq: int = 0, exclude_flags: int = 0, include_flags: int | None = None, *, chrom: str | None = None,
               beg: int = 0, end: int = 0, flag_eq=(), read_group: str | None = None, intervals=None, exclude_intervals: bool = False,
               max_records: int = 0, key_window: tuple[int, int] | None = None) -> bytes:
    """same filters as BamFile.view, on SAM text (header lines dropped)"""
    keep = []
    rg = (b"RG:Z:" + read_group.encode()) if read_group else None
    cb = chrom.encode() if chrom else None
    for l in sam.splitlines(keepends=True):
        if l.startswith(b"@"):
            continue
        t = l.rstrip(b"\r\n").split(b"\t")
        if len(t) < 6:
            continue
        f, q = int(t[1]), int(t[4])
        if q < mapq or (f & exclude_flags) or (include_flags and (f & include_flags) != include_flags):
            continue
        if cb is not None and t[2] != cb:
            continue
        if flag_eq and f not in flag_eq:
            continue
        if key_window is not None:                                  # template window: max(POS, PNEXT) of a pair on one reference
            pos0 = int(t[3]) - 1; key = pos0
            if (f & 1) and not (f & 8) and len(t) > 7 and t[6] in (b"=", t[2]) and int(t[7]) - 1 > pos0:
                key = int(t[7]) - 1
            if not key_window[0] <= key < key_window[1]:
                continue
        if end > 0 or intervals is not None:
            pos0 = int(t[3]) - 1; span = ref_span(t[5])
            if end > 0 and (pos0 + 1 > end or pos0 + span < beg):
                continue
            if intervals is not None and overlaps_any(intervals, pos0, span) == bool(exclude_intervals):
                continue
        if rg is not None and rg not in t[11:]:
            continue
        keep.append(l if l.endswith(b"\n") else l + b"\n")
        if max_records and len(keep) >= max_records:
            break
    return b"".join(keep)
