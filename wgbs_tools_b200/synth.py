"""Seeded synthetic inputs for the hot path.

Everything here is deterministic in ``seed`` (``numpy.random.default_rng``) so that tests, the bench and the
CPU baseline see byte-identical inputs.  Nothing reads /root/reference.

Formats follow the reference's on-disk definitions:
  * CpG dictionary  ``chr \\t locus \\t idx``  (reference src/python/init_genome.py:246-260, 1-based locus of the C,
    1-based running index)
  * SAM text as ``samtools view`` prints it (the stdin of reference src/pipeline_wgbs/patter.cpp:381)
  * pat text ``chr \\t idx \\t pattern \\t count`` (reference docs/pat_format.md:3-47)
  * beta ``uint8[N,2]`` (reference docs/beta_format.md:3-10)
"""
from __future__ import annotations

import numpy as np

_A, _C, _G, _T, _N = (ord(c) for c in "ACGTN")


class Genome:
    """One synthetic chromosome: bases (1-based, index 0 unused), CpG loci, per-CpG methylation probability."""

    def __init__(self, chrom: str, length: int, loci: np.ndarray, first_idx: int, bases: np.ndarray | None,
                 meth_p: np.ndarray):
        self.chrom = chrom
        self.length = int(length)
        self.loci = loci.astype(np.uint32)
        self.first_idx = int(first_idx)
        self.bases = bases
        self.meth_p = meth_p

    @property
    def n_cpg(self) -> int:
        return int(self.loci.size)

    def idx(self) -> np.ndarray:
        return np.arange(self.first_idx, self.first_idx + self.n_cpg, dtype=np.int64)

    def dict_text(self) -> bytes:
        """``chr\\tlocus\\tidx`` lines (plain text form of CpG.bed.gz)."""
        c = self.chrom.encode()
        return b"".join(b"%s\t%d\t%d\n" % (c, l, i) for l, i in zip(self.loci.tolist(), self.idx().tolist()))


def make_loci(rng, length: int, n_target: int | None = None, density: float = 0.01, island_frac: float = 0.01,
              island_density: float = 0.1) -> np.ndarray:
    """Sorted CpG loci (1-based position of the C), pairwise >= 2 apart, last locus <= length-1."""
    slots = (length - 2) // 2                       # candidate C positions 1,3,5,... so CpGs never overlap
    p = np.full(slots, density * 2, dtype=np.float32)
    n_isl = max(1, int(length * island_frac / 1000))
    starts = rng.integers(0, max(1, slots - 500), size=n_isl)
    for s in starts.tolist():
        p[s:s + 500] = island_density * 2
    if n_target is not None:
        p *= n_target / float(p.sum())
    pick = rng.random(slots, dtype=np.float32) < p
    loci = (np.flatnonzero(pick).astype(np.int64) * 2 + 1 + rng.integers(0, 2))
    loci = loci[loci <= length - 1]
    if n_target is not None:                        # hit the target exactly (tests rely on N)
        if loci.size > n_target:
            loci = np.sort(rng.choice(loci, size=n_target, replace=False))
        while loci.size < n_target:
            extra = rng.integers(1, length - 1, size=(n_target - loci.size) * 2)
            allc = np.unique(np.concatenate([loci, extra]))
            keep = np.ones(allc.size, bool)
            keep[1:] = np.diff(allc) >= 2
            # enforce gap >= 2 greedily
            out = [int(allc[0])]
            for v in allc[1:].tolist():
                if v - out[-1] >= 2:
                    out.append(v)
            loci = np.array(out[:n_target] if len(out) >= n_target else out, dtype=np.int64)
    return loci


def make_genome(seed: int, chrom: str = "chrT", length: int = 5_000_000, n_cpg: int | None = None,
                first_idx: int = 1, with_bases: bool = True) -> Genome:
    rng = np.random.default_rng(seed)
    loci = make_loci(rng, length, n_target=n_cpg)
    # regional methylation: one Beta(0.5,0.5) draw per 1 kb window
    region_p = rng.beta(0.5, 0.5, size=length // 1000 + 2).astype(np.float32)
    meth_p = region_p[loci // 1000]
    bases = None
    if with_bases:
        # no G anywhere first => no accidental "CG"; then G's where the previous base is not C; then the CpGs
        bases = np.array([_A, _C, _T], dtype=np.uint8)[rng.integers(0, 3, size=length + 2)]
        g = (rng.random(length + 2) < 0.25)
        g[1:] &= bases[:-1] != _C
        g[0] = False
        bases[g] = _G
        # a G we just placed may sit right after a position that we later turn into C (a locus) - fine, that IS the CpG
        nxt = loci + 1
        bases[loci] = _C
        bases[nxt] = _G
        # kill accidental CGs created by G placed after an original C that was itself overwritten: re-scan
        acc = np.flatnonzero((bases[:-1] == _C) & (bases[1:] == _G))
        is_locus = np.zeros(length + 2, bool)
        is_locus[loci] = True
        bad = acc[~is_locus[acc]]
        bases[bad + 1] = _A
        bases[0] = _N
    return Genome(chrom, length, loci, first_idx, bases, meth_p)


# ----------------------------------------------------------------------------------------------------------------
# bisulfite reads -> SAM text
# ----------------------------------------------------------------------------------------------------------------

def _convert(g: Genome, rng, start: np.ndarray, bottom: np.ndarray, rlen: int, cpg_at: np.ndarray,
             mp_at: np.ndarray) -> np.ndarray:
    """Forward-strand bases of reads [start, start+rlen) after bisulfite conversion of their strand."""
    pos = start[:, None] + np.arange(rlen)[None, :]
    b = g.bases[pos]
    r = rng.random(pos.shape, dtype=np.float32)
    top = ~bottom[:, None]
    # OT: C -> T unless methylated CpG C (non-CpG C converts w.p. 0.995)
    isC = (b == _C) & top
    keepC = np.where(cpg_at[pos], r < mp_at[pos], r >= 0.995)
    b = np.where(isC & ~keepC, _T, b)
    # OB: G -> A unless it is the G of a methylated CpG
    isG = (b == _G) & ~top
    pm1 = pos - 1
    keepG = np.where(cpg_at[pm1], r < mp_at[pm1], r >= 0.995)
    b = np.where(isG & ~keepG, _A, b)
    return b.astype(np.uint8)


def make_sam(g: Genome, n_reads: int, seed: int, paired: bool = True, rlen: int = 150, indel_frac: float = 0.03,
             single_frac: float = 0.01, qual: bool = True, name_prefix: str = "r") -> bytes:
    """Coordinate-sorted SAM text (no header) of ``n_reads`` records on ``g``.

    PE: flags 99/147 (OT) or 83/163 (OB); SE: 0/16.  3 % of records carry one I / D / S event,
    1 % of PE templates lose a mate (singletons).  QNAMEs are unique per template.
    """
    rng = np.random.default_rng(seed)
    L = g.length
    cpg_at = np.zeros(L + 2, bool)
    cpg_at[g.loci] = True
    mp_at = np.zeros(L + 2, np.float32)
    mp_at[g.loci] = g.meth_p
    chrom = g.chrom.encode()
    recs = []  # (pos, line)
    if paired:
        n_t = (n_reads + 1) // 2
        ins = np.clip(rng.normal(300, 50, n_t), rlen, 600).astype(np.int64)
        t0 = rng.integers(1, L - 700, size=n_t)
        bottom = rng.random(n_t) < 0.5
        p1 = t0                                   # leftmost mate
        p2 = t0 + ins - rlen                      # rightmost mate
        s1 = _convert(g, rng, p1, bottom, rlen, cpg_at, mp_at)
        s2 = _convert(g, rng, p2, bottom, rlen, cpg_at, mp_at)
        # OT template: left mate is read1 fwd (99), right is read2 rev (147). OB: left is read2 fwd (163), right read1 rev (83)
        fl_left = np.where(bottom, 163, 99)
        fl_right = np.where(bottom, 83, 147)
        drop = rng.random(n_t) < single_frac
        drop_left = rng.random(n_t) < 0.5
        starts = np.concatenate([p1, p2]); seqs = np.concatenate([s1, s2]); flags = np.concatenate([fl_left, fl_right])
        mates = np.concatenate([p2, p1]); tid = np.concatenate([np.arange(n_t), np.arange(n_t)])
        tlen = np.concatenate([ins, -ins])
        keep = np.concatenate([~(drop & drop_left), ~(drop & ~drop_left)])
    else:
        n_t = n_reads
        starts = rng.integers(1, L - 700, size=n_t)
        bottom = rng.random(n_t) < 0.5
        seqs = _convert(g, rng, starts, bottom, rlen, cpg_at, mp_at)
        flags = np.where(bottom, 16, 0); mates = np.zeros(n_t, np.int64); tid = np.arange(n_t); tlen = np.zeros(n_t, np.int64)
        keep = np.ones(n_t, bool)
    n = starts.size
    ev = rng.random(n) < indel_frac
    ev_kind = rng.integers(0, 4, size=n)          # 0 I, 1 D, 2 S-left, 3 S-right
    ev_len = rng.integers(1, 9, size=n)
    ev_at = rng.integers(10, rlen - 20, size=n)
    order = np.argsort(starts, kind="stable")
    seq_s = seqs.view("S%d" % rlen).ravel()
    q = b"F" * rlen if qual else b"*"
    rnd_bases = np.frombuffer(b"ACGT", np.uint8)
    out = []
    pfx = name_prefix.encode()
    for i in order.tolist():
        if not keep[i]:
            continue
        pos = int(starts[i]); s = seq_s[i]; cigar = b"%dM" % rlen
        if len(s) < rlen:                         # numpy strips trailing NULs only; bases are never NUL
            s = s.ljust(rlen, b"N")
        if ev[i]:
            k = int(ev_len[i]); a = int(ev_at[i]); kind = int(ev_kind[i])
            if kind == 0:                         # insertion: a M, k I, rest M  (read stays rlen long)
                insb = rnd_bases[rng.integers(0, 4, size=k)].tobytes()
                s = s[:a] + insb + s[a:rlen - k]
                cigar = b"%dM%dI%dM" % (a, k, rlen - a - k)
            elif kind == 1:                       # deletion: k reference bases after offset a are skipped
                s = s[:a] + s[a + k:] + g.bases[pos + rlen: pos + rlen + k].tobytes()
                cigar = b"%dM%dD%dM" % (a, k, rlen - a)
            elif kind == 2:                       # soft clip left: first k read bases are junk, POS moves right
                s = rnd_bases[rng.integers(0, 4, size=k)].tobytes() + s[k:]
                pos = pos + k
                cigar = b"%dS%dM" % (k, rlen - k)
            else:
                s = s[:rlen - k] + rnd_bases[rng.integers(0, 4, size=k)].tobytes()
                cigar = b"%dM%dS" % (rlen - k, k)
        if paired:
            rnext, pnext, tl = b"=", int(mates[i]), int(tlen[i])
        else:
            rnext, pnext, tl = b"*", 0, 0
        out.append(b"%s%d\t%d\t%s\t%d\t60\t%s\t%s\t%d\t%d\t%s\t%s\tNM:i:0\n" % (
            pfx, tid[i], flags[i], chrom, pos, cigar, rnext, pnext, tl, s, q))
        recs.append(pos)
    # soft clips shift POS: restore coordinate order (stable)
    o2 = np.argsort(np.array(recs, dtype=np.int64), kind="stable")
    return b"".join(out[j] for j in o2.tolist())


def make_np_sam(g: Genome, n_reads: int, seed: int, rlen: int = 150, dot_frac: float = 0.3) -> bytes:
    """SE reads with MM:Z:C+m?/C+h? (or the implicit '.' convention) + ML:B:C tags on CpG-context C's
    (modification-aware, un-converted bases) -- the MM/ML branch of reference src/pipeline_wgbs/ont.cpp."""
    rng = np.random.default_rng(seed)
    L = g.length
    starts = np.sort(rng.integers(1, L - 700, size=n_reads))
    chrom = g.chrom.encode()
    comp = bytes.maketrans(b"ACGTN", b"TGCAN")
    out = []
    for i in range(n_reads):
        pos = int(starts[i]); bottom = bool(rng.random() < 0.5)
        fwd = g.bases[pos:pos + rlen].tobytes()
        cigar = b"%dM" % rlen
        r = rng.random()
        if r < 0.03:                               # one deletion
            a = int(rng.integers(10, rlen - 20)); k = int(rng.integers(1, 6))
            fwd = fwd[:a] + g.bases[pos + a + k: pos + k + rlen].tobytes()
            cigar = b"%dM%dD%dM" % (a, k, rlen - a)
        elif r < 0.06:                             # one insertion
            a = int(rng.integers(10, rlen - 20)); k = int(rng.integers(1, 6))
            fwd = fwd[:a] + bytes(rng.choice(list(b"ACGT"), size=k).tolist()) + fwd[a:rlen - k]
            cigar = b"%dM%dI%dM" % (a, k, rlen - a - k)
        elif r < 0.08:
            k = int(rng.integers(1, 9))
            cigar = b"%dS%dM" % (k, rlen - k); pos += k
        orig = fwd.translate(comp)[::-1] if bottom else fwd   # read as sequenced
        # C's in the original orientation; modification calls on those in CpG context
        cs = [j for j, ch in enumerate(orig) if ch == _C]
        cpg_c = [n for n, j in enumerate(cs) if j + 1 < len(orig) and orig[j + 1] == _G]
        dot = rng.random() < dot_frac
        keep = [n for n in cpg_c if rng.random() < 0.9]
        tags = b""
        if keep:
            deltas = []; prev = -1
            for n in keep:
                deltas.append(n - prev - 1); prev = n
            d = b",".join(b"%d" % x for x in deltas)
            mlm = rng.integers(0, 256, size=len(keep)); mlh = rng.integers(0, 256, size=len(keep))
            suff = b"." if dot else b"?"
            style = rng.integers(0, 4)
            if style == 0:     # m and h
                tags = b"\tMM:Z:C+m%s,%s;C+h%s,%s;\tML:B:C,%s,%s" % (suff, d, suff, d, b",".join(b"%d" % x for x in mlm), b",".join(b"%d" % x for x in mlh))
            elif style == 1:   # h first
                tags = b"\tMM:Z:C+h%s,%s;C+m%s,%s;\tML:B:C,%s,%s" % (suff, d, suff, d, b",".join(b"%d" % x for x in mlh), b",".join(b"%d" % x for x in mlm))
            elif style == 2:   # m only
                tags = b"\tMM:Z:C+m%s,%s;\tML:B:C,%s" % (suff, d, b",".join(b"%d" % x for x in mlm))
            else:              # m only, no ML (implicit 255) -- Biomodal-like
                tags = b"\tMm:Z:C+m%s,%s;" % (suff, d)
        elif rng.random() < 0.5:
            tags = b"\tMM:Z:C+m.;" if dot else b"\tMM:Z:C+m?;"
        out.append(b"np%d\t%d\t%s\t%d\t60\t%s\t*\t0\t0\t%s\t%s\tNM:i:0%s\n" % (
            i, 16 if bottom else 0, chrom, pos, cigar, fwd, b"F" * len(fwd), tags))
    o = np.argsort(np.array([int(l.split(b"\t", 4)[3]) for l in out]), kind="stable")
    return b"".join(out[j] for j in o.tolist())


# ----------------------------------------------------------------------------------------------------------------
# pat / beta / blocks
# ----------------------------------------------------------------------------------------------------------------

def make_pat_records(seed: int, n_reads: int, n_cpg: int, first_idx: int = 1, mean_len: float = 4.0, max_len: int = 30):
    """(idx int64[R], pattern list/array, count int64[R]) sorted by (idx, pattern) and collapsed.

    Returned as numpy: idx[R], off[R+1] into a uint8 symbol pool of ASCII chars, count[R]."""
    rng = np.random.default_rng(seed)
    region_p = rng.beta(0.5, 0.5, size=n_cpg // 50 + 2).astype(np.float32)
    lens = np.minimum(rng.geometric(1.0 / mean_len, size=n_reads), max_len).astype(np.int64)
    start = rng.integers(0, n_cpg, size=n_reads)
    lens = np.minimum(lens, n_cpg - start)
    off = np.zeros(n_reads + 1, np.int64); off[1:] = np.cumsum(lens)
    tot = int(off[-1])
    rid = np.repeat(np.arange(n_reads), lens)
    site = start[rid] + (np.arange(tot) - off[rid])
    r = rng.random(tot, dtype=np.float32); r2 = rng.random(tot, dtype=np.float32)
    sym = np.where(r < region_p[site // 50], ord("C"), ord("T")).astype(np.uint8)
    sym[r2 < 0.03] = ord(".")
    sym[(r2 >= 0.03) & (r2 < 0.04)] = ord("H")
    # first / last symbol must not be '.'
    first = off[:-1]; last = off[1:] - 1
    sym[first] = np.where(sym[first] == ord("."), ord("T"), sym[first])
    sym[last] = np.where(sym[last] == ord("."), ord("C"), sym[last])
    pats = [sym[off[i]:off[i + 1]].tobytes() for i in range(n_reads)]
    idx = start + first_idx
    order = sorted(range(n_reads), key=lambda i: (int(idx[i]), pats[i]))
    out_idx, out_pat, out_cnt = [], [], []
    for i in order:
        if out_idx and out_idx[-1] == idx[i] and out_pat[-1] == pats[i]:
            out_cnt[-1] += 1
        else:
            out_idx.append(int(idx[i])); out_pat.append(pats[i]); out_cnt.append(1)
    return np.array(out_idx, np.int64), out_pat, np.array(out_cnt, np.int64)


def pat_text(chrom: str, idx, pats, counts) -> bytes:
    c = chrom.encode()
    return b"".join(b"%s\t%d\t%s\t%d\n" % (c, i, p, n) for i, p, n in zip(idx.tolist(), pats, counts.tolist()))


def make_pat_text_fast(seed: int, n_reads: int, n_cpg: int, chrom: str = "chr1", first_idx: int = 1, mean_len: float = 4.0,
                       max_len: int = 30, max_count: int = 3) -> bytes:
    """Large sorted pat text without the Python-level collapse (counts drawn, duplicates allowed - the
    consumers (pat2beta, homog) do not require uniqueness)."""
    rng = np.random.default_rng(seed)
    region_p = rng.beta(0.5, 0.5, size=n_cpg // 50 + 2).astype(np.float32)
    lens = np.minimum(rng.geometric(1.0 / mean_len, size=n_reads), max_len).astype(np.int64)
    start = np.sort(rng.integers(0, n_cpg, size=n_reads))
    lens = np.maximum(1, np.minimum(lens, n_cpg - start))
    off = np.zeros(n_reads + 1, np.int64); off[1:] = np.cumsum(lens)
    tot = int(off[-1])
    rid = np.repeat(np.arange(n_reads), lens)
    site = start[rid] + (np.arange(tot) - off[rid])
    r = rng.random(tot, dtype=np.float32); r2 = rng.random(tot, dtype=np.float32)
    sym = np.where(r < region_p[site // 50], ord("C"), ord("T")).astype(np.uint8)
    sym[r2 < 0.03] = ord("."); sym[(r2 >= 0.03) & (r2 < 0.04)] = ord("H")
    first = off[:-1]; last = off[1:] - 1
    sym[first] = np.where(sym[first] == ord("."), ord("T"), sym[first])
    sym[last] = np.where(sym[last] == ord("."), ord("C"), sym[last])
    cnt = rng.integers(1, max_count + 1, size=n_reads)
    c = chrom.encode() + b"\t"
    idx = (start + first_idx).tolist(); cl = cnt.tolist(); o = off.tolist(); sb = sym.tobytes()
    return b"".join(b"%s%d\t%s\t%d\n" % (c, idx[i], sb[o[i]:o[i + 1]], cl[i]) for i in range(n_reads))


def make_betas(seed: int, n_files: int, n_sites: int, n_tissues: int = 20, plateau: int = 50) -> list[np.ndarray]:
    """uint8[n_sites,2] arrays: cov ~ Poisson(30) clipped to 0..255, meth ~ Binomial(cov, p) with p piecewise-constant
    per tissue (20 'tissues' x replicates)."""
    rng = np.random.default_rng(seed)
    nt = min(n_tissues, n_files)
    n_pl = n_sites // plateau + 2
    base = rng.beta(0.5, 0.5, size=n_pl)
    tissue_p = []
    for _ in range(nt):
        p = base.copy()
        flip = rng.random(n_pl) < 0.1
        p[flip] = rng.beta(0.5, 0.5, size=int(flip.sum()))
        tissue_p.append(np.repeat(p, plateau)[:n_sites])
    out = []
    for f in range(n_files):
        cov = np.clip(rng.poisson(30, size=n_sites), 0, 255)
        meth = rng.binomial(cov, tissue_p[f % nt])
        out.append(np.stack([meth, cov], axis=1).astype(np.uint8))
    return out


def make_blocks(seed: int, first_idx: int, n_cpg: int, mean_len: float = 8.0) -> np.ndarray:
    """int32[B,2] (startCpG, endCpG) tiling [first_idx, first_idx+n_cpg) with Geometric(mean_len) block sizes."""
    rng = np.random.default_rng(seed)
    sizes = rng.geometric(1.0 / mean_len, size=int(n_cpg / mean_len * 1.5) + 16)
    edges = np.concatenate([[0], np.cumsum(sizes)])
    edges = edges[edges < n_cpg]
    edges = np.concatenate([edges, [n_cpg]]) + first_idx
    return np.stack([edges[:-1], edges[1:]], axis=1).astype(np.int32)


def blocks_text(chrom: str, blocks: np.ndarray, loci: np.ndarray | None = None, first_idx: int = 1) -> bytes:
    c = chrom.encode()
    out = []
    for s, e in blocks.tolist():
        if loci is not None:
            a = int(loci[s - first_idx]); b = int(loci[e - 1 - first_idx]) + 1
        else:
            a, b = s * 100, e * 100
        out.append(b"%s\t%d\t%d\t%d\t%d\n" % (c, a, b, s, e))
    return b"".join(out)
