"""CSI index of a .pat.gz, and region reads through it.

The reference indexes every pat file with `tabix -Cf -b 2 -e 2 -m 12 X.pat.gz` (src/python/index.py:12-19,83-93): a CSI
index (min_shift 12) over column 2, the CpG index, with column 1 as the sequence name.  tabix is not available here, so
this module writes the same structure from the SAM/tabix specification (CSIv1, section 5.3 of the SAM spec; tabix aux
block) following htslib's construction rules:
  * one interval [idx-1, idx) per line; bin = reg2bin at depth n_lvls; a chunk per maximal run of consecutive records
    falling into the same bin, from the virtual offset of its first record to the virtual offset behind its last;
  * n_lvls: htslib >= 1.11 deepens the binning until it spans 100 Gbp when the file names no sequence lengths
    (tbx.c adjust_n_lvls): 9 levels for min_shift 12;
  * the pseudo-bin (n_bins + 1) with the sequence's file range and record count; per-bin `loff` from the 2^min_shift linear
    index; bins whose chunks span < 64 KiB of compressed file are merged into an existing parent; adjacent chunks sharing
    a BGZF block are merged.
Bins are written in ascending bin number (htslib writes them in its hash-table order: readers do not care, so index BYTES
are not comparable with tabix's -- index CONTENT is).  Parity is therefore pinned by round trip: every region query
through the index returns exactly the lines a full scan returns (tests/test_host_logic.py).

No GPU is involved: this is the on-disk format next to the hot path (docs/pat_format.md:43-47)."""
from __future__ import annotations

import struct
import zlib

import numpy as np

from .patio import BGZF_EOF, bgzf_compress

MIN_SHIFT = 12
MAX_REF_LEN = 100 * 1024 ** 3            # htslib's default when no sequence lengths are known
MIN_MARKER_DIST = 0x10000


def n_levels(min_shift: int = MIN_SHIFT, max_len: int = MAX_REF_LEN) -> int:
    """tabix -m MIN_SHIFT: (31 - min_shift + 2) / 3 levels, deepened until 2^(min_shift + 3 n) covers max_len + 256"""
    n = (31 - min_shift + 2) // 3
    s = 1 << (min_shift + 3 * n)
    while max_len + 256 > s:
        n += 1; s <<= 3
    return n


def bgzf_blocks(raw: bytes):
    """[(compressed offset, compressed size, uncompressed size)] of every BGZF block"""
    out = []; off = 0; n = len(raw)
    while off + 18 <= n:
        if raw[off:off + 4] != b"\x1f\x8b\x08\x04":
            raise ValueError(f"not a BGZF file (bad block header at {off})")
        xlen = struct.unpack_from("<H", raw, off + 10)[0]
        x = off + 12; bsize = None
        while x + 4 <= off + 12 + xlen:
            si1, si2, slen = raw[x], raw[x + 1], struct.unpack_from("<H", raw, x + 2)[0]
            if si1 == 66 and si2 == 67 and slen == 2:
                bsize = struct.unpack_from("<H", raw, x + 4)[0] + 1
            x += 4 + slen
        if bsize is None or off + bsize > n:
            raise ValueError(f"corrupt BGZF block at {off}")
        out.append((off, bsize, struct.unpack_from("<I", raw, off + bsize - 4)[0]))
        off += bsize
    return out


def _inflate(raw: bytes, block) -> bytes:
    off, bsize, _ = block
    xlen = struct.unpack_from("<H", raw, off + 10)[0]
    return zlib.decompress(raw[off + 12 + xlen: off + bsize - 8], -15)


def reg2bin(beg: np.ndarray, end: np.ndarray, min_shift: int, n_lvls: int) -> np.ndarray:
    """hts_reg2bin, vectorised: the smallest bin containing [beg, end)"""
    beg = np.asarray(beg, np.int64); end = np.asarray(end, np.int64) - 1
    out = np.zeros(beg.shape, np.int64); done = np.zeros(beg.shape, bool)
    s = min_shift; t = ((1 << (3 * n_lvls)) - 1) // 7
    for l in range(n_lvls, 0, -1):
        hit = ~done & ((beg >> s) == (end >> s))
        out[hit] = t + (beg[hit] >> s); done |= hit
        s += 3; t -= 1 << (3 * (l - 1))
    return out


def _bin_level(b: int) -> int:
    l = 0
    while b:
        b = (b - 1) >> 3; l += 1
    return l


def _bin_first(l: int) -> int:
    return ((1 << (3 * l)) - 1) // 7


def reg2bins(beg: int, end: int, min_shift: int, n_lvls: int) -> list[int]:
    """every bin that may hold records overlapping [beg, end)"""
    out = []; end -= 1
    s = min_shift + 3 * n_lvls; t = 0
    for l in range(n_lvls + 1):
        b, e = t + (beg >> s), t + (end >> s)
        out += range(b, e + 1)
        s -= 3; t += 1 << (3 * l)
    return out


class CsiIndex:
    def __init__(self, min_shift: int, n_lvls: int, names: list[str], bins: list[dict], n_no_coor: int = 0, conf=(0, 1, 2, 2, ord("#"), 0)):
        self.min_shift, self.n_lvls, self.names, self.bins, self.n_no_coor, self.conf = min_shift, n_lvls, names, bins, n_no_coor, conf

    @property
    def meta_bin(self) -> int:
        return ((1 << (3 * (self.n_lvls + 1))) - 1) // 7 + 1

    # ---- build ------------------------------------------------------------------------------------------------------
    @classmethod
    def build(cls, path: str, min_shift: int = MIN_SHIFT) -> "CsiIndex":
        raw = open(path, "rb").read()
        blocks = bgzf_blocks(raw)
        data = b"".join(_inflate(raw, b) for b in blocks)
        ustart = np.zeros(len(blocks) + 1, np.int64); ustart[1:] = np.cumsum([b[2] for b in blocks])
        coff = np.array([b[0] for b in blocks] + [len(raw)], np.int64)

        def voff(upos: np.ndarray) -> np.ndarray:
            """virtual offset of an uncompressed position, as bgzf_tell reports it: a position at the very end of a block
            is offset 0 of the block that follows it (which may be an empty EOF block between chromosome parts)"""
            upos = np.asarray(upos, np.int64)
            k = np.searchsorted(ustart, upos, side="left")
            k = np.where(ustart[np.minimum(k, len(blocks))] == upos, k, k - 1)
            return (coff[k] << 16) | (upos - ustart[k])
        n_lvls = n_levels(min_shift)
        a = np.frombuffer(data, np.uint8)
        nl = np.flatnonzero(a == 10)
        ends = nl if (a.size == 0 or a[-1] == 10) else np.append(nl, a.size)
        starts = np.concatenate([[0], ends[:-1] + 1]) if ends.size else np.zeros(0, np.int64)
        keep = (ends > starts) & (a[np.minimum(starts, max(a.size - 1, 0))] != ord("#")) if ends.size else np.zeros(0, bool)
        starts, ends = starts[keep], ends[keep]                      # empty lines and '#' meta lines are not records
        if starts.size == 0:
            return cls(min_shift, n_lvls, [], [])
        tabs = np.flatnonzero(a == 9)
        k1 = np.searchsorted(tabs, starts)
        if tabs.size < 2 or k1.max() + 1 >= tabs.size or np.any(tabs[k1 + 1] >= ends):
            raise ValueError("pat line with fewer than 3 columns")
        t1, t2 = tabs[k1], tabs[k1 + 1]
        # column 2 as integer
        val = np.zeros(starts.size, np.int64); ln = t2 - t1 - 1
        if ln.min() < 1 or ln.max() > 18:
            raise ValueError("bad CpG index column")
        for d in range(int(ln.max())):
            m = ln > d
            c = a[t1[m] + 1 + d].astype(np.int64) - 48
            if np.any((c < 0) | (c > 9)):
                raise ValueError("non-numeric CpG index")
            val[m] = val[m] * 10 + c
        beg = np.maximum(val - 1, 0); end = np.maximum(val, 1)
        # sequence names: contiguous runs (anything else is "chromosome blocks not continuous" for tabix, too)
        names, tid = [], np.zeros(starts.size, np.int32)
        i = 0; n = starts.size
        name_of = lambda j: data[starts[j]:t1[j]]
        while i < n:
            nm = name_of(i); lo, step = i, 1
            while lo + step < n and name_of(lo + step) == nm:
                lo += step; step *= 2
            hi = min(n - 1, lo + step)
            while lo < hi:                                            # last line of the run: name(lo) == nm, name(hi) may differ
                mid = (lo + hi + 1) // 2
                if name_of(mid) == nm:
                    lo = mid
                else:
                    hi = mid - 1
            if nm.decode() in names:
                raise ValueError("chromosome blocks not continuous")
            tid[i:lo + 1] = len(names); names.append(nm.decode())
            i = lo + 1
        for t in range(len(names)):
            b = beg[tid == t]
            if np.any(np.diff(b) < 0):
                raise ValueError("unsorted positions")
        rec_off = voff(starts)                                       # start of every record
        bins_of = reg2bin(beg, end, min_shift, n_lvls)
        out = []
        file_end = len(raw) << 16                                    # bgzf_tell once the reader has run off the end
        first_off = int(rec_off[0])
        for t in range(len(names)):
            sel = np.flatnonzero(tid == t); b = bins_of[sel]
            cut = np.flatnonzero(np.diff(b) != 0) + 1                # a chunk per run of equal bins
            run_s = np.concatenate([[0], cut]); run_e = np.concatenate([cut, [sel.size]])
            r0 = rec_off[sel]
            # htslib closes a run at the START offset of the record that follows it (idx->z.last_off), which for the
            # last run of a sequence is the first record of the next sequence or the end of the file
            nxt = np.concatenate([r0[1:], [rec_off[sel[-1] + 1] if sel[-1] + 1 < n else file_end]])
            bd: dict[int, dict] = {}
            for s_, e_ in zip(run_s.tolist(), run_e.tolist()):
                bd.setdefault(int(b[s_]), {"loff": 0, "chunks": []})["chunks"].append((int(r0[s_]), int(nxt[e_ - 1])))
            seq_beg = int(r0[0]) if t else first_off
            seq_end = int(nxt[-1])
            # linear index (2^min_shift windows): offset of the first record overlapping each window, gaps filled backwards
            win = (beg[sel] >> min_shift)
            uw, first_i = np.unique(win, return_index=True)
            lin = np.full(int(uw.max()) + 1, -1, np.int64); lin[uw] = r0[first_i]
            # update_loff: leading empty windows take the sequence's first offset, later gaps the previous window's
            prev = seq_beg
            for w in range(lin.size):
                if lin[w] < 0:
                    lin[w] = prev
                prev = lin[w]
            for bn, rec in bd.items():
                l = _bin_level(bn)
                bot = (bn - _bin_first(l)) << ((n_lvls - l) * 3)
                rec["loff"] = int(lin[bot]) if bot < lin.size else 0
            cls._compress(bd, n_lvls)
            meta = ((1 << (3 * (n_lvls + 1))) - 1) // 7 + 1
            bd[meta] = {"loff": 0, "chunks": [(seq_beg, seq_end), (int(sel.size), 0)]}
            out.append(bd)
        return cls(min_shift, n_lvls, names, out)

    @staticmethod
    def _compress(bd: dict, n_lvls: int):
        """hts.c compress_binning: small bins into an existing parent, then adjacent chunks in one BGZF block"""
        for l in range(n_lvls, 0, -1):
            start = _bin_first(l)
            for bn in sorted(k for k in bd if k >= start and _bin_level(k) == l):
                ch = bd[bn]["chunks"]
                if l < n_lvls:
                    ch.sort()
                if (ch[-1][1] >> 16) - (ch[0][0] >> 16) < MIN_MARKER_DIST:
                    par = (bn - 1) >> 3
                    if par in bd:
                        bd[par]["chunks"] += ch
                        del bd[bn]
        if 0 in bd:
            bd[0]["chunks"].sort()
        for rec in bd.values():
            ch = rec["chunks"]; m = [ch[0]]
            for u, v in ch[1:]:
                if (m[-1][1] >> 16) >= (u >> 16):
                    if m[-1][1] < v:
                        m[-1] = (m[-1][0], v)
                else:
                    m.append((u, v))
            rec["chunks"] = m

    # ---- file -------------------------------------------------------------------------------------------------------
    def to_bytes(self) -> bytes:
        nm = b"".join(s.encode() + b"\0" for s in self.names)
        aux = struct.pack("<7i", *self.conf, len(nm)) + nm
        out = [b"CSI\1", struct.pack("<3i", self.min_shift, self.n_lvls, len(aux)), aux, struct.pack("<i", len(self.names))]
        for bd in self.bins:
            out.append(struct.pack("<i", len(bd)))
            for bn in sorted(bd):
                rec = bd[bn]
                out.append(struct.pack("<IQi", bn, rec["loff"], len(rec["chunks"])))
                out += [struct.pack("<QQ", u, v) for u, v in rec["chunks"]]
        out.append(struct.pack("<Q", self.n_no_coor))
        return bgzf_compress(b"".join(out), threads=1)

    def save(self, path: str):
        with open(path, "wb") as f:
            f.write(self.to_bytes())

    @classmethod
    def load(cls, path: str) -> "CsiIndex":
        raw = open(path, "rb").read()
        d = b"".join(_inflate(raw, b) for b in bgzf_blocks(raw))
        if d[:4] != b"CSI\1":
            raise ValueError(f"{path}: not a CSI index")
        min_shift, n_lvls, l_aux = struct.unpack_from("<3i", d, 4)
        aux = d[16:16 + l_aux]; p = 16 + l_aux
        conf = struct.unpack_from("<6i", aux, 0); l_nm = struct.unpack_from("<i", aux, 24)[0]
        names = [s.decode() for s in aux[28:28 + l_nm].split(b"\0") if s]
        n_ref = struct.unpack_from("<i", d, p)[0]; p += 4
        bins = []
        for _ in range(n_ref):
            n_bin = struct.unpack_from("<i", d, p)[0]; p += 4
            bd = {}
            for _ in range(n_bin):
                bn, loff, nch = struct.unpack_from("<IQi", d, p); p += 16
                bd[bn] = {"loff": loff, "chunks": [struct.unpack_from("<QQ", d, p + 16 * j) for j in range(nch)]}
                p += 16 * nch
            bins.append(bd)
        n_no = struct.unpack_from("<Q", d, p)[0] if p + 8 <= len(d) else 0
        return cls(min_shift, n_lvls, names, bins, n_no, conf)

    # ---- query ------------------------------------------------------------------------------------------------------
    def chunks(self, name: str, beg: int, end: int) -> list[tuple[int, int]]:
        """virtual-offset ranges that may contain records of `name` overlapping the 0-based half-open [beg, end)"""
        if name not in self.names:
            return []
        bd = self.bins[self.names.index(name)]
        # lowest offset a record overlapping `beg` can have: loff of the smallest existing bin containing beg's window
        min_off = 0
        b = _bin_first(self.n_lvls) + (beg >> self.min_shift)
        while True:
            if b in bd:
                min_off = bd[b]["loff"]; break
            if b == 0:
                break
            b = (b - 1) >> 3
        out = []
        for bn in reg2bins(beg, end, self.min_shift, self.n_lvls):
            if bn in bd:
                out += [(u, v) for u, v in bd[bn]["chunks"] if v > min_off]
        out.sort()
        merged = []
        for u, v in out:
            if merged and u <= merged[-1][1]:
                merged[-1] = (merged[-1][0], max(merged[-1][1], v))
            else:
                merged.append((u, v))
        return merged


def index_pat(path: str) -> str:
    """`tabix -Cf -b 2 -e 2 -m 12 path` (reference index.py:83-93): writes path + '.csi'"""
    out = path + ".csi"
    CsiIndex.build(path).save(out)
    return out


def read_region(path: str, chrom: str, lo: int, hi: int, index: CsiIndex | None = None) -> bytes:
    """`tabix path chrom:lo-hi`: the lines of `chrom` whose column 2 lies in the closed range [lo, hi], in file order,
    reading only the BGZF blocks the index points at"""
    idx = index or CsiIndex.load(path + ".csi")
    raw = open(path, "rb").read() if isinstance(path, str) else path
    key = chrom.encode() + b"\t"
    out = []
    cache: dict[int, tuple[bytes, int]] = {}

    def block(coff: int):
        if coff not in cache:
            if coff >= len(raw):
                return b"", 0
            xlen = struct.unpack_from("<H", raw, coff + 10)[0]
            bsize = struct.unpack_from("<H", raw, coff + 16)[0] + 1 if raw[coff + 12:coff + 14] == b"BC" else bgzf_blocks(raw[coff:coff + 65536 + 64])[0][1]
            cache[coff] = (zlib.decompress(raw[coff + 12 + xlen: coff + bsize - 8], -15), bsize)
        return cache[coff]
    for u, v in idx.chunks(chrom, max(lo - 1, 0), hi):
        coff, uoff = u >> 16, u & 0xFFFF
        buf = []
        while coff < len(raw) and (coff << 16) < v:
            d, bsize = block(coff)
            stop = (v & 0xFFFF) if coff == (v >> 16) else len(d)
            buf.append(d[uoff:stop]); coff += bsize; uoff = 0
        text = b"".join(buf)
        for l in text.splitlines(keepends=True):
            if not l.startswith(key):
                continue
            t = l.split(b"\t", 3)
            if lo <= int(t[1]) <= hi:
                out.append(l)
    return b"".join(out)
