"""BAM files: reader (native: csrc/bam.cu, BGZF inflate on host threads) and a small writer used to build test inputs.

`BamFile.view(chrom, ...)` returns the SAM text `samtools view BAM chrom -q Q -F X [-f Y]` prints (reference
bam2pat.py:165) -- the input of the pileup front end."""
from __future__ import annotations

import ctypes as C
import re
import struct

import weakref

import numpy as np

from ._lib import ViewOpts, check, lib
from .patio import bgzf_compress


def view_opts(refs, chrom: str | None = None, mapq: int = 0, exclude_flags: int = 0, include_flags: int = 0, beg: int = 0, end: int = 0,
              flag_eq=(), read_group: str | None = None, intervals=None, exclude_intervals: bool = False, max_records: int = 0,
              key_window: tuple[int, int] | None = None):
    """wgbs_view_opts for one `samtools view` stage; returns (opts, arrays to keep alive) or (None, None) when nothing can pass.
    key_window: 0-based half-open window on the template key max(POS, PNEXT) (see wgbs_view_opts in include/wgbs_b200.h)"""
    vo = ViewOpts()
    vo.refid = -1 if chrom is None else refs.index(chrom)
    vo.min_mapq, vo.exclude_flags, vo.include_flags, vo.beg, vo.end = mapq, exclude_flags, include_flags or 0, beg, end
    vo.n_flag_eq = len(flag_eq)
    for i, f in enumerate(flag_eq):
        vo.flag_eq[i] = f
    vo.read_group = read_group.encode() if read_group else None
    keep = None
    if intervals is not None:
        keep = (np.ascontiguousarray(intervals[0], np.int64), np.ascontiguousarray(intervals[1], np.int64))
        vo.iv_beg, vo.iv_end, vo.n_iv = keep[0].ctypes.data, keep[1].ctypes.data, keep[0].size
        vo.iv_exclude = int(exclude_intervals)
        if keep[0].size == 0:                       # samtools -L with no interval on this reference prints nothing
            if not exclude_intervals:
                return None, None
            vo.n_iv = 0
    vo.max_records = max_records
    if key_window is not None:
        vo.key_beg, vo.key_end = int(key_window[0]), int(key_window[1])
        if vo.key_end <= 0:
            return None, None
    return vo, keep


class BamFile:
    def __init__(self, path: str, threads: int = 0):
        h = C.c_void_p()
        check(lib.wgbs_bam_open(path.encode(), threads, C.byref(h)))
        self.h = h.value
        self.refs = [lib.wgbs_bam_ref_name(self.h, i).decode() for i in range(lib.wgbs_bam_nref(self.h))]

    @property
    def header(self) -> str:
        return lib.wgbs_bam_header(self.h).decode(errors="replace")

    def nrecords(self, chrom: str | None = None) -> int:
        return int(lib.wgbs_bam_nrecords(self.h, -1 if chrom is None else self.refs.index(chrom)))

    def view(self, chrom: str | None = None, **kw):
        """(as_array=True: the text as a numpy uint8 array over the library's buffer instead of a bytes copy)
        `samtools view BAM chrom[:beg-end] -q mapq -F exclude_flags [-f include_flags] [-r read_group]`, optionally
        `| awk '$2 == flag_eq[0] || ...'`, `-M -L bed` (intervals = (starts, ends) 0-based half-open, sorted, merged) or
        `bedtools intersect -v` (exclude_intervals), `| head -max_records`."""
        as_array = kw.pop("as_array", False)
        vo, keep = view_opts(self.refs, chrom, **kw)
        if vo is None:
            return np.zeros(0, np.uint8) if as_array else b""
        ptr = C.c_void_p(); n = C.c_size_t(); nr = C.c_uint64()
        check(lib.wgbs_bam_view_ex(self.h, C.byref(vo), C.byref(ptr), C.byref(n), C.byref(nr)))
        if as_array:
            # the library's own buffer as a uint8 array, released when the array is: no copy of gigabytes of text on one thread
            if not n.value:
                lib.wgbs_host_free(ptr)
                return np.zeros(0, np.uint8)
            a = np.ctypeslib.as_array((C.c_uint8 * n.value).from_address(ptr.value))
            weakref.finalize(a, lib.wgbs_host_free, ptr.value)
            return a
        try:
            return C.string_at(ptr, n.value)
        finally:
            lib.wgbs_host_free(ptr)

    def close(self):
        if self.h:
            lib.wgbs_bam_close(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def _part_args(refs, ref_lens):
    names = (C.c_char_p * max(len(refs), 1))(*[r.encode() for r in refs])
    lens = (C.c_int32 * max(len(refs), 1))(*[int(x) for x in (ref_lens if ref_lens is not None else [0] * len(refs))])
    return names, lens


class BamPart(BamFile):
    """A window of a .bam (consecutive whole BGZF blocks) decoded on host threads: viewed like a BamFile.  `tail` = inflated
    offset of the first byte that is not part of a complete record (wgbs_bam_open_part)."""

    def __init__(self, data, refs=None, ref_lens=None, first_record: int = 0, threads: int = 0):
        a = np.frombuffer(data, np.uint8)
        h = C.c_void_p(); tail = C.c_uint64()
        if refs is None:
            check(lib.wgbs_bam_open_part(a.ctypes.data, a.size, 0, None, None, 1, 0, threads, C.byref(h), C.byref(tail)))
        else:
            names, lens = _part_args(refs, ref_lens)
            check(lib.wgbs_bam_open_part(a.ctypes.data, a.size, len(refs), names, lens, 0, int(first_record), threads, C.byref(h), C.byref(tail)))
        self.h, self.tail = h.value, int(tail.value)
        self.refs = [lib.wgbs_bam_ref_name(self.h, i).decode() for i in range(lib.wgbs_bam_nref(self.h))]
        self.ref_lens = list(ref_lens) if ref_lens is not None else None
        self.inflated_bytes = int(lib.wgbs_bam_inflated_bytes(self.h))

    def last_record(self):
        """(refid, 0-based POS) of the last complete record; refid -2: no record"""
        r = C.c_int(); p = C.c_int64()
        check(lib.wgbs_bam_last_record(self.h, C.byref(r), C.byref(p)))
        return r.value, p.value

    def first_key(self, refid: int, key: int, **view_kw):
        """inflated offset of the first record of `refid` that passes the view filters and has template key >= key, or None"""
        vo, keep = view_opts(self.refs, None, **view_kw)
        if vo is None:
            return None
        off = C.c_uint64(); found = C.c_int()
        check(lib.wgbs_bam_first_key(self.h, C.byref(vo), refid, int(key), C.byref(off), C.byref(found)))
        return int(off.value) if found.value else None


def bgzf_block_table(path: str):
    """(coff, csize, usize) int64 arrays of every BGZF block of a file, by hopping over the block headers (18 bytes + the 4-byte
    ISIZE at the end of each block are read; nothing is inflated)"""
    import struct
    coff, csize, usize = [], [], []
    with open(path, "rb") as f:
        f.seek(0, 2); fsz = f.tell(); off = 0
        while off + 28 <= fsz:
            f.seek(off); h = f.read(18)
            if len(h) < 18 or h[:4] != b"\x1f\x8b\x08\x04":
                raise ValueError(f"{path}: not a BGZF file (bad block header at {off})")
            xlen = struct.unpack_from("<H", h, 10)[0]
            if h[12:14] == b"BC" and xlen == 6:
                bs = struct.unpack_from("<H", h, 16)[0] + 1
            else:                                                   # other extra subfields first: walk them
                f.seek(off + 12); x = f.read(xlen); bs = 0; k = 0
                while k + 4 <= xlen:
                    sl = struct.unpack_from("<H", x, k + 2)[0]
                    if x[k:k + 2] == b"BC" and sl == 2:
                        bs = struct.unpack_from("<H", x, k + 4)[0] + 1; break
                    k += 4 + sl
            if bs < 28 or off + bs > fsz:
                raise ValueError(f"{path}: corrupt BGZF block at {off}")
            f.seek(off + bs - 4); us = struct.unpack("<I", f.read(4))[0]
            coff.append(off); csize.append(bs); usize.append(us); off += bs
        if off != fsz:
            raise ValueError(f"{path}: {fsz - off} trailing bytes after the last BGZF block")
    return np.array(coff, np.int64), np.array(csize, np.int64), np.array(usize, np.int64)


def probe_block(f, table, b: int, n_ref: int, span: int = 3):
    """(inflated offset within block b.., refid, pos) of the first record that starts at or after the beginning of BGZF block b
    of the open file f, or None when no record starts in the rest of the file (wgbs_bam_probe on a few blocks; more when a
    record is longer than that)"""
    coff, csize, usize = table
    nb = coff.size
    while b < nb:
        e = min(b + span, nb)
        f.seek(int(coff[b])); data = f.read(int(coff[e - 1] + csize[e - 1] - coff[b]))
        a = np.frombuffer(data, np.uint8)
        off = C.c_uint64(); refid = C.c_int(); pos = C.c_int64(); found = C.c_int()
        check(lib.wgbs_bam_probe(a.ctypes.data, a.size, n_ref, C.byref(off), C.byref(refid), C.byref(pos), C.byref(found)))
        if found.value:
            return int(off.value), refid.value, pos.value
        if e == nb:
            return None
        span *= 4
    return None


def chrom_block_ranges(path: str, table, n_ref: int) -> list[tuple[int, int]]:
    """ranges[c] = (lo, hi): the BGZF blocks [lo, hi) of a coordinate-sorted .bam that hold every record of reference c, found
    without a .bai by binary search with wgbs_bam_probe.  With first[c] = the first block b such that the first record STARTING
    in b.. belongs to a reference >= c (unmapped records count as beyond every reference), the records of c start in blocks
    >= first[c] - 1 and end before the block in which the first record of a later reference starts has ended -- that block, not
    first[c + 1], is the upper end: a record may be longer than a block."""
    coff, csize, usize = table
    nb = coff.size
    uoff = np.concatenate([[0], np.cumsum(usize)])
    first, rec_block = [], []
    with open(path, "rb") as f:
        memo = {}

        def probe(b):
            if b not in memo:
                r = probe_block(f, table, b, n_ref) if b < nb else None
                if r is None:
                    memo[b] = ((1 << 30), nb)
                else:
                    at = int(uoff[b]) + r[0]                                     # inflated offset of that record in the file
                    memo[b] = ((1 << 30) if r[1] < 0 else r[1], int(np.searchsorted(uoff, at, side="right")) - 1)
            return memo[b]
        lo_all = 0
        for c in range(n_ref + 1):
            lo, hi = lo_all, nb                      # first b in [lo, nb] with rid(b) >= c   (rid(nb) = infinity)
            while lo < hi:
                m = (lo + hi) // 2
                if probe(m)[0] >= c:
                    hi = m
                else:
                    lo = m + 1
            first.append(lo); rec_block.append(probe(lo)[1] if lo < nb else nb); lo_all = lo
    return [(max(first[c] - 1, 0), min(rec_block[c + 1] + 1, nb)) for c in range(n_ref)]


def stream_parts(path: str, open_part, view_kw_for, budget: int, blocks: tuple[int, int] | None = None, refs0=None, table=None):
    """Read a coordinate-sorted .bam that does not fit in memory as a sequence of parts and yield
        (part, chrom, key_window, chrom_done)
    such that piling up, for every yielded item, the records of `chrom` in `part` that pass the view filters AND whose template
    key (max(POS, PNEXT), wgbs_view_opts) lies in key_window, gives every template of the file exactly once with all its
    records in one part.  chrom_done: no later item carries this chromosome.  The part is closed after its items are consumed.
        open_part(bytes, refs, ref_lens, first_record) -> part object (BamPart / DeviceBamPart): refs is None for the first part
        view_kw_for(chrom) -> the view filters (keyword arguments of view_opts) the caller applies to that chromosome
        budget: inflated bytes per part (a part is at least one block; it grows when a single pile of deferred templates does not
        fit -- deferred are the templates whose key is not yet below the POS of the part's last record)
    Why it is exact: the file is sorted, so when a part ends inside chromosome c at POS P, every record with POS < P is in this
    part or an earlier one.  Templates with key < P are complete (both mates have POS <= key); the others are deferred, and the
    next part starts at the block holding the first deferred record (or the cut-off record), found by first_key().
    blocks / refs0: restrict the pass to BGZF blocks [blocks[0], blocks[1]) (a chromosome that ENDS inside the range is complete;
    the caller ignores chromosomes it did not ask for)."""
    coff, csize, usize = table if table is not None else bgzf_block_table(path)      # (the caller may have scanned the file already)
    nb = coff.size
    uoff = np.concatenate([[0], np.cumsum(usize)])
    refs = ref_lens = None
    b = 0; first_record = 0; prev_hi: dict[int, int] = {}
    grow = 1; at_start = True
    if blocks is not None and blocks[0] > 0:
        # only blocks [blocks[0], blocks[1]) of the file (one rank's share of a multi-GPU run; refs0 = the file's reference names):
        # the first record that starts in the first block is found by a probe
        refs, ref_lens = list(refs0), [0] * len(refs0)
        with open(path, "rb") as f:
            r = probe_block(f, (coff, csize, usize), blocks[0], len(refs))
        if r is None:
            return
        b = blocks[0]; first_record = r[0]; at_start = False
        while first_record >= usize[b] and b + 1 < nb:                # (the probe may have had to look beyond its first block)
            first_record -= int(usize[b]); b += 1
    if blocks is not None:
        nb = min(nb, blocks[1])
        coff, csize, usize, uoff = coff[:nb], csize[:nb], usize[:nb], uoff[:nb + 1]
    with open(path, "rb") as f:
        while b < nb:
            e = int(np.searchsorted(uoff, uoff[b] + budget * grow, side="right")) - 1
            e = min(max(e, b + 1), nb)
            f.seek(int(coff[b])); data = f.read(int(coff[e - 1] + csize[e - 1] - coff[b]))
            part = open_part(data, None if at_start else refs, ref_lens, first_record)     # the first part carries the BAM header
            del data
            if refs is None:
                refs, ref_lens = list(part.refs), [0] * len(part.refs)
            final = e == nb
            L, P = part.last_record()
            try:
                present = [c for c in range(len(refs)) if part.nrecords(refs[c]) > 0]
                restart = part.tail
                if not final and L >= 0:
                    r = part.first_key(L, P, **view_kw_for(refs[L]))
                    if r is not None:
                        restart = min(restart, r)
                if not final:
                    nblk = int(np.searchsorted(uoff, uoff[b] + restart, side="right")) - 1       # the block holding that offset
                    if nblk <= b and restart < part.inflated_bytes:
                        # the deferred templates reach back into the first block of this part: nothing would be gained -- take more
                        if e == nb:
                            final = True
                        else:
                            grow *= 2
                            continue
                for c in present:
                    done = final or c != L
                    hi = (1 << 40) if done else P
                    lo = prev_hi.pop(c, 0)
                    if not done:
                        prev_hi[c] = hi
                    if hi > lo:
                        yield part, refs[c], (lo, hi), done
            finally:
                part.close()
            if final:
                break
            grow = 1; at_start = False
            first_record = int(uoff[b] + restart - uoff[nblk])
            b = nblk


class DeviceBam:
    """A .bam decoded ON THE GPU (csrc/bamdev.cu): the compressed bytes are uploaded, BGZF blocks inflated by one warp each,
    records located and filtered in HBM.  Same interface and byte-identical views as BamFile; `view_dev` keeps the SAM
    text in device memory for Context.pileup_sam."""

    def __init__(self, ctx, path: str):
        h = C.c_void_p()
        check(lib.wgbs_dbam_open_file(ctx.h, path.encode(), C.byref(h)))
        self.ctx, self.h = ctx, h.value
        self.refs = [lib.wgbs_dbam_ref_name(self.h, i).decode() for i in range(lib.wgbs_dbam_nref(self.h))]

    @classmethod
    def from_bytes(cls, ctx, data) -> "DeviceBam":
        """data: the bytes of a .bam file (bytes / numpy uint8 / anything with the buffer protocol; pinned memory uploads fastest)"""
        a = np.frombuffer(data, np.uint8)
        h = C.c_void_p()
        check(lib.wgbs_dbam_open(ctx.h, a.ctypes.data, a.size, C.byref(h)))
        b = cls.__new__(cls)
        b.ctx, b.h = ctx, h.value
        b.refs = [lib.wgbs_dbam_ref_name(b.h, i).decode() for i in range(lib.wgbs_dbam_nref(b.h))]
        return b

    @property
    def header(self) -> str:
        return lib.wgbs_dbam_header(self.h).decode(errors="replace")

    @property
    def inflated_bytes(self) -> int:
        return int(lib.wgbs_dbam_inflated_bytes(self.h))

    def nrecords(self, chrom: str | None = None) -> int:
        return int(lib.wgbs_dbam_nrecords(self.h, -1 if chrom is None else self.refs.index(chrom)))

    def view_dev(self, chrom: str | None = None, **kw):
        """SAM text in device memory (a DevBuf; len() == number of bytes), or None when nothing can pass"""
        from .api import DevBuf
        vo, keep = view_opts(self.refs, chrom, **kw)
        if vo is None:
            return None
        ptr = C.c_void_p(); n = C.c_size_t(); nr = C.c_uint64()
        check(lib.wgbs_dbam_view(self.ctx.h, self.h, C.byref(vo), C.byref(ptr), C.byref(n), C.byref(nr)))
        return DevBuf.adopt(self.ctx, ptr.value, n.value)

    def pileup(self, index, chrom: str | None = None, view=None, min_cpg: int = 1, clip: int = 0, paired: int = -1, nanopore: bool = False,
               np_thresh: float = 0.67, cpc_call: str = "C", combine_mods: bool = False, mbias: bool = False, keep_names: bool = False,
               ctx=None):
        """(ctx: the Context to run on -- another stream of the same GPU may pile up from this file's resident records; default: the
        one that opened the file)
        `samtools view BAM chrom ... | [match_maker |] patter ...` without leaving the device (wgbs_pileup_dbam): the records that
        pass `view` (keyword arguments of view_opts) go to the pileup kernels.  Returns (Pats, stats) like Context.pileup_sam;
        (None, stats with lines == 0) when nothing can pass."""
        from ._lib import PileupOpts
        from .api import Pats
        keys = ["lines", "pairs", "empty", "short", "invalid", "paired", "nanopore", "templates"]
        vo, keep = view_opts(self.refs, chrom, **(view or {}))
        if vo is None:
            return None, dict(zip(keys, [0] * 8))
        o = PileupOpts(min_cpg, clip, paired, int(nanopore), int(combine_mods), np_thresh, cpc_call.encode(), int(keep_names))
        h = C.c_void_p(); st = (C.c_uint64 * 8)()
        mb = np.zeros((2, 2, 1000, 2), np.int32) if mbias else None
        cx = ctx or self.ctx
        check(lib.wgbs_pileup_dbam(cx.h, index.h, self.h, C.byref(vo), C.addressof(o), C.byref(h), C.addressof(st), mb.ctypes.data if mbias else None))
        stats = dict(zip(keys, [int(x) for x in st]))
        if mbias:
            stats["mbias"] = mb
        return Pats(cx, h.value), stats

    def view(self, chrom: str | None = None, **kw) -> bytes:
        d = self.view_dev(chrom, **kw)
        if d is None:
            return b""
        try:
            return d.to_host().tobytes() if d.nbytes else b""
        finally:
            d.free()

    def close(self):
        if self.h:
            lib.wgbs_dbam_close(self.ctx.h, self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


class DeviceBamPart(DeviceBam):
    """A window of a .bam decoded on the GPU (wgbs_dbam_open_part): the part's bytes are uploaded, inflated and indexed in HBM;
    used like a DeviceBam (view / view_dev / pileup), plus what bamio.stream_parts asks of a part"""

    def __init__(self, ctx, data, refs=None, ref_lens=None, first_record: int = 0):
        a = np.frombuffer(data, np.uint8)
        h = C.c_void_p(); tail = C.c_uint64()
        if refs is None:
            check(lib.wgbs_dbam_open_part(ctx.h, a.ctypes.data, a.size, 0, None, None, 1, 0, C.byref(h), C.byref(tail)))
        else:
            names, lens = _part_args(refs, ref_lens)
            check(lib.wgbs_dbam_open_part(ctx.h, a.ctypes.data, a.size, len(refs), names, lens, 0, int(first_record), C.byref(h), C.byref(tail)))
        self.ctx, self.h, self.tail = ctx, h.value, int(tail.value)
        self.refs = [lib.wgbs_dbam_ref_name(self.h, i).decode() for i in range(lib.wgbs_dbam_nref(self.h))]

    def last_record(self):
        r = C.c_int(); p = C.c_int64()
        check(lib.wgbs_dbam_last_record(self.ctx.h, self.h, C.byref(r), C.byref(p)))
        return r.value, p.value

    def first_key(self, refid: int, key: int, **view_kw):
        vo, keep = view_opts(self.refs, None, **view_kw)
        if vo is None:
            return None
        off = C.c_uint64(); found = C.c_int()
        check(lib.wgbs_dbam_first_key(self.ctx.h, self.h, C.byref(vo), refid, int(key), C.byref(off), C.byref(found)))
        return int(off.value) if found.value else None


# ----------------------------------------------------------------------------------------------------------------------
# writer (test inputs): SAM text -> BAM bytes
# ----------------------------------------------------------------------------------------------------------------------
_SEQ = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}
_CIG = {c: i for i, c in enumerate("MIDNSHP=X")}


def _tag_bytes(tag: bytes) -> bytes:
    name, ty, val = tag.split(b":", 2)
    if ty == b"A":
        return name + b"A" + val[:1]
    if ty == b"i":
        v = int(val)
        for code, fmt, lo, hi in ((b"c", "<b", -128, 127), (b"C", "<B", 0, 255), (b"s", "<h", -32768, 32767), (b"S", "<H", 0, 65535), (b"i", "<i", -2**31, 2**31 - 1), (b"I", "<I", 0, 2**32 - 1)):
            if lo <= v <= hi:
                return name + code + struct.pack(fmt, v)
    if ty == b"f":
        return name + b"f" + struct.pack("<f", float(val))
    if ty in (b"Z", b"H"):
        return name + ty + val + b"\0"
    if ty == b"B":
        sub = val[:1]; items = [x for x in val[2:].split(b",") if x] if len(val) > 1 else []
        fmt = {b"c": "b", b"C": "B", b"s": "h", b"S": "H", b"i": "i", b"I": "I", b"f": "f"}[sub]
        conv = float if sub == b"f" else int
        return name + b"B" + sub + struct.pack("<I", len(items)) + struct.pack("<%d%s" % (len(items), fmt), *[conv(x) for x in items])
    raise ValueError(tag)


def _records(sam: bytes, rid: dict) -> bytes:
    """BAM records (SAM spec 4.2) of the alignment lines of `sam`"""
    out = []
    for line in sam.splitlines():
        if not line or line.startswith(b"@"):
            continue
        t = line.split(b"\t")
        name, flag, rname, pos, mapq, cigar, rnext, pnext, tlen, seq, qual = t[:11]
        ops = re.findall(rb"(\d+)([MIDNSHP=X])", cigar) if cigar != b"*" else []
        cig = b"".join(struct.pack("<I", int(n) << 4 | _CIG[o.decode()]) for n, o in ops)
        l_seq = 0 if seq == b"*" else len(seq)
        if l_seq:
            codes = _SEQ_LUT[np.frombuffer(seq, np.uint8)]
            if l_seq & 1:
                codes = np.append(codes, np.uint8(0))
            sq = ((codes[0::2] << 4) | codes[1::2]).astype(np.uint8).tobytes()
        else:
            sq = b""
        ql = (b"\xff" * l_seq) if qual == b"*" else (np.frombuffer(qual, np.uint8) - 33).astype(np.uint8).tobytes()
        r = rid.get(rname.decode(), -1)
        nr = r if rnext == b"=" else rid.get(rnext.decode(), -1)
        end = int(pos) - 1 + sum(int(n) for n, o in ops if o in b"MDN=X") if ops else int(pos)
        bin_ = _reg2bin(int(pos) - 1, max(end, int(pos)))
        core = struct.pack("<iiBBHHHiiii", r, int(pos) - 1, len(name) + 1, int(mapq), bin_, len(ops), int(flag), l_seq, nr, int(pnext) - 1, int(tlen))
        body = core + name + b"\0" + cig + sq + ql + b"".join(_tag_bytes(x) for x in t[11:])
        out.append(struct.pack("<i", len(body)) + body)
    return b"".join(out)


_SEQ_LUT = np.zeros(256, np.uint8)
for _c, _i in _SEQ.items():
    _SEQ_LUT[ord(_c)] = _i


def sam_to_bam(sam: bytes, refs: list[tuple[str, int]], header_text: str | None = None, procs: int = 1) -> bytes:
    """SAM records (no header lines needed) + reference list -> BGZF-compressed BAM bytes.  procs > 1: the records are
    built by that many forked worker processes (call it before CUDA is initialised in this process)."""
    if header_text is None:
        header_text = "@HD\tVN:1.6\tSO:coordinate\n" + "".join(f"@SQ\tSN:{n}\tLN:{l}\n" for n, l in refs)
    rid = {n: i for i, (n, _) in enumerate(refs)}
    out = [b"BAM\1", struct.pack("<i", len(header_text)), header_text.encode(), struct.pack("<i", len(refs))]
    for n, l in refs:
        out += [struct.pack("<i", len(n) + 1), n.encode() + b"\0", struct.pack("<i", l)]
    if procs > 1 and len(sam) > (1 << 20):
        import multiprocessing as mp
        cuts = [0]
        for k in range(1, procs * 4):
            c = sam.find(b"\n", len(sam) * k // (procs * 4)) + 1
            if c > cuts[-1]:
                cuts.append(c)
        cuts.append(len(sam))
        with mp.get_context("fork").Pool(procs) as pool:
            out += pool.starmap(_records, [(sam[a:b], rid) for a, b in zip(cuts[:-1], cuts[1:])])
    else:
        out.append(_records(sam, rid))
    return bgzf_compress(b"".join(out), threads=max(8, procs))


def _reg2bin(beg: int, end: int) -> int:
    end -= 1
    for shift, base in ((14, 4681), (17, 585), (20, 73), (23, 9), (26, 1)):
        if beg >> shift == end >> shift:
            return base + (beg >> shift)
    return 0
