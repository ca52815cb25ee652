"""BAM files: reader (native: csrc/bam.cu, BGZF inflate on host threads) and a small writer used to build test inputs.

`BamFile.view(chrom, ...)` returns the SAM text `samtools view BAM chrom -q Q -F X [-f Y]` prints (reference
bam2pat.py:165) -- the input of the pileup front end."""
from __future__ import annotations

import ctypes as C
import re
import struct

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

    def view(self, chrom: str | None = None, **kw) -> bytes:
        """`samtools view BAM chrom[:beg-end] -q mapq -F exclude_flags [-f include_flags] [-r read_group]`, optionally
        `| awk '$2 == flag_eq[0] || ...'`, `-M -L bed` (intervals = (starts, ends) 0-based half-open, sorted, merged) or
        `bedtools intersect -v` (exclude_intervals), `| head -max_records`."""
        vo, keep = view_opts(self.refs, chrom, **kw)
        if vo is None:
            return b""
        ptr = C.c_void_p(); n = C.c_size_t(); nr = C.c_uint64()
        check(lib.wgbs_bam_view_ex(self.h, C.byref(vo), C.byref(ptr), C.byref(n), C.byref(nr)))
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
               np_thresh: float = 0.67, cpc_call: str = "C", combine_mods: bool = False, mbias: bool = False, keep_names: bool = False):
        """`samtools view BAM chrom ... | [match_maker |] patter ...` without leaving the device (wgbs_pileup_dbam): the records that
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
        check(lib.wgbs_pileup_dbam(self.ctx.h, index.h, self.h, C.byref(vo), C.addressof(o), C.byref(h), C.addressof(st), mb.ctypes.data if mbias else None))
        stats = dict(zip(keys, [int(x) for x in st]))
        if mbias:
            stats["mbias"] = mb
        return Pats(self.ctx, h.value), stats

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
