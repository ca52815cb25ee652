"""Python face of the C ABI: a Context per GPU, device-resident pat records, and the per-step operators.

Buffers cross the boundary as raw addresses: numpy arrays / bytes for host data, ``DevBuf`` (or any object with a
``data_ptr()`` such as a torch CUDA tensor) for device data."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import PileupOpts, WgbsError, check, lib


def _addr(x):
    """address of a host numpy array / bytes-like, or of a device buffer (DevBuf / torch tensor)."""
    if x is None:
        return None
    if isinstance(x, DevBuf):
        return x.ptr
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    if isinstance(x, np.ndarray):
        if not x.flags["C_CONTIGUOUS"]:
            raise ValueError("array must be C-contiguous")
        return x.ctypes.data
    if isinstance(x, bytes):
        return C.cast(C.c_char_p(x), C.c_void_p).value
    if isinstance(x, (bytearray, memoryview)):
        # the address of the caller's own buffer (never of a temporary copy): the caller keeps `x` alive for the call,
        # and writes through the pointer land in `x`
        mv = memoryview(x)
        if not mv.contiguous:
            raise ValueError("buffer must be contiguous")
        if mv.nbytes == 0:
            return None
        return np.frombuffer(mv, np.uint8).ctypes.data
    if isinstance(x, int):
        return x
    raise TypeError(type(x))


class DevBuf:
    """A device allocation owned by a Context (stream-ordered)."""

    def __init__(self, ctx: "Context", nbytes: int):
        self.ctx, self.nbytes = ctx, int(nbytes)
        p = C.c_void_p()
        check(lib.wgbs_dev_alloc(ctx.h, self.nbytes, C.byref(p)))
        self.ptr = p.value

    @classmethod
    def adopt(cls, ctx: "Context", ptr: int, nbytes: int) -> "DevBuf":
        """take ownership of a device buffer the library allocated (released with wgbs_dev_free)"""
        b = cls.__new__(cls)
        b.ctx, b.nbytes, b.ptr = ctx, int(nbytes), ptr
        return b

    def __len__(self):
        return self.nbytes

    def free(self):
        if self.ptr:
            lib.wgbs_dev_free(self.ctx.h, self.ptr)
            self.ptr = None

    def to_host(self, dtype=np.uint8) -> np.ndarray:
        out = np.empty(self.nbytes // np.dtype(dtype).itemsize, dtype)
        check(lib.wgbs_memcpy(self.ctx.h, out.ctypes.data, self.ptr, out.nbytes))
        return out


class DevView:
    """A byte range inside somebody else's device buffer (no ownership): what patio.pat_pieces yields for device-resident text"""

    def __init__(self, ptr: int, nbytes: int):
        self.ptr, self.nbytes = int(ptr), int(nbytes)

    def __len__(self):
        return self.nbytes


class Pats:
    """Device-resident pat records (idx, len, count, 2-bit symbol pool)."""

    def __init__(self, ctx: "Context", handle):
        self.ctx, self.h = ctx, handle

    def __len__(self):
        n = C.c_uint64(); w = C.c_uint64()
        check(lib.wgbs_pats_count(self.h, C.byref(n), C.byref(w)))
        return n.value

    @property
    def pool_words(self) -> int:
        n = C.c_uint64(); w = C.c_uint64()
        check(lib.wgbs_pats_count(self.h, C.byref(n), C.byref(w)))
        return w.value

    def download(self):
        n, w = len(self), self.pool_words
        idx = np.empty(n, np.uint32); ln = np.empty(n, np.uint32); cnt = np.empty(n, np.uint32)
        off = np.empty(max(n, 1), np.uint32); pool = np.empty(max(w, 1), np.uint32)
        check(lib.wgbs_pats_download(self.ctx.h, self.h, idx.ctypes.data, ln.ctypes.data, cnt.ctypes.data, off.ctypes.data, pool.ctypes.data))
        return idx.view(np.int32), ln, cnt.view(np.int32), off[:n], pool[:w]

    def patterns(self) -> list[bytes]:
        """decode the symbol pool back to ASCII patterns (host-side, for tests and small outputs)."""
        idx, ln, cnt, off, pool = self.download()
        lut = np.frombuffer(b".CHT", np.uint8)
        out = []
        for i in range(idx.size):
            L = int(ln[i]); w = pool[off[i]: off[i] + (L + 15) // 16]
            sh = (30 - 2 * np.arange(16, dtype=np.uint32))[None, :]
            sym = ((w[:, None] >> sh) & 3).reshape(-1)[:L]
            out.append(lut[sym].tobytes())
        return out

    def collapse(self, long: bool = False, mode: int | None = None) -> "Pats":
        """sort -k2,2n -k3,3 | uniq -c (in place); long=True: only sort (by idx, pattern, read name), no uniq (--long).
        mode: WGBS_COLLAPSE_* (2 = patterns may end in '.', 3 = merge adjacent identical records only, no sort)."""
        if mode is not None:
            check(lib.wgbs_collapse_ex(self.ctx.h, self.h, int(mode)))
        else:
            check((lib.wgbs_collapse_long if long else lib.wgbs_collapse)(self.ctx.h, self.h))
        return self

    def cview(self, bstart, bend, strict: bool = False, strip: bool = False, no_gaps: bool = False, min_cpgs: int = 1, pre=None) -> "Pats":
        """`cview --blocks_path/--sites ... [--strict] [--strip] [--no_gaps] [--min_cpgs N]` on these records (file order).
        bstart/bend: blocks [startCpG, endCpG) sorted by start.  pre: optional (lo[], hi[]) closed ranges of start indices
        (the `tabix` pre-selection).  Returns new Pats, not yet sorted / collapsed."""
        bs = np.ascontiguousarray(bstart, np.int32); be = np.ascontiguousarray(bend, np.int32)
        pl = ph = None; npre = 0
        if pre is not None:
            pl = np.ascontiguousarray(pre[0], np.int32); ph = np.ascontiguousarray(pre[1], np.int32); npre = pl.size
        h = C.c_void_p()
        check(lib.wgbs_cview(self.ctx.h, self.h, bs.ctypes.data, be.ctypes.data, bs.size, pl.ctypes.data if npre else None,
                             ph.ctypes.data if npre else None, npre, int(strict), int(strip), int(no_gaps), int(min_cpgs), C.byref(h)))
        return Pats(self.ctx, h.value)

    def to_text(self, chrom: str, out=None, long: bool = False):
        """pat text.  out: optional preallocated host uint8 array or DevBuf (then the number of bytes is returned).
        long=True: `chr idx pattern 1 qname` lines (--long)."""
        fmt = lib.wgbs_pats_format_long if long else lib.wgbs_pats_format
        n = C.c_size_t()
        if out is None:
            check(fmt(self.ctx.h, self.h, chrom.encode(), None, 0, C.byref(n)))
            buf = np.empty(max(n.value, 1), np.uint8)
            check(fmt(self.ctx.h, self.h, chrom.encode(), buf.ctypes.data, n.value, C.byref(n)))
            return buf[:n.value].tobytes()
        cap = out.nbytes
        check(fmt(self.ctx.h, self.h, chrom.encode(), _addr(out), cap, C.byref(n)))
        return n.value

    def free(self):
        if self.h:
            lib.wgbs_pats_free(self.ctx.h, self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Index:
    """Device-resident CpG dictionary of one chromosome / region."""

    def __init__(self, ctx: "Context", loci, first_idx: int = 1):
        a = np.ascontiguousarray(loci, np.uint32)
        h = C.c_void_p()
        check(lib.wgbs_index_load(ctx.h, a.ctypes.data, a.size, int(first_idx), C.byref(h)))
        self.ctx, self.h, self.n, self.first_idx = ctx, h.value, a.size, int(first_idx)

    def free(self):
        if self.h:
            lib.wgbs_index_free(self.ctx.h, self.h)
            self.h = None


class Context:
    """One per GPU.  ``stream``: a raw cudaStream_t (int) to run on, e.g. torch.cuda.current_stream().cuda_stream."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self.h = lib.wgbs_create(int(device), C.c_void_p(stream) if stream else None)
        if not self.h:
            raise WgbsError(lib.wgbs_last_error().decode(errors="replace"))
        self.device = device

    def close(self):
        if self.h:
            lib.wgbs_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def sync(self):
        check(lib.wgbs_sync(self.h))

    @property
    def launches(self) -> int:
        return int(lib.wgbs_launch_count(self.h))

    def prof(self, on: bool = True):
        check(lib.wgbs_prof_enable(self.h, int(on)))

    def prof_report(self) -> dict:
        """{kernel: (launches, total_ms)} since prof(True)"""
        buf = C.create_string_buffer(1 << 16)
        n = lib.wgbs_prof_report(self.h, buf, len(buf))
        check(n)
        out = {}
        for line in buf.raw[:n].decode().splitlines():
            k, c, ms = line.split("\t")
            out[k] = (int(c), float(ms))
        return out

    # ---- memory -------------------------------------------------------------------------------------------------
    def alloc(self, nbytes: int) -> DevBuf:
        return DevBuf(self, nbytes)

    def upload(self, data) -> DevBuf:
        a = np.frombuffer(data, np.uint8) if isinstance(data, (bytes, bytearray, memoryview)) else np.ascontiguousarray(data)
        b = DevBuf(self, a.nbytes)
        check(lib.wgbs_memcpy(self.h, b.ptr, a.ctypes.data, a.nbytes))
        return b

    def bgzf_inflate(self, data) -> DevBuf:
        """the bytes of a BGZF file (host) -> its inflated bytes in device memory (one warp per BGZF block, its lanes on spans of the bit stream; deflate, ISIZE and
        CRC32 are verified)"""
        a = np.frombuffer(data, np.uint8)
        p = C.c_void_p(); n = C.c_size_t()
        check(lib.wgbs_bgzf_inflate(self.h, a.ctypes.data, a.size, C.byref(p), C.byref(n)))
        return DevBuf.adopt(self, p.value, n.value)

    # ---- pileup -------------------------------------------------------------------------------------------------
    def load_index(self, loci, first_idx: int = 1) -> "Index":
        return Index(self, loci, first_idx)

    def pileup_sam(self, index: "Index", sam, min_cpg: int = 1, clip: int = 0, paired: int = -1, nanopore: bool = False,
                   np_thresh: float = 0.67, cpc_call: str = "C", combine_mods: bool = False, nbytes: int | None = None,
                   mbias: bool = False, keep_names: bool = False):
        """SAM text (bytes or DevBuf) -> (Pats of templates, stats dict).  The reference's
        `samtools view ... | [match_maker |] patter DICT REGION ...` for one chromosome."""
        if isinstance(sam, DevBuf):
            p, n = sam.ptr, sam.nbytes if nbytes is None else nbytes
        else:
            a = np.frombuffer(sam, np.uint8)
            p, n = a.ctypes.data, a.size
        o = PileupOpts(min_cpg, clip, paired, int(nanopore), int(combine_mods), np_thresh, cpc_call.encode(), int(keep_names))
        h = C.c_void_p(); st = (C.c_uint64 * 8)()
        mb = np.zeros((2, 2, 1000, 2), np.int32) if mbias else None
        check(lib.wgbs_pileup_sam_mbias(self.h, index.h, p, n, C.addressof(o), C.byref(h), C.addressof(st), mb.ctypes.data if mbias else None))
        keys = ["lines", "pairs", "empty", "short", "invalid", "paired", "nanopore", "templates"]
        stats = dict(zip(keys, [int(x) for x in st]))
        if mbias:
            stats["mbias"] = mb          # [OT|OB][mate][position][meth|unmeth]
        return Pats(self, h.value), stats

    def sort_pairs(self, keys: np.ndarray, vals: np.ndarray):
        k = self.upload(np.ascontiguousarray(keys, np.uint32)); v = self.upload(np.ascontiguousarray(vals, np.uint32))
        check(lib.wgbs_sort_pairs_u32(self.h, k.ptr, v.ptr, keys.size))
        ko, vo = k.to_host(np.uint32), v.to_host(np.uint32)
        k.free(); v.free()
        return ko, vo

    # ---- segment ------------------------------------------------------------------------------------------------
    def segment(self, betas, dists, chunks, max_cpg: int, max_bp: int, pseudo: float):
        """betas: list of K uint8[nsites,2] arrays (numpy, DevBuf or torch CUDA tensors) over one site range;
        dists: uint32[nsites]; chunks: iterable of (start, n) relative to the arrays.
        Returns a list of int64 border arrays (relative to each chunk's start), like `segmentor`'s stdout."""
        K = len(betas)
        keep = [np.ascontiguousarray(b, np.uint8) if isinstance(b, np.ndarray) else b for b in betas]
        ptrs = (C.c_void_p * K)(*[_addr(b) for b in keep])
        d = np.ascontiguousarray(dists, np.uint32) if isinstance(dists, np.ndarray) else dists
        nsites = d.size if isinstance(d, np.ndarray) else d.nbytes // 4
        ch = np.ascontiguousarray(np.asarray(list(chunks), dtype=np.uint32).reshape(-1, 2))
        tot = int((ch[:, 1].astype(np.int64) + 1).sum())
        borders = np.empty(max(tot, 1), np.int32); nb = np.empty(max(ch.shape[0], 1), np.int32)
        check(lib.wgbs_segment(self.h, C.addressof(ptrs), K, _addr(d), nsites, ch.ctypes.data, ch.shape[0], int(max_cpg), int(max_bp),
                               float(pseudo), borders.ctypes.data, nb.ctypes.data))
        out = []; o = 0
        for c in range(ch.shape[0]):
            out.append(borders[o:o + nb[c]].astype(np.int64)); o += int(ch[c, 1]) + 1
        return out

    def glibc_log2_probe(self, p: np.ndarray):
        p = np.ascontiguousarray(p, np.float32)
        a = np.empty(p.size, np.float32); b = np.empty(p.size, np.float64)
        check(lib.wgbs_glibc_log2_probe(self.h, p.ctypes.data, p.size, a.ctypes.data, b.ctypes.data))
        return a, b

    # ---- pat ----------------------------------------------------------------------------------------------------
    def pats_from_text(self, text, nbytes: int | None = None) -> Pats:
        """text: bytes (host), or DevBuf / DevView (device-resident pat text)."""
        if isinstance(text, (DevBuf, DevView)):
            p, n = text.ptr, text.nbytes if nbytes is None else nbytes
        else:
            a = np.frombuffer(text, np.uint8)
            p, n = a.ctypes.data, a.size
        h = C.c_void_p()
        check(lib.wgbs_pats_from_text(self.h, p, n, C.byref(h)))
        return Pats(self, h.value)

    def pat2beta(self, pats: Pats, start: int, end: int, meth_cov=None, zero_first: bool = True):
        """device int32[end-start,2] (meth, cover) counts; returns the DevBuf (allocated if not given)."""
        n = end - start
        buf = meth_cov if meth_cov is not None else DevBuf(self, max(n, 1) * 8)
        check(lib.wgbs_pat2beta(self.h, pats.h, start, end, _addr(buf), int(zero_first)))
        return buf

    def trim(self, meth_cov, n: int, nbits: int = 8) -> np.ndarray:
        out = np.empty((n, 2), np.uint8 if nbits == 8 else np.uint16)
        check(lib.wgbs_trim(self.h, _addr(meth_cov), n, nbits, out.ctypes.data))
        return out

    def pat2beta_text(self, text: bytes, start: int, end: int, nbits: int = 8, want_counts: bool = False):
        """pat text (host) -> beta array (host), the whole of reference pat2beta.py:32-37 in one call."""
        n = end - start
        a = np.frombuffer(text, np.uint8)
        out = np.empty((n, 2), np.uint8 if nbits == 8 else np.uint16)
        mc = np.empty((n, 2), np.int32) if want_counts else None
        check(lib.wgbs_pat2beta_text(self.h, a.ctypes.data, a.size, start, end, nbits, out.ctypes.data,
                                     mc.ctypes.data if mc is not None else None))
        return (out, mc) if want_counts else out

    def beta_to_blocks(self, beta: np.ndarray, bstart, bend, out_bits: int | None = 8, want_sums: bool = False):
        """per-block (meth, cover) sums of a beta array (uint8[N,2] or uint16[N,2]); returns the trimmed `.bin`/`.lbeta`
        rows (out_bits 8 / 16), the raw int64 sums (want_sums), or both (reference beta_to_blocks.py:101-150)."""
        b = np.ascontiguousarray(beta)
        if b.dtype not in (np.uint8, np.uint16) or b.ndim != 2 or b.shape[1] != 2:
            raise ValueError("beta must be uint8[N,2] or uint16[N,2]")
        bs = np.ascontiguousarray(bstart, np.int32); be = np.ascontiguousarray(bend, np.int32)
        out = np.zeros((bs.size, 2), np.uint8 if out_bits == 8 else np.uint16) if out_bits else None
        sums = np.zeros((bs.size, 2), np.int64) if want_sums else None
        check(lib.wgbs_beta_to_blocks(self.h, b.ctypes.data, 8 * b.dtype.itemsize, b.shape[0], bs.ctypes.data, be.ctypes.data, bs.size,
                                      int(out_bits or 0), out.ctypes.data if out is not None else None,
                                      sums.ctypes.data if sums is not None else None))
        return (out, sums) if (out is not None and want_sums) else (sums if want_sums else out)

    def homog(self, pats: Pats, blocks: np.ndarray, rng, min_cpgs: int, inclusive: bool = False) -> np.ndarray:
        bs = np.ascontiguousarray(blocks[:, 0], np.int32); be = np.ascontiguousarray(blocks[:, 1], np.int32)
        r = np.ascontiguousarray(rng, np.float32); nb = r.size - 1
        out = np.empty((bs.size, nb), np.int32)
        check(lib.wgbs_homog(self.h, pats.h, bs.ctypes.data, be.ctypes.data, bs.size, r.ctypes.data, nb, int(min_cpgs),
                             int(inclusive), out.ctypes.data))
        return out
