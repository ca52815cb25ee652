"""pat / beta file I/O: .pat[.gz] text in, BGZF (.pat.gz) out, .beta out.

BGZF follows the SAM spec section 4.1 the way htslib's bgzip writes it: <= 0xff00 input bytes per block, raw deflate at
level 6, 'BC' extra field with the block size, and the 28-byte EOF block at the end of every file -- so the `cat` of
per-chromosome parts (reference bam2pat.py:408) carries one EOF block per part.  Compressed BYTES depend on the zlib
build and are not pinned; the decompressed bytes are."""
from __future__ import annotations

import gzip
import struct
import zlib
from concurrent.futures import ThreadPoolExecutor

BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
BGZF_BLOCK = 0xFF00


def _bgzf_block(data: bytes, level: int = 6) -> bytes:
    co = zlib.compressobj(level, zlib.DEFLATED, -15)
    comp = co.compress(data) + co.flush()
    bsize = len(comp) + 25
    return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", bsize)
            + comp + struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data)))


def bgzf_compress(data: bytes, threads: int = 8, level: int = 6) -> bytes:
    chunks = [data[i:i + BGZF_BLOCK] for i in range(0, len(data), BGZF_BLOCK)]
    if threads > 1 and len(chunks) > 4:
        with ThreadPoolExecutor(threads) as ex:                # zlib releases the GIL
            blocks = list(ex.map(lambda c: _bgzf_block(c, level), chunks))
    else:
        blocks = [_bgzf_block(c, level) for c in chunks]
    return b"".join(blocks) + BGZF_EOF


def read_pat_text(path: str) -> bytes:
    """`gunzip -cd X.pat.gz` / `cat X.pat` (reference pat2beta.py:17-22)"""
    if path.endswith(".pat.gz"):
        with gzip.open(path, "rb") as f:
            return f.read()
    if path.endswith(".pat"):
        with open(path, "rb") as f:
            return f.read()
    raise ValueError(f"Invalid pat suffix: {path}")


def is_bgzf(head: bytes) -> bool:
    """gzip member with the 'BC' extra subfield first (what bgzip writes; SAM spec 4.1)"""
    return len(head) >= 18 and head[:4] == b"\x1f\x8b\x08\x04" and head[12:14] == b"BC"


def read_pat_device(ctx, path: str):
    """the pat text of a bgzip-compressed X.pat.gz as a device buffer (DevBuf): only the compressed bytes cross PCIe, the BGZF
    blocks are inflated in HBM (csrc/bamdev.cu).  Returns None when the file is not BGZF (plain gzip / uncompressed)."""
    if not path.endswith(".pat.gz"):
        return None
    with open(path, "rb") as f:
        raw = f.read()
    if not is_bgzf(raw[:18]):
        return None
    return ctx.bgzf_inflate(raw)


def read_pat_device_shard(ctx, path: str, rank: int, world: int):
    """One rank's share of a bgzip-compressed X.pat.gz under torchrun, inflated in HBM: the rank reads only ITS run of BGZF blocks
    (+ the block before and the block after), and owns the lines that START inside its run -- the same rule on both sides of every
    boundary, so the ranks' shares partition the records.  Returns (DevBuf to free, DevView of the owned lines), or None when the
    file is not BGZF (the caller then shards the host text, dist.shard_lines).  pat2beta and homog are sums over records
    (reference pat2beta.py / homog.cpp:84 shard by chromosome; any record split gives the same sums)."""
    from .api import DevView
    from .bamio import bgzf_block_table
    if not path.endswith(".pat.gz"):
        return None
    with open(path, "rb") as f:
        if not is_bgzf(f.read(18)):
            return None
    coff, csize, usize = bgzf_block_table(path)
    nb = int(coff.size)
    b0, b1 = nb * rank // world, nb * (rank + 1) // world
    lo, hi = max(b0 - 1, 0), min(b1 + 1, nb)                      # neighbours: where the first / last owned line starts and ends
    if hi <= lo:
        return None
    with open(path, "rb") as f:
        f.seek(int(coff[lo])); raw = f.read(int(coff[hi - 1] + csize[hi - 1] - coff[lo]))
    buf = ctx.bgzf_inflate(raw + BGZF_EOF)
    S = int(usize[lo:b0].sum()); E = S + int(usize[b0:b1].sum()); n = len(buf)

    def line_start_at_or_after(pos: int) -> int:
        """first offset >= pos where a line starts (offset 0, or the byte behind a newline)"""
        if pos <= 0:
            return 0
        if pos >= n:
            return n
        a = max(pos - 1, 0); w = min(n - a, 1 << 20)
        import numpy as np
        from ._lib import check, lib
        win = np.empty(w, np.uint8)
        check(lib.wgbs_memcpy(ctx.h, win.ctypes.data, buf.ptr + a, w))
        k = win.tobytes().find(b"\n")
        if k < 0:
            if a + w >= n:
                return n
            raise ValueError("pat line longer than 1 MiB")
        return a + k + 1
    start = line_start_at_or_after(S) if b0 > 0 else 0
    end = line_start_at_or_after(E) if b1 < nb else n
    return buf, DevView(buf.ptr + start, max(end - start, 0))


def pat_pieces(ctx, text, limit: int | None = None):
    """A pat text (bytes, or a DevBuf of device-resident text) as pieces of at most `limit` bytes that end at line ends: one
    wgbs_pats_from_text call takes < 4 GiB, a 30x pat file is more.  pat2beta and homog are sums over records, so the pieces are
    simply processed one after the other.  Yields the text itself when it fits (the common case costs nothing)."""
    import os
    from .api import DevBuf, DevView
    limit = limit or int(os.environ.get("WGBS_PAT_CHUNK_BYTES", 2 << 30))
    n = len(text)
    if n <= limit:
        yield text
        return
    dev = isinstance(text, (DevBuf, DevView))
    mv = None if dev else memoryview(text)
    lo = 0
    while lo < n:
        hi = min(lo + limit, n)
        if hi < n:                                                   # back up to the last line end inside [lo, hi)
            if dev:
                import numpy as np
                from ._lib import check, lib
                w = min(hi - lo, 1 << 20)
                tail = np.empty(w, np.uint8)
                check(lib.wgbs_memcpy(ctx.h, tail.ctypes.data, text.ptr + hi - w, w))
                k = tail.tobytes().rfind(b"\n")
                cut = hi - w + k + 1 if k >= 0 else lo
            else:
                cut = text.rfind(b"\n", lo, hi) + 1
            if cut <= lo:
                raise ValueError("pat line longer than the chunk size (WGBS_PAT_CHUNK_BYTES)")
            hi = cut
        yield DevView(text.ptr + lo, hi - lo) if dev else mv[lo:hi]
        lo = hi


def splitextgz(name: str) -> str:
    for suf in (".pat.gz", ".pat"):
        if name.endswith(suf):
            return name[: -len(suf)]
    return name
