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


def splitextgz(name: str) -> str:
    for suf in (".pat.gz", ".pat"):
        if name.endswith(suf):
            return name[: -len(suf)]
    return name
