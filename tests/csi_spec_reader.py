"""TEST INFRASTRUCTURE: an independent reader of BGZF + CSI files, written from the specifications only (SAM spec section 4.1
"The BGZF compression format", section 5.3 "C source code for computing bin number and overlapping bins" / CSIv1 layout, and the tabix
aux block of the tabix paper / tbx.c documentation) -- it shares no code with wgbs_tools_b200/csi.py.  It does what `tabix FILE
chr:beg-end` does with a .csi: validate the file structurally, compute the candidate bins of a query, collect and merge their chunks,
seek by virtual offset, inflate block by block with zlib, and filter the lines.  htslib itself is not available in this image; this
is the closest stand-in for "can tabix read what we wrote"."""
import gzip
import struct
import zlib


def _bgzf_block_at(raw: bytes, coff: int):
    """(block size, inflated bytes) of the BGZF block at compressed offset coff; every field the spec fixes is checked"""
    assert raw[coff:coff + 4] == b"\x1f\x8b\x08\x04", "gzip magic / CM / FLG.FEXTRA"
    xlen = struct.unpack_from("<H", raw, coff + 10)[0]
    p = coff + 12; bsize = None
    while p < coff + 12 + xlen:
        si1, si2, slen = raw[p], raw[p + 1], struct.unpack_from("<H", raw, p + 2)[0]
        if (si1, si2) == (66, 67):
            assert slen == 2
            bsize = struct.unpack_from("<H", raw, p + 4)[0] + 1
        p += 4 + slen
    assert p == coff + 12 + xlen and bsize is not None, "BC subfield"
    cdata = raw[coff + 12 + xlen:coff + bsize - 8]
    crc, isize = struct.unpack_from("<II", raw, coff + bsize - 8)
    data = zlib.decompress(cdata, -15)
    assert len(data) == isize and zlib.crc32(data) & 0xFFFFFFFF == crc and isize <= 65536
    return bsize, data


def validate_bgzf(raw: bytes) -> int:
    """walk the whole file block by block; it must end with the 28-byte EOF marker block; returns the number of blocks"""
    off = n = 0
    while off < len(raw):
        bs, _ = _bgzf_block_at(raw, off)
        off += bs; n += 1
    assert off == len(raw)
    assert raw[-28:] == bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"), "EOF marker"
    return n


class Csi:
    def __init__(self, path: str):
        raw = gzip.open(path, "rb").read()                       # a .csi is itself BGZF-compressed
        assert raw[:4] == b"CSI\x01"
        self.min_shift, self.depth, l_aux = struct.unpack_from("<iii", raw, 4)
        aux = raw[16:16 + l_aux]
        # tabix aux: format, col_seq, col_beg, col_end, meta, skip, l_nm, names (NUL-terminated, concatenated)
        self.format, self.col_seq, self.col_beg, self.col_end, self.meta, self.skip, l_nm = struct.unpack_from("<iiiiiii", aux, 0)
        names = aux[28:28 + l_nm]
        assert len(names) == l_nm and (l_nm == 0 or names[-1] == 0)
        self.names = [x.decode() for x in names.split(b"\0")[:-1]]
        p = 16 + l_aux
        n_ref = struct.unpack_from("<i", raw, p)[0]; p += 4
        assert n_ref == len(self.names)
        self.bins = []
        for _ in range(n_ref):
            n_bin = struct.unpack_from("<i", raw, p)[0]; p += 4
            d = {}
            for _ in range(n_bin):
                b, loff, n_chunk = struct.unpack_from("<IQi", raw, p); p += 16
                ch = [struct.unpack_from("<QQ", raw, p + 16 * k) for k in range(n_chunk)]; p += 16 * n_chunk
                assert b not in d
                d[b] = (loff, ch)
            self.bins.append(d)
        assert len(raw) - p in (0, 8)                            # optional n_no_coor

    def reg2bins(self, beg: int, end: int):
        """SAM spec 5.3, reg2bins for CSI: bins that may overlap the 0-based half-open interval [beg, end)"""
        out = []
        end -= 1
        s = self.min_shift + self.depth * 3
        t = 0
        for l in range(self.depth + 1):
            b = t + (beg >> s); e = t + (end >> s)
            out.extend(range(b, e + 1))
            s -= 3; t += 1 << (l * 3)
        return out

    def max_bin(self) -> int:
        return ((1 << (self.depth + 1) * 3) - 1) // 7


def query(pat_gz: str, csi: Csi, chrom: str, lo: int, hi: int) -> bytes:
    """`tabix pat_gz chrom:lo-hi` (1-based closed on the column col_beg == col_end): the matching lines, in file order"""
    if chrom not in csi.names:
        return b""
    tid = csi.names.index(chrom)
    raw = open(pat_gz, "rb").read()
    beg0, end0 = lo - 1, hi                                      # 0-based half-open
    bins = csi.bins[tid]
    pseudo = csi.max_bin() + 1
    chunks = []
    for b in csi.reg2bins(beg0, end0):
        if b in bins and b != pseudo:
            chunks += bins[b][1]
    chunks.sort()
    out = []
    seen_end = 0
    for cb, ce in chunks:
        cb = max(cb, seen_end)                                   # overlapping / adjacent chunks: never read a record twice
        if cb >= ce:
            continue
        seen_end = ce
        coff, uoff = cb >> 16, cb & 0xFFFF
        buf = b""
        while (coff << 16) < ce:
            bs, data = _bgzf_block_at(raw, coff)
            stop = (ce & 0xFFFF) if coff == (ce >> 16) else len(data)
            buf += data[uoff:stop]
            coff += bs; uoff = 0
        for line in buf.splitlines(keepends=True):
            t = line.split(b"\t")
            if t[csi.col_seq - 1].decode() != chrom:
                continue
            v = int(t[csi.col_beg - 1])
            if lo <= v <= hi:
                out.append(line)
    return b"".join(out)
