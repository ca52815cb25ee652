"""Device BAM front end (csrc/bamdev.cu), the part that can be pinned WITHOUT a GPU: inflate_core.cuh / bam_core.cuh are
host + device code, so tests/bamdev_core_check.cpp compiles them with g++ and this file checks them against zlib, printf,
the source SAM text and the library's host BAM reader (csrc/bam.cu, itself round-trip tested in test_bam_ingest.py).
The warp-cooperative part of the inflater runs here under a 32-lane lock-step emulation (one thread per lane, every
shuffle / ballot / __syncwarp a barrier)."""
import os
import re
import struct
import subprocess
import zlib

import numpy as np
import pytest

from wgbs_tools_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def check(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("bamdev") / "bamdev_core_check")
    r = subprocess.run(["g++", "-std=c++20", "-O2", "-o", exe, os.path.join(ROOT, "tests", "bamdev_core_check.cpp"), "-lz", "-lpthread"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    return exe


def frame(comp: bytes, data: bytes) -> bytes:
    assert len(comp) + 26 <= 65536
    return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", len(comp) + 25) + comp
            + struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data)))


def frame_shifted(comp: bytes, data: bytes, k: int) -> bytes:
    """the same BGZF block with a second extra subfield of k payload bytes in front of 'BC': moves the deflate payload by 4 + k bytes"""
    xlen = 6 + 4 + k
    return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff" + struct.pack("<H", xlen) + b"XX" + struct.pack("<H", k) + bytes(k) + b"BC\x02\x00"
            + struct.pack("<H", len(comp) + 12 + xlen + 8 - 1) + comp + struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data)))


def block(data: bytes, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, mem=8) -> bytes:
    co = zlib.compressobj(level, zlib.DEFLATED, -15, mem, strategy)
    return frame(co.compress(data) + co.flush(), data)


@pytest.fixture(scope="module")
def sam():
    g = synth.make_genome(7, "chrT", 300_000)
    return g, synth.make_sam(g, 6000, 3, paired=True)


def test_inflate_core_matches_zlib_on_every_block_type(check, sam, tmp_path):
    """stored / fixed / dynamic blocks, long codes, overlapping matches (dist < len), several deflate blocks per BGZF block,
    empty blocks; every block through the one-lane build, a sample through the emulated warp"""
    from wgbs_tools_b200.patio import BGZF_EOF
    _, s = sam
    rng = np.random.default_rng(1)
    datas = [s[:60000], s[60000:125000], b"", b"a", b"ab" * 30000, b"\0" * 65000, rng.integers(0, 256, 50000, dtype=np.uint8).tobytes(),
             rng.integers(0, 4, 65000, dtype=np.uint8).tobytes(), bytes(range(256)) * 200, s[200000:260000]]
    parts = []
    for d in datas:
        for lvl in (0, 1, 6, 9):
            if lvl == 0 and len(d) > 65000:
                continue
            parts.append(block(d, lvl))
        parts += [block(d, 6, zlib.Z_FIXED), block(d, 6, zlib.Z_HUFFMAN_ONLY), block(d, 6, zlib.Z_RLE), block(d, 9, zlib.Z_DEFAULT_STRATEGY, 1)]
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    d = s[300000:340000]
    parts.append(frame(co.compress(d[:10000]) + co.flush(zlib.Z_SYNC_FLUSH) + co.compress(d[10000:]) + co.flush(zlib.Z_FULL_FLUSH) + co.flush(), d))
    # the emulated warp is slow (thread barriers): put one block of every kind first
    order = [0, 4, 5, 6, len(parts) - 1, 36, 12, 20] + [i for i in range(len(parts)) if i not in (0, 4, 5, 6, len(parts) - 1, 36, 12, 20)]
    # every alignment of the payload (the bit reader starts inside a 16-byte chunk; a fixed / stored block header leaves it
    # with as few as 5 bits before the first probe)
    shifted = []
    for k in range(16):
        for st in (zlib.Z_FIXED, zlib.Z_DEFAULT_STRATEGY):
            d = s[1000 * k:1000 * k + 20000]
            co = zlib.compressobj(6, zlib.DEFLATED, -15, 8, st)
            shifted.append(frame_shifted(co.compress(d) + co.flush(), d, k))
        shifted.append(frame_shifted(zlib.compressobj(0, zlib.DEFLATED, -15).compress(s[:3000]) + zlib.compressobj(0, zlib.DEFLATED, -15).flush(), s[:3000], k) if False else
                       frame_shifted((lambda c: c.compress(s[:3000]) + c.flush())(zlib.compressobj(0, zlib.DEFLATED, -15)), s[:3000], k))
    p = tmp_path / "mix.bgzf"
    p.write_bytes(b"".join(parts[i] for i in order) + b"".join(shifted) + BGZF_EOF)
    r = subprocess.run([check, "inflate", str(p), "8"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert f"blocks {len(parts) + len(shifted) + 1} emu 8 " in r.stdout and r.stdout.strip().endswith("mismatches 0")
    # the team decoder (inflate3_core.cuh: bgzf_team_decode_k): one lane and lock-step teams of 8 / 16 / 32 lanes
    r = subprocess.run([check, "inflate3", str(p)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert f"blocks {len(parts) + len(shifted) + 1} " in r.stdout and r.stdout.strip().endswith("mismatches 0")


def test_inflate_core_rejects_what_zlib_rejects(check, sam, tmp_path, built_lib):
    """corrupt payloads: the verdict (ok / error) must agree with zlib block by block; never a crash"""
    _, s = sam
    rng = np.random.default_rng(3)
    good = block(s[:30000])
    parts = []
    for k in range(40):
        b = bytearray(good)
        pos = int(rng.integers(18, len(b) - 8))
        b[pos] ^= 1 << int(rng.integers(0, 8))
        parts.append(bytes(b))
    short = bytearray(good); short[-4:] = struct.pack("<I", 29999)            # ISIZE smaller than the stream
    longer = bytearray(good); longer[-4:] = struct.pack("<I", 30001)
    badcrc = bytearray(good); badcrc[-8] ^= 1                                  # the gzip trailer's CRC32 (htslib checks it)
    parts += [bytes(short), bytes(longer), bytes(badcrc)]
    p = tmp_path / "bad.bgzf"
    p.write_bytes(b"".join(parts))
    r = subprocess.run([check, "inflate", str(p), "2"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr                            # 0 = no DISAGREEMENT with zlib (+ CRC32 check)
    assert r.stdout.strip().endswith("mismatches 0")
    r = subprocess.run([check, "inflate3", str(p)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("mismatches 0"), r.stdout + r.stderr      # the team decoder: same verdicts
    # and the verdict on the last three is "reject"
    p2 = tmp_path / "bad3.bgzf"; p2.write_bytes(b"".join(parts[-3:]))
    for i in range(3):
        q = tmp_path / f"one{i}.bgzf"; q.write_bytes(parts[-3 + i] + good)
        from wgbs_tools_b200 import bamio
        from wgbs_tools_b200._lib import WgbsError
        with pytest.raises(WgbsError, match="inflate failed"):
            bamio.BamFile(str(q), threads=1)                                 # the host reader (zlib + crc32) rejects them too


def test_fmt_g_matches_printf(check):
    r = subprocess.run([check, "fmtg", "400000", "7"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip() == "fmtg mismatches 0", r.stdout + r.stderr


EXTRA = (b"x1\t0\tchrT\t500\t7\t10M2I5M3D20M4S\t*\t0\t0\t" + b"ACGTN" * 8 + b"A\t*\tXA:A:q\tXB:i:-5\tXC:i:300\tXD:i:70000\tXE:f:1.5"
         b"\tXF:Z:hello world\tXG:H:1AE3\tML:B:C,1,2,255\tXH:B:s,-3,400\tXI:B:f,0.25,2,1e-07,3.14159\tMM:Z:C+m?,0,1;\tRG:Z:grp1\n")
UNMAPPED = b"x2\t4\t*\t0\t0\t*\t*\t0\t0\t*\t*\n"


@pytest.mark.parametrize("seg,depth", [(64, 4), (1000, 4), (16384, 4), (1 << 20, 4), (50, 0), (1000, -3), (16384, -4)])
def test_record_boundaries_and_sam_text_round_trip(check, sam, tmp_path, built_lib, seg, depth):
    """SAM -> BAM (records spanning BGZF blocks) -> inflate core -> segment guess / walk / repair -> format core == SAM.
    depth 0 makes every guess wrong (entry = segment base): the repair loop alone must still find every record.
    depth < 0 injects isolated wrong guesses into otherwise good ones."""
    from wgbs_tools_b200 import bamio
    g, s = sam
    if depth == 0:
        s = s[: s.index(b"\n", 40000) + 1]
    full = EXTRA + s + UNMAPPED
    p = tmp_path / "t.bam"
    p.write_bytes(bamio.sam_to_bam(full, [("chrM", 16571), ("chrT", g.length)]))
    r = subprocess.run([check, "view", str(p), str(seg), str(depth)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    assert r.returncode == 0, r.stderr.decode()
    assert r.stdout == full
    stats = dict(zip(r.stderr.decode().split()[::2], r.stderr.decode().split()[1::2]))
    assert int(stats["records"]) == full.count(b"\n")
    if depth > 0:
        assert int(stats["wrong_guesses"]) == 0 and int(stats["rounds"]) == 1
    elif depth < 0:                                       # every |depth|-th guess deliberately wrong (near and far): one repair round,
        assert int(stats["wrong_guesses"]) > 0 and int(stats["rounds"]) == 2      # the wrong ones cannot spread to their neighbours
    else:
        assert int(stats["wrong_guesses"]) > 0


def test_long_records_spanning_many_segments(check, tmp_path, built_lib):
    """ONT-sized records (tens of KB, ML arrays of thousands of values): most segments hold no record start at all"""
    from wgbs_tools_b200 import bamio
    rng = np.random.default_rng(11)
    lines = []
    pos = 100
    for i in range(40):
        n = int(rng.integers(200, 60000))
        seq = bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), n))
        qual = bytes(rng.integers(33, 74, n, dtype=np.uint8))
        nm = int(seq.count(b"C") // 3)
        mm = b"MM:Z:C+m?" + b"".join(b",%d" % int(x) for x in rng.integers(0, 3, nm)) + b";"
        ml = b"ML:B:C" + b"".join(b",%d" % int(x) for x in rng.integers(0, 256, nm))
        lines.append(b"read%d\t%d\tchrT\t%d\t60\t%dM\t*\t0\t0\t%s\t%s\t%s\t%s\tqs:f:%s\n" % (i, 16 * (i & 1), pos, n, seq, qual, mm, ml, repr(float(np.float32(rng.random() * 40))).encode()))
        pos += int(rng.integers(1, 3000))
    full = b"".join(lines)
    p = tmp_path / "ont.bam"
    p.write_bytes(bamio.sam_to_bam(full, [("chrT", 10_000_000)]))
    with bamio.BamFile(str(p), threads=2) as b:
        host = b.view()
    r = subprocess.run([check, "view", str(p), "16384", "4"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    assert r.returncode == 0, r.stderr.decode()
    assert r.stdout == host                               # float tags print as %g on both sides


def test_filters_match_the_host_reader(check, sam, tmp_path, built_lib):
    """-q / -F / -f / region / FLAG equality / -r RG / -L / bedtools -v / head -N: core (bam_core.cuh passes()) == bam.cu"""
    from wgbs_tools_b200 import bamio
    g, s = sam
    rng = np.random.default_rng(5)
    # read groups on a third of the records, mapq varied
    lines = s.splitlines()
    for i in range(0, len(lines), 3):
        lines[i] += b"\tRG:Z:grp1" if i % 2 == 0 else b"\tXX:i:5\tRG:Z:other"
    for i in range(0, len(lines), 7):
        t = lines[i].split(b"\t"); t[4] = b"3"; lines[i] = b"\t".join(t)
    full = EXTRA + b"\n".join(lines) + b"\n" + UNMAPPED
    p = tmp_path / "f.bam"
    p.write_bytes(bamio.sam_to_bam(full, [("chrM", 16571), ("chrT", g.length)]))
    starts = np.sort(rng.integers(0, g.length - 2000, 40)); ends = starts + rng.integers(1, 1500, 40)
    keep = np.concatenate([[True], starts[1:] >= ends[:-1]]); starts, ends = starts[keep], ends[keep]
    iv = tmp_path / "iv.txt"
    iv.write_text("".join(f"{a} {b}\n" for a, b in zip(starts, ends)))
    cases = [
        dict(chrom="chrT", mapq=10, exclude_flags=1796, include_flags=3),
        dict(chrom="chrT", beg=20_000, end=21_000),
        dict(chrom="chrT", flag_eq=(99, 147)),
        dict(chrom="chrT", read_group="grp1"),
        dict(chrom="chrT", intervals=(starts, ends)),
        dict(chrom="chrT", intervals=(starts, ends), exclude_intervals=True, mapq=5),
        dict(chrom=None, max_records=200),
        dict(chrom="chrM"),
        dict(chrom=None, read_group="nosuch"),
        dict(chrom="chrT", key_window=(100_000, 200_000)),
        dict(chrom="chrT", key_window=(0, 100_000), mapq=10, exclude_flags=1796, include_flags=3),
        dict(chrom="chrT", key_window=(200_000, 1 << 40), flag_eq=(99, 147)),
    ]
    with bamio.BamFile(str(p), threads=2) as b:
        for kw in cases:
            exp = b.view(**kw)
            argv = [str(-1 if kw.get("chrom") is None else b.refs.index(kw["chrom"])), str(kw.get("mapq", 0)), str(kw.get("exclude_flags", 0)),
                    str(kw.get("include_flags", 0)), str(kw.get("beg", 0)), str(kw.get("end", 0)),
                    ",".join(map(str, kw["flag_eq"])) if kw.get("flag_eq") else "-", kw.get("read_group") or "-",
                    str(iv) if "intervals" in kw else "-", str(int(kw.get("exclude_intervals", False))), str(kw.get("max_records", 0))]
            if "key_window" in kw:
                argv += [str(kw["key_window"][0]), str(kw["key_window"][1])]
            r = subprocess.run([check, "view", str(p), "16384", "4"] + argv, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
            assert r.returncode == 0, r.stderr.decode()
            assert r.stdout == exp, kw
            assert kw.get("chrom") == "chrM" or kw.get("read_group") == "nosuch" or exp


def test_part_record_table_equals_the_host_part_reader(check, sam, tmp_path, built_lib):
    """the device's part mode (dbam_open_impl with a cut-off last record: first bad record = tail, segments behind it ignored),
    emulated on the CPU, == wgbs_bam_open_part: same records, same tail, for parts starting at a probed record in mid-file"""
    from wgbs_tools_b200 import bamio
    g, s = sam
    lines = s.splitlines()
    # a few very long records so that cut-off records span several 16 KiB segments and whole BGZF blocks
    big = b"L1\t0\tchrT\t150000\t60\t90000M\t*\t0\t0\t" + b"ACGT" * 22500 + b"\t" + b"F" * 90000
    lines = sorted(lines + [big, big.replace(b"L1", b"L2").replace(b"150000", b"150700")], key=lambda l: int(l.split(b"\t")[3]))
    full = b"\n".join(lines) + b"\n"
    p = tmp_path / "p.bam"
    p.write_bytes(bamio.sam_to_bam(full, [("chrT", g.length)]))
    table = bamio.bgzf_block_table(str(p)); nb = table[0].size
    raw = p.read_bytes()
    with open(p, "rb") as f:
        for b0, n in [(3, 2), (5, 1), (nb // 2, 3), (nb // 2 + 1, 7), (nb - 4, 4), (1, nb - 1)]:
            pr = bamio.probe_block(f, table, b0, 1)
            assert pr is not None
            data = raw[int(table[0][b0]): int(table[0][b0 + n - 1] + table[1][b0 + n - 1])]
            host = bamio.BamPart(data, ["chrT"], [g.length], pr[0], threads=2)
            exp = host.view()
            for seg in (4096, 16384):
                r = subprocess.run([check, "part", str(p), str(seg), "4", str(b0), str(n), str(pr[0]), "chrT"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
                assert r.returncode == 0, r.stderr.decode()
                assert r.stdout == exp, (b0, n, seg)
                assert r.stderr.decode().strip().endswith(f"nrec {host.nrecords()} tail {host.tail}"), (r.stderr.decode(), host.tail)
            host.close()


def test_first_key_core_equals_the_host_reader(check, sam, tmp_path, built_lib):
    """bam_first_key_k's logic (shared record code: template_key + passes) == wgbs_bam_first_key of the host reader"""
    from wgbs_tools_b200 import bamio
    g, s = sam
    p = tmp_path / "k.bam"
    p.write_bytes(bamio.sam_to_bam(s, [("chrT", g.length)]))
    part = bamio.BamPart(p.read_bytes(), threads=2)
    for kw in (dict(), dict(mapq=10, exclude_flags=1796, include_flags=3), dict(flag_eq=(99, 147))):
        for key in (0, 1000, 150_000, 299_000, 10**9):
            exp = part.first_key(0, key, **kw)
            argv = ["0", str(kw.get("mapq", 0)), str(kw.get("exclude_flags", 0)), str(kw.get("include_flags", 0)), "0", "0",
                    ",".join(map(str, kw["flag_eq"])) if kw.get("flag_eq") else "-", "-", "-", "0", "0", "0", "0", "firstkey", str(key)]
            r = subprocess.run([check, "view", str(p), "16384", "4"] + argv, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
            assert r.returncode == 0, r.stderr.decode()
            assert int(r.stdout.decode().strip()) == (-1 if exp is None else exp), (kw, key)
    part.close()



def deep_code_block(rng, data: bytes, maxbits: int = 15, lens=None) -> bytes:
    """a BGZF block holding ONE dynamic-Huffman deflate block of literals only whose literal/length code is deliberately deep
    (many codes longer than 10 bits: large second-level tables in the two-phase decoder; some sets exceed its arena and must take
    the warp-per-block decoder).  Written bit by bit here -- zlib never emits such shapes."""
    n = 257
    leaves = [0]
    while lens is None and len(leaves) < n:
        i, j = int(rng.integers(len(leaves))), int(rng.integers(len(leaves)))
        if leaves[j] > leaves[i] and rng.integers(3):
            i = j                                                  # prefer deep leaves: a skewed tree
        if leaves[i] >= maxbits:
            i = next(k for k, l in enumerate(leaves) if l < maxbits)
        leaves[i] += 1; leaves.append(leaves[i])
    lens = [int(x) for x in rng.permutation(leaves if lens is None else lens)]
    cnt = [0] * 17
    for l in lens:
        cnt[l] += 1
    cnt[0] = 0; code = 0; nxt = [0] * 17
    for b in range(1, 16):
        code = (code + cnt[b - 1]) << 1; nxt[b] = code
    codes = []
    for l in lens:
        codes.append(nxt[l]); nxt[l] += 1
    out = bytearray(); acc = 0; nacc = 0

    def put(v, nb):                                                # nb bits of v, least significant first (RFC 1951 3.1.1)
        nonlocal acc, nacc
        acc |= v << nacc; nacc += nb
        while nacc >= 8:
            out.append(acc & 0xFF); acc >>= 8; nacc -= 8

    def put_code(c, l):                                            # Huffman codes are packed most significant bit first
        put(int(format(c, f"0{l}b")[::-1], 2), l)
    put(1, 1); put(2, 2)                                           # final block, dynamic Huffman
    put(n - 257, 5); put(0, 5); put(19 - 4, 4)                     # HLIT, HDIST (one distance code), HCLEN (all 19 lengths)
    for sym in (16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15):
        put(0 if sym >= 16 else 4, 3)                              # code-length code: symbols 0..15 get 4 bits each (code = the symbol)
    for l in lens + [1]:                                           # the literal/length lengths, then the single distance code
        put_code(l, 4)
    for b in data:
        put_code(codes[b], lens[b])
    put_code(codes[256], lens[256])
    if nacc:
        out.append(acc & 0xFF)
    return frame(bytes(out), data)


def test_two_phase_decoder_tables_and_fallback(check, tmp_path):
    """(1) the two-level Huffman tables of the two-phase decoder on 4 000 random complete code sets, built by one lane and by teams of
    1 / 8 / 32 lanes (team_tables): every symbol decodes to itself through root + second level, sets that exceed the arena are handed
    back (E_FALLBACK) by both alike;
    (2) hand-written deflate blocks with deep literal codes: output == zlib's, and some of them do take the fallback decoder"""
    from wgbs_tools_b200.patio import BGZF_EOF
    r = subprocess.run([check, "tables", "4000", "7"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("mismatches 0"), r.stdout + r.stderr
    rng = np.random.default_rng(5)
    parts = []
    for k in range(60):
        data = rng.integers(0, 256, int(rng.integers(1, 3000)), dtype=np.uint8).tobytes()
        blk = deep_code_block(rng, data, maxbits=int(rng.integers(11, 16)))
        assert zlib.decompress(blk[18:-8], -15) == data            # the hand-written stream is a valid deflate stream
        parts.append(blk)
    # 250 codes of 14-15 bits fill 8 different 10-bit prefixes: 8 second-level tables of 32 entries do not fit next to the roots
    wide = list(range(1, 8)) + [15] * 244 + [14] * 6
    for k in range(3):
        data = rng.integers(0, 256, 2000, dtype=np.uint8).tobytes()
        blk = deep_code_block(rng, data, lens=wide)
        assert zlib.decompress(blk[18:-8], -15) == data
        parts.insert(5 * k + 1, blk)
    p = tmp_path / "deep.bgzf"
    p.write_bytes(b"".join(parts) + BGZF_EOF)
    r = subprocess.run([check, "inflate3", str(p)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("mismatches 0"), r.stdout + r.stderr
    fb = int(r.stderr.split("fallbacks")[1].split()[0])
    assert fb > 0, "no block exceeded the arena: the fallback path was not exercised"


def test_mm_ml_tags_found_in_bam_records_like_in_sam_text(check, sam, tmp_path, built_lib):
    """bam_core.cuh find_np_tags (MM/ML mode of the direct route, bamdev.cu bam_np_tags_k) against the rule the SAM tokenizer applies to
    text (sam.cu tag_kind; reference ont.cpp:418-438): the LAST "MM:Z:" / "Mm:Z:" field and the LAST "ML:B:C" / "Ml:B:C" field -- other
    array subtypes, other tags in front, behind and between, empty arrays, no tags at all"""
    import re
    from wgbs_tools_b200 import bamio
    g, _ = sam
    seq = g.bases[1000:1040].tobytes()
    odd = [b"MM:Z:C+m?,0,1;\tML:B:C,250,3", b"Mm:Z:C+m.,1;\tMl:B:C,200", b"XA:i:5\tMM:Z:C+h?,0;\tXB:Z:abc\tML:B:C,255\tXC:B:s,1,2", b"MM:Z:C+m?,5;\tMM:Z:C+m?,0;\tML:B:C,9\tML:B:C,7,8",
           b"MM:Z:C+m?,0;\tML:B:c,100", b"MM:Z:C+m?,0;", b"XZ:Z:none", b"MM:Z:C+m?,0;\tML:B:S,300", b"XF:f:1.5\tXH:H:1AE3\tMM:Z:C+C?,0;\tXI:B:f,0.25,2\tML:B:C,77\tXJ:A:q",
           b"MM:Z:C+m?;\tML:B:C", b"ML:B:C,5\tMM:Z:C+m.,0;", b"MM:Z:\tML:B:C,1", b"MN:Z:x\tMK:B:C,1,2"]
    lines = [b"r%d\t0\tchrT\t%d\t60\t40M\t*\t0\t0\t%s\t*\t%s\n" % (k, 1001 + k, seq, t) for k, t in enumerate(odd)]
    lines.append(b"bare\t0\tchrT\t2000\t60\t40M\t*\t0\t0\t%s\t*\n" % seq)
    p = tmp_path / "tags.bam"; p.write_bytes(bamio.sam_to_bam(b"".join(lines), [("chrT", g.length)]))
    r = subprocess.run([check, "nptags", str(p)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)
    assert r.returncode == 0, r.stderr
    want = []
    for l in lines:
        f = l.rstrip(b"\n").split(b"\t")
        mm = [x[5:] for x in f[11:] if re.match(rb"M[Mm]:Z:", x)]
        ml = [x[6:] for x in f[11:] if re.match(rb"M[Ll]:B:C", x)]
        want.append(b"\t".join([f[0], mm[-1] if mm else b"-", (ml[-1].lstrip(b",") or b".") if ml else b"-"]) + b"\n")
    assert r.stdout == b"".join(want)


def test_team_decoder_is_exact_however_badly_the_lanes_guess(sam, tmp_path):
    """inflate3_core.cuh: a lane's guessed walk is only USED once its predecessor has arrived exactly at its anchor; lanes whose anchor is
    missed are dropped and the predecessor walks on.  With a run-up of 8 bits instead of 1024 (-DWGBS_SYNC_BITS=8) four lanes in five
    are dropped -- the output must still equal zlib's on every block type, for teams of 1 / 8 / 16 / 32 lanes, and the verdicts on
    corrupt streams must still agree"""
    from wgbs_tools_b200.patio import BGZF_EOF
    exe = str(tmp_path / "check_sb8")
    r = subprocess.run(["g++", "-std=c++20", "-O2", "-DWGBS_SYNC_BITS=8", "-o", exe, os.path.join(ROOT, "tests", "bamdev_core_check.cpp"), "-lz", "-lpthread"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    _, s = sam
    rng = np.random.default_rng(2)
    parts = [block(s[:60000]), block(s[60000:125000], 9), block(s[130000:190000], 1), block(rng.integers(0, 4, 65000, dtype=np.uint8).tobytes()),
             block(b"ab" * 30000), block(s[200000:260000], 6, zlib.Z_FIXED), block(s[:3000], 0), block(b"")]
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    d = s[300000:360000]
    parts.append(frame(co.compress(d[:20000]) + co.flush(zlib.Z_SYNC_FLUSH) + co.compress(d[20000:45000]) + co.flush(zlib.Z_FULL_FLUSH) + co.compress(d[45000:]) + co.flush(), d))
    good = block(s[:50000])
    for k in range(12):
        b = bytearray(good); b[int(rng.integers(18, len(b) - 8))] ^= 1 << int(rng.integers(0, 8)); parts.append(bytes(b))
    p = tmp_path / "m.bgzf"; p.write_bytes(b"".join(parts) + BGZF_EOF)
    r = subprocess.run([exe, "inflate3", str(p)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
    assert r.returncode == 0 and r.stdout.strip().endswith("mismatches 0"), r.stdout + r.stderr
    started, dropped = (int(x) for x in re.search(r"team 32: deflate blocks \d+ lanes started (\d+) dropped (\d+)", r.stderr).groups())
    assert dropped * 2 > started, "the short run-up was meant to make most guesses fail"
