#!/usr/bin/env python
"""Fuzz the team decoder (csrc/inflate3_core.cuh) + byte replay on the CPU against zlib:  python tools/fuzz_inflate3.py [rounds] [seed]
Every round writes ~120 BGZF blocks -- random content kinds (SAM-like text, small alphabets, runs, noise), sizes, zlib levels /
strategies / memLevels, random Z_SYNC_FLUSH / Z_FULL_FLUSH points (several deflate blocks, empty stored blocks), a share of them with
one flipped bit -- and runs tests/bamdev_core_check.cpp `inflate3` on it, built with the shipped run-up and with an 8-bit run-up (most
lanes dropped): teams of 1 / 8 / 16 / 32 lanes must give zlib's bytes, or zlib's verdict, on every block."""
import os
import struct
import subprocess
import sys
import tempfile
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wgbs_tools_b200 import synth  # noqa: E402
from wgbs_tools_b200.patio import BGZF_EOF  # noqa: E402


def frame(comp: bytes, data: bytes) -> bytes:
    return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", len(comp) + 25) + comp
            + struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data)))


def main():
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.default_rng(seed)
    g = synth.make_genome(7, "chrT", 400_000)
    sam = synth.make_sam(g, 8000, 3, paired=True)
    with tempfile.TemporaryDirectory() as td:
        exes = []
        for sb in (None, 8):
            exe = os.path.join(td, f"check_{sb}")
            cmd = ["g++", "-std=c++20", "-O2", "-o", exe, os.path.join(ROOT, "tests", "bamdev_core_check.cpp"), "-lz", "-lpthread"] + ([f"-DWGBS_SYNC_BITS={sb}"] if sb else [])
            subprocess.run(cmd, check=True)
            exes.append(exe)
        for r in range(rounds):
            parts = []
            for _ in range(120):
                kind = int(rng.integers(0, 6)); n = int(rng.integers(0, 65000))
                if kind == 0:
                    o = int(rng.integers(0, len(sam) - n)); d = sam[o:o + n]
                elif kind == 1:
                    d = rng.integers(0, int(rng.integers(2, 9)), n, dtype=np.uint8).tobytes()
                elif kind == 2:
                    d = bytes(rng.integers(65, 91, int(rng.integers(1, 40)), dtype=np.uint8).tolist()) * (n // 8 + 1); d = d[:n]
                elif kind == 3:
                    d = rng.integers(0, 256, n // 4, dtype=np.uint8).tobytes()
                elif kind == 4:
                    o = int(rng.integers(0, len(sam) - n)); d = bytearray(sam[o:o + n])
                    for k in rng.integers(0, max(n, 1), n // 50):
                        d[int(k)] = int(rng.integers(0, 256))
                    d = bytes(d)
                else:
                    d = b"\0" * n
                co = zlib.compressobj(int(rng.integers(0, 10)), zlib.DEFLATED, -15, int(rng.integers(1, 10)),
                                      [zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED][int(rng.integers(0, 5))])
                comp = b""; p = 0
                for cut in sorted(int(x) for x in rng.integers(0, max(len(d), 1), int(rng.integers(0, 4)))):
                    comp += co.compress(d[p:cut]) + co.flush([zlib.Z_SYNC_FLUSH, zlib.Z_FULL_FLUSH][int(rng.integers(0, 2))]); p = cut
                comp += co.compress(d[p:]) + co.flush()
                if len(comp) + 26 > 65536:
                    continue
                blk = bytearray(frame(comp, d))
                if rng.random() < 0.15 and len(comp) > 4:
                    blk[18 + int(rng.integers(0, len(comp)))] ^= 1 << int(rng.integers(0, 8))
                parts.append(bytes(blk))
            path = os.path.join(td, "f.bgzf")
            open(path, "wb").write(b"".join(parts) + BGZF_EOF)
            for exe in exes:
                res = subprocess.run([exe, "inflate3", path], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=1800)
                ok = res.returncode == 0 and res.stdout.strip().endswith("mismatches 0")
                print(f"round {r} {os.path.basename(exe)}: {res.stdout.strip()} | {res.stderr.strip().splitlines()[-2] if res.stderr.strip() else ''}", flush=True)
                if not ok:
                    keep = os.path.join(ROOT, "gpurun_out", f"fuzz_inflate3_fail_{seed}_{r}.bgzf")
                    os.makedirs(os.path.dirname(keep), exist_ok=True); open(keep, "wb").write(open(path, "rb").read())
                    print(res.stderr[-2000:]); raise SystemExit(f"MISMATCH: file kept as {keep}")
    print("ok")


if __name__ == "__main__":
    main()
