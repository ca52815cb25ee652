"""Fuzzer for bamio.stream_parts (run by hand: `python tools/fuzz_stream.py SEED [runs]`): random coordinate-sorted BAMs -- long
reads spanning many BGZF blocks, mates far apart, piles of records at one position, empty chromosomes, unmapped tails -- read as
parts of random sizes (from one block up), whole file or chromosome block range by chromosome block range (`runs`); the union of
the yielded views must equal the whole-file views and no QNAME may be split over two items.  TEST INFRASTRUCTURE, CPU only."""
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from wgbs_tools_b200 import bamio
random.seed(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
RUNS = len(sys.argv) > 2 and sys.argv[2] == "runs"
def rnd_sam(nchrom):
    recs = []
    lens = []
    for c in range(nchrom):
        L = random.choice([2000, 50_000, 300_000]); lens.append((f"c{c}", L))
        n = random.choice([0, 3, 200, 3000])
        rows = []
        for i in range(n):
            rl = random.choice([30, 150, 150, 150, 5000]) if random.random() < 0.97 else 150_000
            pos = random.randint(1, max(1, L - 10))
            seq = "A" * (min(rl, 400) if rl < 100_000 else rl) ; cig = f"{len(seq)}M"
            name = f"q{c}_{i}"
            if random.random() < 0.6:   # pair
                far = random.random() < 0.1
                p2 = random.randint(1, L - 1) if far else min(L - 1, pos + random.randint(0, 500))
                a, b = min(pos, p2), max(pos, p2)
                f1, f2 = (99, 147) if random.random() < 0.5 else (83, 163)
                if far: f1 &= ~2; f2 &= ~2
                rows.append((a, f"{name}\t{f1}\tc{c}\t{a}\t60\t{cig}\t=\t{b}\t{b-a+100}\t{seq}\t*"))
                if random.random() < 0.95:
                    rows.append((b, f"{name}\t{f2}\tc{c}\t{b}\t60\t{cig}\t=\t{a}\t{-(b-a+100)}\t{seq}\t*"))
            else:
                rows.append((pos, f"{name}\t{random.choice([0,16])}\tc{c}\t{pos}\t60\t{cig}\t*\t0\t0\t{seq}\t*"))
            if random.random() < 0.02:  # same-position pile
                for k in range(random.randint(5, 60)):
                    rows.append((pos, f"{name}_d{k}\t0\tc{c}\t{pos}\t60\t{cig}\t*\t0\t0\t{seq}\t*"))
        rows.sort(key=lambda r: r[0])
        recs += [r[1] for r in rows]
    recs += ["u\t4\t*\t0\t0\t*\t*\t0\t0\tACGT\t*"] * random.choice([0, 2])
    return ("\n".join(recs) + "\n").encode() if recs else b"", lens
bad = 0
for it in range(40):
    sam, lens = rnd_sam(random.randint(1, 4))
    if not sam: continue
    path = f"/tmp/fuzz_stream_{os.getpid()}.bam"
    open(path, "wb").write(bamio.sam_to_bam(sam, lens))
    kw = random.choice([dict(), dict(mapq=10, exclude_flags=1796), dict(exclude_flags=1796, include_flags=3)])
    with bamio.BamFile(path, threads=2) as b:
        whole = {c: b.view(c, **kw) for c in b.refs}
    budget = random.choice([1, 70_000, 200_000, 1_000_000])
    got = {c: [] for c in whole}; where = {}; k = 0
    opener = lambda d, r, l, f: bamio.BamPart(d, r, l, f, threads=2)
    try:
        if RUNS:
            table = bamio.bgzf_block_table(path); names = [x[0] for x in lens]
            ranges = bamio.chrom_block_ranges(path, table, len(names))
            for ci, c in enumerate(names):
                rng = ranges[ci]
                if rng[1] <= rng[0]: continue
                for part, chrom, win, done in bamio.stream_parts(path, opener, lambda c: kw, budget, blocks=rng, refs0=names):
                    if chrom != c: continue
                    t = part.view(chrom, key_window=win, **kw)
                    for l in t.splitlines():
                        q = l.split(b"\t", 1)[0]
                        if where.setdefault((chrom, q), k) != k: bad += 1; print("SPLIT", it, q)
                    got[chrom].append(t); k += 1
        else:
            for part, chrom, win, done in bamio.stream_parts(path, opener, lambda c: kw, budget):
                t = part.view(chrom, key_window=win, **kw)
                for l in t.splitlines():
                    q = l.split(b"\t", 1)[0]
                    if where.setdefault((chrom, q), k) != k: bad += 1; print("SPLIT", it, q)
                got[chrom].append(t); k += 1
        for c in whole:
            if sorted(b"".join(got[c]).splitlines()) != sorted(whole[c].splitlines()):
                bad += 1; print("MISMATCH", it, c, budget, kw, len(whole[c].splitlines()), sum(len(x.splitlines()) for x in got[c]))
    except Exception as e:
        bad += 1; print("EXC", it, budget, repr(e)[:300])
print("bad", bad)
