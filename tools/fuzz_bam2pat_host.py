"""Fuzzer for bam2pat's host logic (run by hand: `python tools/fuzz_bam2pat_host.py SEED`): random BAMs of one library type (paired
or single-end) through the CLI three ways -- whole file, streamed in parts of random size (--bam_decode stream), chromosomes piled
up in template windows -- with the oracle port standing in for the device (tests/test_bam2pat_host.py PortContext).  The three
outputs (.pat.gz text, .beta) must be identical.  TEST INFRASTRUCTURE, CPU only."""
import sys, os, gzip, random, tempfile, pathlib, importlib.util
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import test_bam2pat_host as T
from wgbs_tools_b200 import api, bamio, synth, bam2pat
api.Context = T.PortContext
spec = importlib.util.spec_from_file_location("fz", os.path.join(ROOT, "tools", "fuzz_stream.py"))
src = open(os.path.join(ROOT, "tools", "fuzz_stream.py")).read()
# reuse rnd_sam only
ns = {"__file__": os.path.join(ROOT, "tools", "fuzz_stream.py")}
exec(src[:src.index("bad = 0")].replace('random.seed(int(sys.argv[1]) if len(sys.argv) > 1 else 0)', 'pass').replace('RUNS = len(sys.argv) > 2 and sys.argv[2] == "runs"', ''), ns)
random.seed(int(sys.argv[1])); ns['random'].seed(int(sys.argv[1]))
bad = 0
for it in range(12):
    sam, lens = ns['rnd_sam'](random.randint(1, 3))
    if not sam: continue
    # genome names must look like chromosomes: rename c0.. -> chr1..
    for k in range(len(lens)):
        sam = sam.replace(f"\tc{k}\t".encode(), f"\tchr{k+1}\t".encode())
    lens = [(f"chr{k+1}", L) for k, (_, L) in enumerate(lens)]
    def cg(l):
        t = l.split(b"\t")
        if len(t) > 9 and t[9] != b"*" and t[9] != b"ACGT":
            n = len(t[9]); t[9] = (b"CG" * (n // 2 + 1))[:n] if random.random() < 0.5 else (b"TG" * (n // 2 + 1))[:n]
        if PE and t[1] in (b"0", b"16"):
            t[1] = b"73" if t[1] == b"0" else b"89"          # a paired library: singles are paired reads whose mate is unmapped
        if not PE and int(t[1]) & 1:
            t[1] = b"0" if int(t[1]) & 64 else b"16"; t[6] = b"*"; t[7] = b"0"; t[8] = b"0"; t[0] += b"/%d" % (1 if int(t[1]) == 0 else 2)
        return b"\t".join(t)
    PE = random.random() < 0.6
    sam = b"\n".join(cg(l) for l in sam.splitlines() if not l.startswith(b"u\t")) + b"\n"
    tmp = pathlib.Path(tempfile.mkdtemp())
    refdir = tmp / "g"; refdir.mkdir()
    gs = []; first = 1
    for name, L in lens:
        g = synth.make_genome(5, name, max(L, 3000), first_idx=first, with_bases=False); first += g.n_cpg; gs.append(g)
    with gzip.open(refdir / "CpG.bed.gz", "wb") as f: f.write(b"".join(g.dict_text() for g in gs))
    (refdir / "CpG.chrome.size").write_text("".join(f"{g.chrom}\t{g.n_cpg}\n" for g in gs)); (refdir / "chrome.size").write_text("".join(f"{g.chrom}\t{g.length}\n" for g in gs))
    (tmp / "s.bam").write_bytes(bamio.sam_to_bam(sam, lens))
    outs = []
    extra = random.choice([[], ["-F", "1796", "--include_flags", "1"], ["-r", "chr1"], ["--long", "--no_beta"], ["--bottom_strand"], ["--top_strand"],
                           ["--min_cpg", "2", "--clip", "3"], ["-r", "chr1:500-%d" % random.randint(600, 40_000)], ["-l"]])
    for tag, dec, env in (("w", "host", {}), ("s", "stream", {"WGBS_STREAM_BYTES": str(random.choice([1, 100_000, 400_000]))}), ("c", "host", {"WGBS_CHUNK_RECORDS": str(random.choice([50, 700])), "WGBS_GPU_STREAMS": str(random.choice([1, 2, 3]))})):
        os.environ.pop("WGBS_STREAM_BYTES", None); os.environ.pop("WGBS_GPU_STREAMS", None); os.environ["WGBS_CHUNK_RECORDS"] = "0"
        os.environ.update(env)
        o = tmp / tag; o.mkdir()
        try:
            bam2pat.main([str(tmp / "s.bam"), "--genome", str(refdir), "-o", str(o), "--bam_decode", dec, "-q", "0"] + extra)
            pat = gzip.decompress((o / "s.pat.gz").read_bytes()) if (o / "s.pat.gz").exists() else None
            beta = (o / "s.beta").read_bytes() if (o / "s.beta").exists() else None
            outs.append((pat, beta))
        except Exception as e:
            outs.append(("EXC", repr(e)[:200]))
    print("it", it, extra, [len(x[0]) if x[0] else 0 for x in outs])
    if not (outs[0] == outs[1] == outs[2]):
        bad += 1; print("MISMATCH", it, extra, [type(x[0]).__name__ + str(len(x[0]) if x[0] else 0) for x in outs], [x[1] if x[0]=="EXC" else "" for x in outs])
print("bad", bad)
