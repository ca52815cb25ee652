#!/usr/bin/env python
"""The four CLIs under torchrun at 1 / 2 / 4 ... GPUs of one box on a multi-chromosome synthetic data set (BASELINE.json configs 3-5,
scaled by --scale so that generating the inputs does not dominate the GPU time of the call):
    python tools/multigpu_cli.py --scale 0.1 --gpus 1,2,4 --out gpurun_out/r2e_multigpu.json
  bam2pat   a coordinate-sorted BAM over 25 chromosomes (hg38 lengths x scale): chromosomes dealt to the ranks by LPT on their record
            counts, ONE NCCL reduce of the int32 beta counts, parts gathered in chromosome order (reference bam2pat.py:319-346,398-422)
  pat2beta  the .pat.gz of that run: records sharded by line ranges, one reduce
  homog     the same pat over blocks tiling the genome: records sharded, one reduce of the bins (reference homog.cpp:84 shards by --chrom)
  segment   K betas over the whole genome: chunks dealt round-robin, borders all-gathered (reference segment.py:129-165)
Every output must be byte-identical at every GPU count.  work_s is the command's own wall clock (WGBS_TIMING: started once the interpreter, the library and the NCCL communicator are up);
the whole-command wall clock and the start-up of a no-op command at the same GPU count are reported beside it."""
import argparse
import gzip
import hashlib
import json
import os
import subprocess
import sys
import time
from concurrent.futures import ProcessPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
HG38 = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717, 133797422, 135086622, 133275309, 114364328,
        107043718, 101991189, 90338345, 83257441, 80373285, 58617616, 64444167, 46709983, 50818468, 156040895, 57227415, 16569]
NAMES = [f"chr{i}" for i in range(1, 23)] + ["chrX", "chrY", "chrM"]


def _chrom_job(args):
    ci, name, length, first_idx, n_reads, seed = args
    from wgbs_tools_b200 import synth
    g = synth.make_genome(seed + ci, name, length)
    g.first_idx = first_idx
    sam = synth.make_sam(g, n_reads, seed + 100 + ci, paired=True, name_prefix=f"c{ci}_") if n_reads else b""
    return ci, g.loci, sam


def _beta_job(args):
    k, n, seed = args
    from wgbs_tools_b200 import synth
    return k, synth.make_betas(seed + k // 10, 1, n)[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=0.1)
    ap.add_argument("--gpus", default="1,2")
    ap.add_argument("--reads", type=int, default=4_000_000)
    ap.add_argument("--pat_records", type=int, default=30_000_000)
    ap.add_argument("--K", type=int, default=40)
    ap.add_argument("--out", default="gpurun_out/multigpu_cli.json")
    ap.add_argument("--dir", default=None)
    a = ap.parse_args()
    from wgbs_tools_b200 import bamio, synth
    from wgbs_tools_b200.patio import bgzf_compress
    work = a.dir or os.path.join("/dev/shm" if os.path.isdir("/dev/shm") else "/tmp", f"wgbs_mg_{os.getpid()}")
    os.makedirs(work, exist_ok=True)
    t0 = time.time()
    lens = [max(20_000, int(l * a.scale)) for l in HG38]
    tot = sum(lens)
    jobs, first = [], 1
    # loci first (cheap) to know the index ranges; reads proportional to length
    cores = max(1, len(os.sched_getaffinity(0)))
    for ci in range(25):
        jobs.append((ci, NAMES[ci], lens[ci], 1, int(a.reads * lens[ci] / tot) // 2 * 2, 500))
    with ProcessPoolExecutor(min(cores, 25)) as ex:
        res = sorted(ex.map(_chrom_job, jobs))
    loci = [r[1] for r in res]
    n_sites = int(sum(l.size for l in loci))
    refdir = os.path.join(work, "ref"); os.makedirs(refdir, exist_ok=True)
    with gzip.open(os.path.join(refdir, "CpG.bed.gz"), "wb", compresslevel=1) as f:
        idx = 1
        for ci, l in enumerate(loci):
            nm = NAMES[ci].encode()
            f.write(b"".join(b"%s\t%d\t%d\n" % (nm, x, idx + k) for k, x in enumerate(l.tolist())))
            idx += l.size
    open(os.path.join(refdir, "CpG.chrome.size"), "w").write("".join(f"{NAMES[ci]}\t{loci[ci].size}\n" for ci in range(25)))
    open(os.path.join(refdir, "chrome.size"), "w").write("".join(f"{NAMES[ci]}\t{lens[ci]}\n" for ci in range(25)))
    np.save(os.path.join(refdir, "CpG.bed.gz.loci.npy"), np.concatenate(loci).astype(np.uint32))
    sam = b"".join(r[2] for r in res)
    n_rec = sam.count(b"\n")
    bam = bamio.sam_to_bam(sam, list(zip(NAMES, lens)), procs=min(cores, 32))
    bam_path = os.path.join(work, "sample.bam"); open(bam_path, "wb").write(bam)
    del sam, res
    # pat for pat2beta / homog: per chromosome, concatenated in order
    parts = []
    first = 1
    for ci in range(25):
        n = loci[ci].size
        r = int(a.pat_records * n / n_sites)
        if r:
            parts.append(synth.make_pat_text_fast(900 + ci, r, n, chrom=NAMES[ci], first_idx=first))
        first += n
    pat_text = b"".join(parts)
    pat_path = os.path.join(work, "big.pat.gz"); open(pat_path, "wb").write(bgzf_compress(pat_text, min(cores, 32)))
    pat_records = pat_text.count(b"\n")
    del parts, pat_text
    blocks = synth.make_blocks(5, 1, n_sites)
    allloci = np.concatenate(loci)
    bounds = np.concatenate([[1], 1 + np.cumsum([l.size for l in loci])])
    keep = np.searchsorted(bounds, blocks[:, 0], side="right") == np.searchsorted(bounds, blocks[:, 1] - 1, side="right")      # blocks do not span chromosomes
    blocks = blocks[keep]
    cidx = np.searchsorted(bounds, blocks[:, 0], side="right") - 1
    bed = os.path.join(work, "blocks.bed")
    with open(bed, "w") as f:
        for (s, e), c in zip(blocks.tolist(), cidx.tolist()):
            f.write(f"{NAMES[c]}\t{int(allloci[s - 1])}\t{int(allloci[e - 2]) + 2}\t{s}\t{e}\n")
    with ProcessPoolExecutor(min(cores, a.K)) as ex:
        betas = sorted(ex.map(_beta_job, [(k, n_sites, 700) for k in range(a.K)]))
    beta_paths = []
    for k, b in betas:
        p = os.path.join(work, f"t{k}.beta"); b.tofile(p); beta_paths.append(p)
    del betas
    gen_s = time.time() - t0
    print(f"[mg] inputs: {n_rec:,} records in a {os.path.getsize(bam_path) / 1e6:.0f} MB BAM, {n_sites:,} CpGs, pat {pat_records:,} records, {blocks.shape[0]:,} blocks, "
          f"{a.K} betas ({gen_s:.0f}s)", file=sys.stderr, flush=True)

    def launch(n, cmd, outdir):
        os.makedirs(outdir, exist_ok=True)
        base = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1", "--master-port", str(29500 + n),
                "-m", "wgbs_tools_b200.cli"] if n > 1 else [sys.executable, "-m", "wgbs_tools_b200.cli"]
        t = time.time()
        r = subprocess.run(base + cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900, env=dict(os.environ, WGBS_TIMING="1"))
        return time.time() - t, r

    def digest(paths):
        h = {}
        for p in paths:
            data = gzip.open(p, "rb").read() if p.endswith(".gz") else open(p, "rb").read()
            h[os.path.basename(p)] = hashlib.sha256(data).hexdigest()[:16]
        return h

    results = {"inputs": {"records": n_rec, "bam_bytes": os.path.getsize(bam_path), "sites": n_sites, "pat_records": pat_records, "blocks": int(blocks.shape[0]), "K": a.K,
                          "scale": a.scale, "generated_in_s": gen_s}, "runs": {}}
    ref_digest = {}
    gpu_counts = [int(x) for x in a.gpus.split(",") if x]
    for n in gpu_counts:
        out = os.path.join(work, f"out{n}")
        run = {}
        t_noop = min(launch(n, ["noop"], out)[0] for _ in range(2))
        run["startup_s"] = t_noop
        cmds = {
            "bam2pat": (["bam2pat", bam_path, "--genome", refdir, "-o", out, "-f"], [f"{out}/sample.pat.gz", f"{out}/sample.beta"], n_rec, "reads/s"),
            "pat2beta": (["pat2beta", pat_path, "--genome", refdir, "-o", out, "-f"], [f"{out}/big.beta"], n_sites, "CpG-sites/s"),
            "homog": (["homog", pat_path, "-b", bed, "-o", out, "-f", "--genome", refdir], [f"{out}/big.uxm.bed.gz"], n_sites, "CpG-sites/s"),
            "segment": (["segment", "--betas", *beta_paths, "--genome", refdir, "-o", f"{out}/seg.bed"], [f"{out}/seg.bed"], n_sites, "CpG-sites/s"),
        }
        for name, (cmd, outs, units, unit) in cmds.items():
            try:
                t, r = launch(n, cmd, out)
                ok = r.returncode == 0 and all(os.path.isfile(p) for p in outs)
                d = digest(outs) if ok else {}
                if n == gpu_counts[0]:
                    ref_digest[name] = d
                import re
                m = re.search(r"\] \w+: ([0-9.]+) s after start-up", r.stderr)
                work_s = float(m.group(1)) if m else max(t - t_noop, 1e-3)       # the command's own clock (process group up before it starts)
                run[name] = {"wall_s": t, "work_s": work_s, "rate": units / work_s, "unit": unit, "rc": r.returncode, "identical_to_first_run": d == ref_digest.get(name), "digest": d}
                if not ok:
                    run[name]["stderr_tail"] = r.stderr[-1500:]
            except Exception as e:
                run[name] = {"error": repr(e)}
            print(f"[mg] N={n} {name}: {json.dumps({k: v for k, v in run[name].items() if k not in ('digest', 'stderr_tail')})}", file=sys.stderr, flush=True)
        results["runs"][str(n)] = run
    os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
    json.dump(results, open(a.out, "w"), indent=1)
    print(json.dumps({n: {k: (round(v["work_s"], 2) if isinstance(v, dict) and "work_s" in v else v) for k, v in r.items()} for n, r in results["runs"].items()}))
    if a.dir is None:
        subprocess.run(["rm", "-rf", work])


if __name__ == "__main__":
    main()
