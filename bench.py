#!/usr/bin/env python
"""bench.py -- bam2pat reads/s on BASELINE.json configs[1] (synthetic 150 bp PE WGBS reads, 1M records, chr19-sized
CpG index) on N B200s, next to the reference's own CPU pipeline.

A "step" = one pass of the whole bam2pat hot path over one batch of 1M alignment records per GPU:
    SAM text -> tokenise -> pair mates -> CIGAR/CpG calls -> mate merge -> beta counts (+ NCCL reduce at N>1, + uint8 trim)
             -> sort/collapse -> pat text
`value`  : inputs resident in HBM, outputs left in HBM (device timed, CUDA events, max over ranks).
`e2e`    : the same step through the public API with HOST buffers: pinned SAM text in, pat text + .beta bytes out.
`--impl reference`: the reference's unmodified executables (oracle/_ref, flags of its setup.py) as
    `match_maker | patter | sort -k2,2n -k3,3 | uniq -c | awk`, one pipeline per shard of the same workload, on the host cores.

Multi-GPU (weak scaling): every rank piles up its own 1M-record batch over the same chromosome index (the reference
shards the same way: one process per region, bam2pat.py:343); the only exchange is one NCCL reduce(sum) of the
int32[nCpG,2] beta counts to rank 0, before the non-linear uint8 trim.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CHR = "chr19"
CHR_LEN = 58_617_616
N_CPG = 1_100_000
METRIC = "bam2pat_reads_per_sec"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ----------------------------------------------------------------------------------------------------------------------
# workload
# ----------------------------------------------------------------------------------------------------------------------
_GENOME = None


def genome():
    global _GENOME
    if _GENOME is None:
        from wgbs_tools_b200 import synth
        t = time.time()
        _GENOME = synth.make_genome(19, CHR, CHR_LEN, n_cpg=N_CPG)
        log(f"[bench] genome {CHR}: {CHR_LEN:,} bp, {_GENOME.n_cpg:,} CpGs ({time.time() - t:.1f}s)")
    return _GENOME


def make_batch(n_reads: int, seed: int) -> bytes:
    from wgbs_tools_b200 import synth
    t = time.time()
    sam = synth.make_sam(genome(), n_reads, seed, paired=True)
    log(f"[bench] SAM batch seed {seed}: {sam.count(10):,} records, {len(sam) / 1e6:.1f} MB ({time.time() - t:.1f}s)")
    return sam


# ----------------------------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu: int):
        self.gpu, self.p, self.lines = gpu, None, []

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.p.stdout), daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def stop(self) -> dict:
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], None, set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------------
# the reference arm / CPU baseline  (the ONLY place bench.py executes oracle/: as the thing we are compared with)
# ----------------------------------------------------------------------------------------------------------------------
def reference_run(sam: bytes, shards: int, steps: int, warmup: int, opt: bool = False):
    """Run the reference pipeline on `shards` contiguous shards of the workload concurrently; returns seconds per step."""
    from oracle import harness as H
    if not H.have_ref():
        return None
    g = genome()
    tmp = tempfile.mkdtemp(prefix="wgbsref_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    dpath = os.path.join(tmp, "CpG.bed")
    with open(dpath, "wb") as f:
        f.write(g.dict_text())
    lines = sam.splitlines(keepends=True)
    per = (len(lines) + shards - 1) // shards
    paths = []
    for s in range(shards):
        chunk = lines[s * per:(s + 1) * per]
        if not chunk:
            continue
        p = os.path.join(tmp, f"s{s}.sam")
        with open(p, "wb") as f:
            f.writelines(chunk)
        paths.append(p)
    env = dict(os.environ); env["PATH"] = H.SHIM + os.pathsep + env.get("PATH", ""); env["LC_ALL"] = "C"
    cmd = (f"{H.tool('match_maker', opt)} < {{inp}} | {H.tool('patter', opt)} {dpath} {CHR} --min_cpg 1 --clip 0 2>/dev/null"
           " | sort -k2,2n -k3,3 | uniq -c | awk -v OFS='\\t' '{{print $2,$3,$4,$1}}' > {inp}.pat")
    times = []
    for it in range(warmup + steps):
        t0 = time.time()
        procs = [subprocess.Popen(cmd.format(inp=p), shell=True, env=env, stderr=subprocess.DEVNULL) for p in paths]
        rcs = [p.wait() for p in procs]
        dt = time.time() - t0
        if any(rcs):
            raise RuntimeError("reference pipeline failed")
        if it >= warmup:
            times.append(dt)
    nlines = sum(1 for p in paths for _ in open(p + ".pat", "rb"))
    # what one patter process spends before its first read: loading the CpG dictionary of the chromosome (patter.cpp:14-42; every
    # chromosome worker of the reference pays it once, every shard pipeline here)
    global REF_DICT_LOAD_S
    one = os.path.join(tmp, "one.sam")
    with open(one, "wb") as f:
        f.writelines(lines[:2])
    t0 = time.time()
    subprocess.run(cmd.format(inp=one), shell=True, env=env, stderr=subprocess.DEVNULL)
    REF_DICT_LOAD_S = time.time() - t0
    subprocess.run(["rm", "-rf", tmp])
    return float(np.mean(times)), len(paths), nlines


REF_DICT_LOAD_S = None


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ----------------------------------------------------------------------------------------------------------------------
# roofline bookkeeping: algorithmic bytes per launch of each hot kernel (DESIGN.md section "kernels")
# ----------------------------------------------------------------------------------------------------------------------
def algorithmic_bytes(kernel: str, n_rec: int, text_bytes: int, n_tmpl: int, seq_end_avg: float = 0.0, out_text_bytes: int = 0):
    """bytes one launch must move at minimum.  n for the sort kernels is not fixed (records when pairing,
    templates when collapsing): use the larger so the fraction is a lower bound."""
    kernel = kernel.strip("()").split("<")[0]            # the profiler reports template instances: nl_scan_k<0>, sam_lines_k<2>
    return {
        "nl_scan_k": text_bytes + 4 * n_rec,             # read the text ONCE; write one newline offset per line
        "sam_lines_k": int(seq_end_avg * n_rec) + 8 * n_rec + 49 * n_rec,   # read each line up to the end of SEQ (QUAL/tags are
                                                         # never read in bisulfite mode) + 2 newline offsets; write 12 words + status
        "sam_scan_k": text_bytes + 8 * n_rec + 49 * n_rec,   # fused tokenizer: the text ONCE; newline offsets written + read back; 12 words + status per record
        "tk_records_k": 48 * n_rec + 52 * n_rec + 32 * n_rec,  # read offsets (+ QNAME/FLAG/POS bytes), write 13 descriptor words
        "rs_onesweep_k": 16 * n_rec,                     # (key,val) read + written
        "rs_global_hist_k": 4 * n_rec,
        "pileup_measure_k": 36 * n_rec,
        # descriptors (44 B) + the CIGAR's sector + one 32-byte SEQ sector per candidate CpG (the two bases of a CpG are read
        # in place from the SAM text; 150 bp x N_CPG / CHR_LEN candidates per read)
        "pileup_call_k": int(n_rec * (44 + 32 + 32 * 150.0 * N_CPG / CHR_LEN)),
        "nl_count_k": text_bytes, "nl_write_k": text_bytes + 4 * n_rec,
        # pairing: a 24-byte slot per table entry (2 x records rounded up to a power of two), hash words + slot index per record;
        # the resolve step also compares the two QNAMEs of every pair (~10 bytes each, one 32-byte sector per name)
        "slots_init_k": 24 * (1 << (2 * n_rec - 1).bit_length()),
        "pair_insert_k": 12 * n_rec + 24 * n_rec,
        "pair_resolve_k": 4 * n_rec + 24 * n_rec + 32 * n_rec + 4 * n_rec,
        # mate overlay: per record idx / len / word offset in, per template idx / len / off / valid out + the pattern words of both mates
        "merge_templates_k": 12 * n_rec + 16 * n_rec + 8 * n_rec,
        "line_write_k": 16 * n_tmpl + 8 * n_tmpl + out_text_bytes,
    }.get(kernel)


def build_roofline(rep: dict, psteps: int, n_rec: int, text_bytes: int, n_tmpl: int, seq_end_avg: float, out_text_bytes: int, peak: float, how: str) -> dict:
    """the `roofline` object of the bench line from the library's per-kernel profile of `psteps` steps ({kernel: (launches, ms)}):
    the dominant kernel = the one with the largest share of the step among those whose algorithmic bytes are defined"""
    tot = sum(v[1] for v in rep.values())
    top = sorted(rep.items(), key=lambda kv: -kv[1][1])
    per = []
    for k, (c, ms) in top[:10]:
        b = algorithmic_bytes(k, n_rec, text_bytes, n_tmpl, seq_end_avg, out_text_bytes)
        if b:
            per.append({"kernel": k, "launches_per_step": c // psteps, "avg_launch_ms": ms / c, "achieved": b / (ms / c / 1e3) / 1e9,
                        "frac": b / (ms / c / 1e3) / 1e9 / peak, "share_of_step": ms / tot, "algorithmic_bytes_per_launch": b})
    breakdown = {k: round(v[1] / psteps, 4) for k, v in top[:10]}
    if not per:
        return {"bound": "hbm", "kernel": top[0][0] if top else None, "achieved": None, "peak": peak, "unit": "GB/s", "frac": None, "traffic": None,
                "peak_source": how, "breakdown_ms_per_step": breakdown}
    d = per[0]
    traffic = None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.isfile(tp):
        traffic = json.load(open(tp)).get(d["kernel"].strip("()").split("<")[0])
    return {"bound": "hbm", "kernel": d["kernel"], "achieved": d["achieved"], "peak": peak, "unit": "GB/s", "frac": d["frac"], "traffic": traffic,
            "peak_source": how, "algorithmic_bytes_per_launch": d["algorithmic_bytes_per_launch"], "avg_launch_ms": d["avg_launch_ms"],
            "share_of_step": d["share_of_step"], "per_kernel": per, "breakdown_ms_per_step": breakdown}


# ----------------------------------------------------------------------------------------------------------------------
# the other hot-path steps (BASELINE.json metric: "pat2beta CpG-sites/sec", homog, segment) -- reported under "extra"
# ----------------------------------------------------------------------------------------------------------------------
def extras(ctx, torch, peak, sam_for_bam=b""):
    import ctypes as C
    from oracle import harness as H
    from wgbs_tools_b200 import synth
    from wgbs_tools_b200._lib import check, lib
    out = {}
    ev = lambda: torch.cuda.Event(enable_timing=True)
    stream = torch.cuda.current_stream()

    def kernel_ms(fn, reps=2):
        """per-kernel device time of one call of fn (the library's own profiler: an event pair around every launch)"""
        ctx.prof(True)
        for _ in range(reps):
            fn()
        rep = ctx.prof_report()
        ctx.prof(False)
        return {k: round(v[1] / reps, 4) for k, v in sorted(rep.items(), key=lambda kv: -kv[1][1])[:8]}

    def dev_time(fn, reps=5, warm=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        a, b = ev(), ev()
        a.record(stream)
        for _ in range(reps):
            fn()
        b.record(stream)
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps / 1e3

    # ---- pat2beta + homog on a sorted pat of 16M records over the chr19-sized index (records + symbols ~390 MB > L2)
    N, R = N_CPG, 16_000_000
    t0 = time.time()
    txt = synth.make_pat_text_fast(3, R, N, chrom=CHR)
    log(f"[bench] pat text: {R:,} records, {len(txt) / 1e6:.1f} MB ({time.time() - t0:.1f}s)")
    d_txt = ctx.upload(txt)
    mc = ctx.alloc(N * 8)
    P = ctx.pats_from_text(d_txt)
    sec_parse = dev_time(lambda: ctx.pats_from_text(d_txt).free())
    sec_p2b = dev_time(lambda: ctx.pat2beta(P, 1, N + 1, meth_cov=mc))
    words = P.pool_words
    p2b_bytes = R * 16 + words * 4 + N * 8          # idx,len,count,off + symbol words + one int32 pair per site written
    out["pat2beta"] = {"records": R, "sites": N, "parse_text_ms": sec_parse * 1e3, "kernel_ms": sec_p2b * 1e3,
                       "kernels_ms": {"parse": kernel_ms(lambda: ctx.pats_from_text(d_txt).free()), "accumulate": kernel_ms(lambda: ctx.pat2beta(P, 1, N + 1, meth_cov=mc))},
                       "sites_per_sec": N / (sec_parse + sec_p2b), "records_per_sec": R / (sec_parse + sec_p2b),
                       "roofline": {"kernel": "pat2beta_k", "bound": "hbm", "achieved": p2b_bytes / sec_p2b / 1e9, "peak": peak, "unit": "GB/s",
                                    "frac": p2b_bytes / sec_p2b / 1e9 / peak, "algorithmic_bytes": p2b_bytes}}
    blocks = synth.make_blocks(5, 1, N)
    rng = np.array([0, 0.334, 0.667, 1], np.float32)
    bs = ctx.upload(np.ascontiguousarray(blocks[:, 0])); be = ctx.upload(np.ascontiguousarray(blocks[:, 1]))
    d_rng = ctx.upload(rng); d_out = ctx.alloc(blocks.shape[0] * 12)
    sec_h = dev_time(lambda: check(lib.wgbs_homog(ctx.h, P.h, bs.ptr, be.ptr, blocks.shape[0], d_rng.ptr, 3, 3, 0, d_out.ptr)))
    out["homog"] = {"records": R, "blocks": int(blocks.shape[0]), "ms": sec_h * 1e3, "records_per_sec": R / sec_h, "sites_per_sec": N / sec_h,
                    "kernels_ms": kernel_ms(lambda: check(lib.wgbs_homog(ctx.h, P.h, bs.ptr, be.ptr, blocks.shape[0], d_rng.ptr, 3, 3, 0, d_out.ptr)))}
    # reference CPU (single process, reference flags) on the same text
    if H.have_ref():
        sub = txt[: txt.index(b"\n", len(txt) // 8) + 1]          # bounded sample: first eighth of the records
        rs = sub.count(b"\n")
        t0 = time.time(); H.ref_stdin2beta(sub, 1, N + 1); c1 = time.time() - t0
        bp = H.write_tmp(synth.blocks_text(CHR, blocks), ".bed")
        t0 = time.time(); H.ref_homog(sub, bp, "0,0.334,0.667,1", 3); c2 = time.time() - t0
        os.remove(bp)
        out["pat2beta"]["cpu_reference"] = {"records_per_sec": rs / c1, "cores": 1, "sample": f"{rs:,} records, stdin2beta 1 {N + 1}"}
        out["homog"]["cpu_reference"] = {"records_per_sec": rs / c2, "cores": 1, "sample": f"{rs:,} records"}
    P.free()
    # ---- segment: K betas x S sites in 60000-site chunks (segment.py defaults: max_cpg 1000, max_bp 2000, pcount 15)
    K, S = 10, 240_000
    betas = synth.make_betas(9, K, S)
    loci = genome().loci[:S]
    dbet = [ctx.upload(b) for b in betas]; dd = ctx.upload(loci)
    chunks = [(s, min(60_000, S - s)) for s in range(0, S, 60_000)]
    t0 = time.time(); res = ctx.segment(dbet, dd, chunks, 1000, 2000, 15); torch.cuda.synchronize(); warm = time.time() - t0
    t0 = time.time(); res = ctx.segment(dbet, dd, chunks, 1000, 2000, 15); torch.cuda.synchronize(); sec_s = time.time() - t0
    # work of the DP: one cost cell per admissible (start, end) pair (max_cpg 1000, max_bp 2000, inside the chunk), K log-likelihood
    # terms per cell (one fp32 divide + log2f + fp64 log2 each: SURVEY 8d -- segment is bound by that arithmetic, not by HBM)
    l64 = loci.astype(np.int64); e_idx = np.arange(S)
    lo_i = np.maximum(np.maximum((e_idx // 60_000) * 60_000, e_idx + 1 - 1000), np.searchsorted(l64, l64 - 2000, side="left"))
    cells = int((e_idx - lo_i + 1).sum())
    hbm_min = 2 * K * S + 8 * S
    out["segment"] = {"K": K, "sites": S, "chunks": len(chunks), "ms": sec_s * 1e3, "sites_per_sec": S / sec_s,
                      "blocks": int(sum(len(r) - 1 for r in res)), "timing": "host wall clock around the C-ABI call (includes D2H of borders)",
                      "cost_cells": cells, "cell_terms_per_sec": cells * K / sec_s,
                      "roofline": {"bound": "arithmetic (fp32 divide + log2f + fp64 log2 per term) and the sequential DP chain per chunk; HBM minimum shown for scale",
                                   "hbm_min_bytes": hbm_min, "achieved": hbm_min / sec_s / 1e9, "peak": peak, "unit": "GB/s", "frac": hbm_min / sec_s / 1e9 / peak}}
    if H.have_ref():
        paths = [H.write_tmp(b.tobytes(), f".{i}.beta") for i, b in enumerate(betas)]
        t0 = time.time(); r0 = H.ref_segmentor(paths, 0, 60_000, 1000, 2000, 15, loci[:60_000]); c3 = time.time() - t0
        out["segment"]["cpu_reference"] = {"sites_per_sec": 60_000 / c3, "cores": 1, "sample": "first 60000-site chunk", "identical_borders": bool(np.array_equal(r0, res[0]))}
        for p in paths:
            os.remove(p)
    for b in dbet + [dd, d_txt, mc, bs, be, d_rng, d_out]:
        b.free()
    # ---- the MM/ML mode of the pileup (BASELINE config 5's flavour: single-end reads with MM:Z / ML:B:C tags, 5mC + 5hmC calls)
    try:
        n_np = 150_000
        t0 = time.time(); npsam = synth.make_np_sam(genome(), n_np, 7)
        log(f"[bench] MM/ML SAM batch: {n_np:,} records, {len(npsam) / 1e6:.1f} MB ({time.time() - t0:.1f}s)")
        d_np = ctx.upload(npsam)
        ix_np = ctx.load_index(genome().loci, 1)

        def np_step():
            Pn, _ = ctx.pileup_sam(ix_np, d_np)              # MM tag on the first line: MM/ML mode is auto-detected like patter does
            Pn.collapse(); Pn.free()
        sec_np = dev_time(np_step, reps=5, warm=2)
        Pn, st_np = ctx.pileup_sam(ix_np, d_np)
        Pn.collapse()
        got = Pn.to_text(CHR); Pn.free()
        out["pileup_mm_ml"] = {"records": n_np, "sam_bytes": len(npsam), "ms": sec_np * 1e3, "reads_per_sec": n_np / sec_np, "nanopore_mode": int(st_np["nanopore"]),
                               "templates": int(st_np["templates"]), "what": "tokenize + MM/ML decode + calls + collapse, inputs resident in HBM",
                               "kernels_ms": kernel_ms(np_step)}
        if H.have_ref():
            sub = npsam[: npsam.index(b"\n", len(npsam) // 8) + 1]
            dp = H.write_tmp(genome().dict_text(), ".CpG.bed")
            sub2 = npsam[: npsam.index(b"\n", len(npsam) // 4) + 1]
            t0 = time.time(); ro, _ = H.ref_patter(sub, dp, CHR, False, nanopore=True); c_np = time.time() - t0
            t0 = time.time(); H.ref_patter(sub2, dp, CHR, False, nanopore=True); c_np2 = time.time() - t0
            os.remove(dp)
            gsub, _ = ctx.pileup_sam(ix_np, sub)
            gsub.collapse()
            # the dictionary load (~3 s for 1.1M CpGs through the tabix stand-in) is taken out by timing two sample sizes
            per_read = max(c_np2 - c_np, 1e-9) / max(sub2.count(10) - sub.count(10), 1)
            out["pileup_mm_ml"]["cpu_reference"] = {"reads_per_sec": 1.0 / per_read, "cores": 1,
                                                    "sample": f"patter --nanopore on {sub.count(10):,} and {sub2.count(10):,} records; rate from the difference (dictionary load excluded)",
                                                    "identical_pat": bool(gsub.to_text(CHR) == H.ref_collapse(ro))}
            gsub.free()
        d_np.free(); ix_np.free()
    except Exception as e:
        out["pileup_mm_ml"] = {"error": repr(e)}
    # ---- BAM ingest (host side: BGZF inflate + BAM -> SAM text on threads); what `samtools view` does in the reference pipeline
    try:
        from wgbs_tools_b200 import bamio
        sub = sam_for_bam[: sam_for_bam.index(b"\n", len(sam_for_bam) // 8) + 1]
        t0 = time.time(); bam = bamio.sam_to_bam(sub, [(CHR, CHR_LEN)]); tw = time.time() - t0
        path = os.path.join("/dev/shm" if os.path.isdir("/dev/shm") else "/tmp", f"wgbs_bench_{os.getpid()}.bam")
        open(path, "wb").write(bam)
        res = {}
        for th in (1, host_threads()):
            t0 = time.time(); bf = bamio.BamFile(path, th); t1 = time.time() - t0
            t0 = time.time(); txt_out = bf.view(CHR, 10, 1796, 3); t2 = time.time() - t0
            nrec = bf.nrecords(); bf.close()
            res[f"threads_{th}"] = {"open_ms": t1 * 1e3, "view_ms": t2 * 1e3, "records_per_sec": nrec / (t1 + t2)}
        os.remove(path)
        out["bam_ingest"] = {"records": nrec, "bam_bytes": len(bam), "sam_bytes": len(txt_out), "identical_to_input_sam": bool(txt_out == sub), **res,
                             "note": "host only (no GPU): inflate + record walk (open) and SAM formatting with -q 10 -F 1796 -f 3 (view)"}
    except Exception as e:
        out["bam_ingest"] = {"error": repr(e)}
    return out


def bam_device_leg(tmp: str, n_rec: int, sam_bytes: int, peak: float, configs: str, steps=8, warmup=3):
    """(child process of the bench) The same batch end to end from the COMPRESSED BAM bytes in pinned host memory (SURVEY 8f-1:
    no SAM-text detour over PCIe): upload + BGZF inflate + record table + view + pileup + pat2beta + collapse + pat text and .beta
    read back.  Outputs are compared with the SAM-text path's (files written by the parent).
    configs: "key:direct:inflate:streams,..." -- route (WGBS_DBAM_DIRECT), decoder (WGBS_INFLATE, empty = default) and the number of
    batches in flight (S Contexts with their own streams on S host threads: the upload of one batch overlaps the kernels of
    another); one JSON line per configuration, printed as soon as it is measured."""
    import ctypes as C
    import torch
    from wgbs_tools_b200._lib import PileupOpts, ViewOpts, check, lib
    from wgbs_tools_b200.api import Context
    bam_bytes = open(os.path.join(tmp, "batch.bam"), "rb").read()
    ref_text = open(os.path.join(tmp, "ref.pat"), "rb").read(); ref_beta = open(os.path.join(tmp, "ref.beta"), "rb").read()
    loci = np.load(os.path.join(tmp, "loci.npy"))
    n_cpg = int(loci.size)
    torch.cuda.set_device(0)
    main_stream = torch.cuda.Stream(); torch.cuda.set_stream(main_stream)
    ctx0 = Context(0, stream=main_stream.cuda_stream)
    ix = ctx0.load_index(loci, 1)
    h_bam = torch.frombuffer(bytearray(bam_bytes), dtype=torch.uint8).pin_memory()
    torch.cuda.synchronize()

    def run_config(S: int) -> dict:
        streams = [main_stream] if S == 1 else [torch.cuda.Stream() for _ in range(S)]
        ctxs = [ctx0] if S == 1 else [Context(0, stream=st.cuda_stream) for st in streams]
        h_text = [torch.empty(max(len(ref_text) * 2, 1 << 20), dtype=torch.uint8).pin_memory() for _ in range(S)]
        h_beta = [torch.empty((n_cpg, 2), dtype=torch.uint8).pin_memory() for _ in range(S)]
        mc = [torch.zeros((n_cpg, 2), dtype=torch.int32, device="cuda") for _ in range(S)]
        out_n = [{} for _ in range(S)]; errs = []
        torch.cuda.synchronize()

        def step(w: int):
            ctx = ctxs[w]
            B = C.c_void_p()
            check(lib.wgbs_dbam_open(ctx.h, h_bam.data_ptr(), h_bam.numel(), C.byref(B)))
            out_n[w]["inflated"] = int(lib.wgbs_dbam_inflated_bytes(B))
            vo = ViewOpts(); vo.refid = 0
            o = PileupOpts(1, 0, -1, 0, 0, 0.67, b"C")
            h = C.c_void_p(); st = (C.c_uint64 * 8)()
            check(lib.wgbs_pileup_dbam(ctx.h, ix.h, B, C.byref(vo), C.addressof(o), C.byref(h), C.addressof(st), None))
            lib.wgbs_dbam_close(ctx.h, B)
            check(lib.wgbs_pat2beta(ctx.h, h, 1, n_cpg + 1, mc[w].data_ptr(), 1))
            check(lib.wgbs_collapse(ctx.h, h))
            n = C.c_size_t()
            check(lib.wgbs_pats_format(ctx.h, h, CHR.encode(), h_text[w].data_ptr(), h_text[w].numel(), C.byref(n)))
            check(lib.wgbs_trim(ctx.h, mc[w].data_ptr(), n_cpg, 8, h_beta[w].data_ptr()))
            lib.wgbs_pats_free(ctx.h, h)
            out_n[w].update(n=n.value, lines=int(st[0]))

        def work(w: int, k: int):
            try:
                torch.cuda.set_device(0)
                for _ in range(k):
                    step(w)
            except Exception as e:
                errs.append(repr(e))

        def run(total: int):
            if S == 1:
                return work(0, total)
            th = [threading.Thread(target=work, args=(w, total // S + (1 if w < total % S else 0))) for w in range(S)]
            for t in th:
                t.start()
            for t in th:
                t.join()

        run(warmup * S)
        torch.cuda.synchronize()
        if errs:
            return {"error": errs[0]}
        same = all(h_text[w][:out_n[w]["n"]].numpy().tobytes() == ref_text and h_beta[w].numpy().tobytes() == ref_beta for w in range(S))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main_stream)
        for st in streams:
            if st is not main_stream:
                st.wait_event(e0)
        run(steps)
        for st in streams:
            if st is not main_stream:
                ev = torch.cuda.Event(); ev.record(st); main_stream.wait_event(ev)
        e1.record(main_stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        ctxs[0].prof(True)
        for _ in range(2):
            step(0)
        rep = ctxs[0].prof_report()
        ctxs[0].prof(False)
        top = sorted(rep.items(), key=lambda kv: -kv[1][1])
        res = {"records": n_rec, "lines_seen": out_n[0]["lines"], "ms_per_step": ms, "reads_per_sec": n_rec / (ms / 1e3), "h2d_bytes_per_step": len(bam_bytes),
               "d2h_bytes_per_step": out_n[0]["n"] + 2 * n_cpg, "sam_text_bytes_equivalent": sam_bytes, "inflated_bytes": out_n[0]["inflated"],
               "route": "direct (BAM records -> pileup kernels, no SAM text)" if os.environ.get("WGBS_DBAM_DIRECT") == "1" else "text (view -> SAM text -> tokenizer)",
               "inflate": os.environ.get("WGBS_INFLATE") or "default (warp per BGZF block)", "batches_in_flight": S,
               "identical_to_sam_text_path": bool(same) and not errs, "errors": errs[:3],
               "breakdown_ms_per_step": {k: round(v[1] / 2, 4) for k, v in top[:12]},
               "mode": ("serial: upload, inflate, view, pileup, read back, one batch after the other" if S == 1 else
                        f"{S} batches in flight on {S} streams (one Context per host thread): uploads overlap the kernels of the other batches")
                       + "; wgbs_dbam_open + wgbs_pileup_dbam from pinned host bytes"}
        infl_name = next((k for k in rep if k.startswith("bgzf_inflate")), None)      # bgzf_inflate_k<2>, bgzf_inflate_team_k<G>
        infl = rep.get(infl_name)
        if infl:
            sec = infl[1] / infl[0] / 1e3
            ab = len(bam_bytes) + out_n[0]["inflated"]          # algorithmic bytes: compressed bytes read + inflated bytes written
            res["roofline"] = {"kernel": infl_name, "bound": "instruction issue of the serial Huffman walk per block; reported against hbm", "achieved": ab / sec / 1e9,
                               "peak": peak, "unit": "GB/s", "frac": ab / sec / 1e9 / peak, "algorithmic_bytes": ab, "avg_launch_ms": sec * 1e3}
        if S > 1:
            for c in ctxs:
                c.close()
        return res

    for cfg in configs.split(","):
        key, direct, inflate, S = cfg.split(":")
        os.environ["WGBS_DBAM_DIRECT"] = direct
        if inflate:
            os.environ["WGBS_INFLATE"] = inflate
        else:
            os.environ.pop("WGBS_INFLATE", None)
        try:
            res = run_config(int(S))
        except Exception as e:
            res = {"error": repr(e)}
        res["key"] = key
        log(f"[bench] {key}: {res.get('ms_per_step', float('nan')):.3f} ms/step, {res.get('reads_per_sec', 0) / 1e6:.1f} M reads/s, identical={res.get('identical_to_sam_text_path')}")
        print(json.dumps(res), flush=True)


def stream_leg(tmp: str, n_rec: int, steps=12, warmup=3):
    """(child process of the bench) The device-resident step of the main line, with S batches in flight on S streams: one
    Context (own stream, own scratch) per host thread -- chromosomes are independent units of work, so several can be in flight on
    one GPU (`bam2pat --gpu_streams S` does the same in the CLI).  A single
    stream leaves the GPU idle while the host reads back sizes between kernels (~10 round trips per step) and runs kernels of
    one wave or less back to back; a second stream fills those gaps.  S = 1 repeats the main line's `value` as the control.
    Timed on the device: a start event every worker stream waits for, an end event that waits for every worker stream."""
    import ctypes as C
    import torch
    from wgbs_tools_b200._lib import PileupOpts, check, lib
    from wgbs_tools_b200.api import Context
    sam = open(os.path.join(tmp, "batch.sam"), "rb").read()
    ref_text = open(os.path.join(tmp, "ref.pat"), "rb").read(); ref_beta = open(os.path.join(tmp, "ref.beta"), "rb").read()
    loci = np.load(os.path.join(tmp, "loci.npy"))
    n_cpg = int(loci.size); text_bytes = len(sam)
    torch.cuda.set_device(0)
    main_stream = torch.cuda.Stream(); torch.cuda.set_stream(main_stream)
    ctx0 = Context(0, stream=main_stream.cuda_stream)
    ix = ctx0.load_index(loci, 1)
    d_sam = torch.frombuffer(bytearray(sam), dtype=torch.uint8).cuda()
    torch.cuda.synchronize()
    out = {"records": n_rec, "steps": steps, "by_streams": {}}
    for S in (1, 2, 4):
        streams = [torch.cuda.Stream() for _ in range(S)]
        ctxs = [Context(0, stream=st.cuda_stream) for st in streams]
        mcs = [torch.zeros((n_cpg, 2), dtype=torch.int32, device="cuda") for _ in range(S)]
        texts = [torch.empty(text_bytes // 4, dtype=torch.uint8, device="cuda") for _ in range(S)]
        betas = [torch.empty((n_cpg, 2), dtype=torch.uint8, device="cuda") for _ in range(S)]
        sizes = [0] * S; errs = []

        def work(w: int, k: int):
            try:
                torch.cuda.set_device(0)
                ctx = ctxs[w]
                for _ in range(k):
                    o = PileupOpts(1, 0, -1, 0, 0, 0.67, b"C")
                    h = C.c_void_p(); st = (C.c_uint64 * 8)()
                    check(lib.wgbs_pileup_sam(ctx.h, ix.h, d_sam.data_ptr(), text_bytes, C.addressof(o), C.byref(h), C.addressof(st)))
                    check(lib.wgbs_pat2beta(ctx.h, h, 1, n_cpg + 1, mcs[w].data_ptr(), 1))
                    check(lib.wgbs_collapse(ctx.h, h))
                    n = C.c_size_t()
                    check(lib.wgbs_pats_format(ctx.h, h, CHR.encode(), texts[w].data_ptr(), texts[w].numel(), C.byref(n)))
                    check(lib.wgbs_trim(ctx.h, mcs[w].data_ptr(), n_cpg, 8, betas[w].data_ptr()))
                    lib.wgbs_pats_free(ctx.h, h)
                    sizes[w] = n.value
            except Exception as e:                      # surfaces in the parent thread
                errs.append(repr(e))

        def run(total_steps: int):
            share = [total_steps // S + (1 if w < total_steps % S else 0) for w in range(S)]
            th = [threading.Thread(target=work, args=(w, share[w])) for w in range(S)]
            for t in th:
                t.start()
            for t in th:
                t.join()

        run(warmup * S)
        torch.cuda.synchronize()
        same = all(texts[w][:sizes[w]].cpu().numpy().tobytes() == ref_text and betas[w].cpu().numpy().tobytes() == ref_beta for w in range(S)) and not errs
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main_stream)
        for st in streams:
            st.wait_event(e0)
        run(steps)
        for st in streams:
            ev = torch.cuda.Event(); ev.record(st); main_stream.wait_event(ev)
        e1.record(main_stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out["by_streams"][str(S)] = {"ms_per_step": ms, "reads_per_sec": n_rec / (ms / 1e3), "identical_outputs": bool(same), "errors": errs[:3]}
        log(f"[bench] streams={S}: {ms:.3f} ms/step, {n_rec / (ms / 1e3) / 1e6:.1f} M reads/s, identical={same}")
        for c in ctxs:
            c.close()
        del mcs, texts, betas
    print(json.dumps(out), flush=True)


def segment_leg():
    """(child process of the bench) segment at a scale where the per-call structure shows: many 60 000-site chunks per call (the DP
    is one warp per chunk, so few chunks = few busy warps), K = 10 and K = 200, with the worst-case wave plan (default) and the
    exact one (WGBS_SEG_PLAN=exact, staged); per-kernel times from the library's profiler; borders compared between the plans."""
    import torch
    from wgbs_tools_b200 import synth
    from wgbs_tools_b200.api import Context
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
    ctx = Context(0, stream=stream.cuda_stream)
    out = {}
    for key, K, nch in (("K10_x64_chunks", 10, 64), ("K200_x4_chunks", 200, 4)):
        S = nch * 60_000
        betas = synth.make_betas(9, K, S)
        loci = synth.make_genome(2, "chr1", S * 110, with_bases=False).loci[:S]
        dbet = [ctx.upload(b) for b in betas]; dd = ctx.upload(loci)
        chunks = [(s, 60_000) for s in range(0, S, 60_000)]
        res = {}
        ref = None
        for plan in ("worst", "exact", "exact+redux"):
            os.environ["WGBS_SEG_PLAN"] = plan.split("+")[0]
            os.environ["WGBS_SEG_DP"] = "redux" if plan.endswith("redux") else "shuffle"
            try:
                ctx.segment(dbet, dd, chunks, 1000, 2000, 15); torch.cuda.synchronize()            # warm-up
                t0 = time.time(); r = ctx.segment(dbet, dd, chunks, 1000, 2000, 15); torch.cuda.synchronize(); sec = time.time() - t0
                ctx.prof(True); ctx.segment(dbet, dd, chunks, 1000, 2000, 15); rep = ctx.prof_report(); ctx.prof(False)
                same = ref is None or all(np.array_equal(a, b) for a, b in zip(ref, r))
                ref = ref or r
                res[plan] = {"ms": sec * 1e3, "sites_per_sec": S / sec, "blocks": int(sum(len(x) - 1 for x in r)), "same_borders_as_worst_plan": bool(same),
                             "kernels_ms": {k: round(v[1], 3) for k, v in sorted(rep.items(), key=lambda kv: -kv[1][1])[:6]},
                             "launches": {k: v[0] for k, v in rep.items() if k.startswith("seg_dp") or k.startswith("seg_cost")}}
            except Exception as e:
                res[plan] = {"error": repr(e)}
            log(f"[bench] segment {key} plan={plan}: {res[plan]}")
        out[key] = {"K": K, "sites": S, "chunks": nch, "timing": "host wall clock around the C-ABI call (betas resident in HBM, borders read back)", **res}
        for b in dbet + [dd]:
            b.free()
    print(json.dumps({"key": "segment_at_scale", **out}), flush=True)
    # ---- pat text parser: default (4 byte-wise passes) vs the two-pass tile parser (WGBS_PATPARSE=tiles, staged)
    try:
        R, N = 8_000_000, N_CPG
        txt = synth.make_pat_text_fast(3, R, N, chrom=CHR)
        d_txt = ctx.upload(txt)
        res = {"records": R, "text_bytes": len(txt)}
        ref = None
        for mode in ("default", "tiles", "tiles_tma"):
            os.environ["WGBS_PATPARSE"] = mode
            try:
                for _ in range(2):
                    ctx.pats_from_text(d_txt).free()
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                for _ in range(5):
                    ctx.pats_from_text(d_txt).free()
                b.record(stream)
                torch.cuda.synchronize()
                ms = a.elapsed_time(b) / 5
                P = ctx.pats_from_text(d_txt)
                arrs = tuple(x.tobytes() for x in P.download())
                P.free()
                same = ref is None or arrs == ref
                ref = ref or arrs
                res[mode] = {"ms": ms, "text_gb_per_s": len(txt) / ms / 1e6, "records_per_sec": R / (ms / 1e3), "same_records_as_default": bool(same)}
            except Exception as e:
                res[mode] = {"error": repr(e)}
            log(f"[bench] pat parse {mode}: {res[mode]}")
        d_txt.free()
        print(json.dumps({"key": "pat_parse", **res}), flush=True)
    except Exception as e:
        print(json.dumps({"key": "pat_parse", "error": repr(e)}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=1_000_000)
    ap.add_argument("--streams", type=int, default=1, help="batches in flight per GPU in the device-resident leg: S Contexts (own stream each) on S host "
                                                             "threads (chromosomes are independent units of work: bam2pat --gpu_streams) [1]")
    ap.add_argument("--no-extras", dest="no_extras", action="store_true", help="skip the pat2beta / homog / segment side measurements")
    ap.add_argument("--only-bam-extra", dest="only_bam", action="store_true", help="of the side measurements run only the device-BAM leg")
    ap.add_argument("--bam-leg", dest="bam_leg", help=argparse.SUPPRESS)       # internal: child process of the device-BAM leg
    ap.add_argument("--bam-configs", dest="bam_configs", default="bam_device:0::1", help=argparse.SUPPRESS)
    ap.add_argument("--segment-leg", dest="segment_leg", action="store_true", help=argparse.SUPPRESS)   # internal: child process, segment at scale
    ap.add_argument("--stream-leg", dest="stream_leg", help=argparse.SUPPRESS)  # internal: child process of the batches-in-flight leg
    ap.add_argument("--sam-bytes", dest="sam_bytes", type=int, default=0, help=argparse.SUPPRESS)
    ap.add_argument("--peak", type=float, default=6650.0, help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.bam_leg:
        return bam_device_leg(args.bam_leg, args.reads, args.sam_bytes, args.peak, args.bam_configs)
    if args.stream_leg:
        return stream_leg(args.stream_leg, args.reads)
    if args.segment_leg:
        return segment_leg()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    workload = f"bam2pat synthetic 150bp PE WGBS, {args.reads:,} records per GPU, {CHR} index ({N_CPG:,} CpGs)"
    config = {"workload": workload, "records_per_gpu": args.reads, "read_len": 150, "paired": True, "n_cpg": N_CPG,
              "sharding": "reads (one batch per GPU), beta counts NCCL-reduced" if args.gpus > 1 else "single GPU",
              "l2": "inputs larger than L2 (SAM batch ~345 MB > 126 MB)"}
    if args.streams > 1:
        config["batches_in_flight"] = args.streams

    # ------------------------------------------------------------------------------------------------------------------
    if args.impl == "reference":
        if rank != 0:
            return
        sam = make_batch(args.reads, 1000)
        n_rec = sam.count(10)
        cores = host_threads()
        shards = max(1, min(cores // 2, 64))                  # one pipeline = patter + 4 light helpers: half the cores as pipelines keeps every core busy
        r = reference_run(sam, shards, args.steps, args.warmup)
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref executables missing"}))
            return
        sec, nsh, _ = r
        val = n_rec / sec
        unit = "reads/s"
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": val, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": val, "unit": unit, "cores": cores, "kind": "reference",
                             "sample": f"whole batch ({n_rec:,} records) as {nsh} concurrent shard pipelines "
                                       "(match_maker|patter|sort|uniq|awk, reference setup.py flags); every pipeline loads the chromosome's "
                                       f"CpG dictionary first, as every chromosome worker of the reference does ({REF_DICT_LOAD_S:.1f} s of each step)"},
            "e2e": {"value": val, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    # ------------------------------------------------------------------------------------------------------------------
    g = genome()
    sam = make_batch(args.reads, 1000 + rank)
    bam_bytes = None
    if rank == 0 and args.gpus == 1 and not args.no_extras:
        # the same batch as a BAM file, for the device-decode leg (forked workers: before CUDA is initialised here)
        try:
            from wgbs_tools_b200 import bamio
            t0 = time.time()
            bam_bytes = bamio.sam_to_bam(sam, [(CHR, CHR_LEN)], procs=max(1, min(host_threads(), 32)))
            log(f"[bench] BAM of the batch: {len(bam_bytes) / 1e6:.1f} MB ({time.time() - t0:.1f}s)")
        except Exception as e:
            log(f"[bench] BAM build failed: {e!r}")
    import torch
    import torch.distributed as dist
    from wgbs_tools_b200.api import Context

    torch.cuda.set_device(local)
    if world > 1:
        import datetime
        # NCCL prints its version banner on STDOUT at communicator creation whenever NCCL_DEBUG >= VERSION; rank 0's stdout
        # must carry exactly one JSON line, so stdout points at stderr while the communicator comes up.
        sys.stdout.flush()
        keep = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=180))
            w = torch.zeros(1, device="cuda")
            dist.all_reduce(w)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(keep, 1)
            os.close(keep)
    n_rec = sam.count(10)
    text_bytes = len(sam)
    # a real (non-default) torch stream is made current and handed to the library, so that torch's CUDA events, the NCCL
    # work and every kernel of ours are ordered on ONE stream (the legacy default stream has handle 0 = "create your own")
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx = Context(local, stream=stream.cuda_stream)
    ix = ctx.load_index(g.loci, 1)
    start, end = 1, g.n_cpg + 1

    h_sam = torch.frombuffer(bytearray(sam), dtype=torch.uint8).pin_memory()
    d_sam = h_sam.cuda(non_blocking=False)
    mc = torch.zeros((g.n_cpg, 2), dtype=torch.int32, device="cuda")
    d_text = torch.empty(text_bytes // 4, dtype=torch.uint8, device="cuda")      # pat text is far smaller than the SAM text
    d_beta = torch.empty((g.n_cpg, 2), dtype=torch.uint8, device="cuda")
    h_text = torch.empty(text_bytes // 4, dtype=torch.uint8).pin_memory()
    h_beta = torch.empty((g.n_cpg, 2), dtype=torch.uint8).pin_memory()
    from wgbs_tools_b200._lib import PileupOpts, check, lib
    import ctypes as C

    last = {}

    def run_step(host: bool, reduce: bool = True, src_ptr: int | None = None):
        src = h_sam if host else d_sam
        o = PileupOpts(1, 0, -1, 0, 0, 0.67, b"C")
        h = C.c_void_p(); st = (C.c_uint64 * 8)()
        check(lib.wgbs_pileup_sam(ctx.h, ix.h, src_ptr if src_ptr is not None else src.data_ptr(), text_bytes, C.addressof(o), C.byref(h), C.addressof(st)))
        check(lib.wgbs_pat2beta(ctx.h, h, start, end, mc.data_ptr(), 1))
        work = None
        if world > 1 and reduce:                              # the one exchange step: int32[N,2] beta counts over NVLink,
            work = dist.reduce(mc, dst=0, op=dist.ReduceOp.SUM, async_op=True)   # overlapped with collapse + formatting
        check(lib.wgbs_collapse(ctx.h, h))
        n = C.c_size_t()
        tout = h_text if host else d_text
        check(lib.wgbs_pats_format(ctx.h, h, CHR.encode(), tout.data_ptr(), tout.numel(), C.byref(n)))
        if work is not None:
            work.wait()                                       # current stream waits for the NCCL stream
        bout = h_beta if host else d_beta
        if rank == 0:
            check(lib.wgbs_trim(ctx.h, mc.data_ptr(), g.n_cpg, 8, bout.data_ptr()))
        last.update(text_bytes=n.value, stats=[int(x) for x in st])
        lib.wgbs_pats_free(ctx.h, h)

    # --streams S: S - 1 more Contexts on their own streams; every worker runs whole device-resident steps (own outputs)
    workers = []
    for _ in range(max(args.streams, 1) - 1):
        st_w = torch.cuda.Stream()
        workers.append(dict(stream=st_w, ctx=Context(local, stream=st_w.cuda_stream), mc=torch.zeros_like(mc), text=torch.empty_like(d_text), beta=torch.empty_like(d_beta)))

    def worker_steps(w: dict, k: int, errs: list):
        try:
            torch.cuda.set_device(local)
            for _ in range(k):
                o = PileupOpts(1, 0, -1, 0, 0, 0.67, b"C")
                h = C.c_void_p(); st = (C.c_uint64 * 8)()
                check(lib.wgbs_pileup_sam(w["ctx"].h, ix.h, d_sam.data_ptr(), text_bytes, C.addressof(o), C.byref(h), C.addressof(st)))
                check(lib.wgbs_pat2beta(w["ctx"].h, h, start, end, w["mc"].data_ptr(), 1))
                check(lib.wgbs_collapse(w["ctx"].h, h))
                n = C.c_size_t()
                check(lib.wgbs_pats_format(w["ctx"].h, h, CHR.encode(), w["text"].data_ptr(), w["text"].numel(), C.byref(n)))
                check(lib.wgbs_trim(w["ctx"].h, w["mc"].data_ptr(), g.n_cpg, 8, w["beta"].data_ptr()))
                lib.wgbs_pats_free(w["ctx"].h, h)
        except Exception as e:
            errs.append(repr(e))

    def run_device_steps(k: int):
        """k device-resident steps: on the main Context alone, or shared out over the main Context and the workers (single-GPU runs)"""
        if not workers or world > 1:
            for _ in range(k):
                run_step(False)
            return
        S = len(workers) + 1
        share = [k // S + (1 if i < k % S else 0) for i in range(S)]
        errs: list = []
        th = [threading.Thread(target=worker_steps, args=(w, share[i + 1], errs)) for i, w in enumerate(workers)]
        for t in th:
            t.start()
        for _ in range(share[0]):
            run_step(False)
        for t in th:
            t.join()
        if errs:
            raise SystemExit(f"worker stream failed: {errs[0]}")

    d_in = [torch.empty_like(d_sam), torch.empty_like(d_sam)]     # double-buffered device copies of the streamed input

    def run_stream(k: int):
        """k end-to-end steps as a stream of batches (the way bam2pat walks chromosomes): batch i+1 is uploaded from pinned
        host memory (wgbs_prefetch, copy stream) while batch i is processed; every batch is uploaded, every result read back"""
        check(lib.wgbs_prefetch(ctx.h, d_in[0].data_ptr(), h_sam.data_ptr(), text_bytes))
        for i in range(k):
            check(lib.wgbs_prefetch_wait(ctx.h))
            if i + 1 < k:
                check(lib.wgbs_prefetch(ctx.h, d_in[(i + 1) % 2].data_ptr(), h_sam.data_ptr(), text_bytes))
            run_step(True, src_ptr=d_in[i % 2].data_ptr())

    def timed(host: bool, steps: int, warmup: int, streamed: bool = False):
        if streamed:
            run_stream(warmup)
        elif not host:
            run_device_steps(warmup * (len(workers) + 1))
        else:
            for _ in range(warmup):
                run_step(host)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        all_launches = lambda: ctx.launches + sum(w["ctx"].launches for w in workers)
        l0 = all_launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for w in workers:                                    # (no workers unless --streams > 1)
            w["stream"].wait_event(e0)
        if streamed:
            run_stream(steps)
        elif not host:
            run_device_steps(steps)
        else:
            for _ in range(steps):
                run_step(host)
        for w in workers:
            ev = torch.cuda.Event(); ev.record(w["stream"]); stream.wait_event(ev)
        e1.record(stream)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), all_launches() - l0

    # nvidia-smi samples every 100 ms; one timed region lasts tens of ms, so the sampler spans both (warm-ups included:
    # the GPU is under the same load throughout)
    cs = ClockSampler(local)
    cs.start()
    ms_dev, launches = timed(False, args.steps, args.warmup)
    ms_serial, _ = timed(True, args.steps, max(3, args.warmup))             # upload, process, read back, one batch after the other
    ms_e2e, _ = timed(True, args.steps, max(3, args.warmup), streamed=True)  # the same K batches with the upload of batch i+1 overlapped
    if d_in[0][:1 << 20].ne(d_sam[:1 << 20]).any().item() or d_in[1][-(1 << 20):].ne(d_sam[-(1 << 20):]).any().item():
        raise SystemExit("streamed input differs from the resident copy")
    if ms_dev + ms_e2e < 1500:                       # keep the GPU busy long enough for a few samples
        t_end = time.time() + 1.0
        while time.time() < t_end:                   # time-bounded => rank-local work only: NO collective in here
            run_step(False, reduce=False)
        torch.cuda.synchronize()
    clocks = cs.stop()
    nrec_t = torch.tensor([n_rec], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(nrec_t)
    total_rec = int(nrec_t.item())
    value = total_rec * args.steps / (ms_dev / 1e3)
    e2e = total_rec * args.steps / (ms_e2e / 1e3)

    # per-kernel breakdown (separate profiled steps: one event pair per launch)
    roof = None
    if rank == 0:
        ctx.prof(True)
        psteps = 3
        for _ in range(psteps):
            run_step(False, reduce=False)
        rep = ctx.prof_report()
        ctx.prof(False)
        tot = sum(v[1] for v in rep.values())
        top = sorted(rep.items(), key=lambda kv: -kv[1][1])
        log("[bench] kernel breakdown (device ms per step, share):")
        for k, (c, ms) in top[:12]:
            log(f"    {k:24s} {c // psteps:4d} launches  {ms / psteps:8.3f} ms  {100 * ms / tot:5.1f}%")
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.isfile(peaks_path):
            peak, how = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, how = 6650.0, "fallback (B200_PROFILING.md)"
        head = sam[:2_000_000].splitlines()[:5000]
        seq_end_avg = float(np.mean([len(b"\t".join(l.split(b"\t")[:10])) + 1 for l in head]))
        roof = build_roofline(rep, psteps, n_rec, text_bytes, last["stats"][7], seq_end_avg, last["text_bytes"], peak, how)

    cpu = None
    if rank == 0 and args.gpus == 1:
        cores = host_threads()
        shards = max(1, min(cores // 2, 64))                  # one pipeline = patter + 4 light helpers: half the cores as pipelines keeps every core busy
        try:
            r = reference_run(sam, shards, 1, 0)
        except Exception as e:  # the baseline must not kill the bench line
            log(f"[bench] cpu baseline failed: {e}")
            r = None
        if r:
            sec, nsh, nlines = r
            fair = None
            try:                                     # the same pipelines built with -O2 (the reference ships without -O): the "fair CPU" figure of SURVEY 8d
                load_shipped = REF_DICT_LOAD_S
                r2 = reference_run(sam, shards, 1, 0, opt=True)
                if r2:
                    fair = {"value": n_rec / r2[0], "unit": "reads/s", "flags": "-O2", "dictionary_load_s": REF_DICT_LOAD_S}
                globals()["REF_DICT_LOAD_S"] = load_shipped
            except Exception as e:
                log(f"[bench] -O2 cpu baseline failed: {e}")
            cpu = {"value": n_rec / sec, "unit": "reads/s", "cores": cores, "kind": "reference", "built_with_O2": fair,
                   "sample": f"whole batch ({n_rec:,} records) once, as {nsh} concurrent shard pipelines of the reference "
                             "executables (match_maker|patter|sort|uniq|awk; reference setup.py flags, i.e. no -O); every pipeline loads the "
                             f"chromosome's CpG dictionary first, as every chromosome worker of the reference does ({REF_DICT_LOAD_S:.1f} s of the run)",
                   "dictionary_load_s": REF_DICT_LOAD_S}

    extra = None
    if rank == 0 and args.gpus == 1 and not args.no_extras:
        try:
            extra = {} if args.only_bam else extras(ctx, torch, roof["peak"] if roof else 6650.0, sam)
        except Exception as e:
            log(f"[bench] extras failed: {e!r}")
            extra = {"error": repr(e)}
        if bam_bytes is not None:
            # in a child process with a time limit: a fault in this (newest) leg can then neither poison this process's CUDA
            # context nor hold up the bench line
            tmp = tempfile.mkdtemp(prefix="wgbsbam_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
            try:
                run_step(True)
                torch.cuda.synchronize()
                open(os.path.join(tmp, "ref.pat"), "wb").write(h_text[:last["text_bytes"]].numpy().tobytes())
                open(os.path.join(tmp, "ref.beta"), "wb").write(h_beta.numpy().tobytes())
                open(os.path.join(tmp, "batch.bam"), "wb").write(bam_bytes)
                np.save(os.path.join(tmp, "loci.npy"), g.loci)
                open(os.path.join(tmp, "batch.sam"), "wb").write(sam)
                try:                                         # S batches in flight on S streams (device-resident step)
                    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--stream-leg", tmp, "--reads", str(n_rec)], stdout=subprocess.PIPE, timeout=150)
                    line = [l for l in r.stdout.decode(errors="replace").splitlines() if l.startswith("{")]
                    extra["batches_in_flight"] = json.loads(line[-1]) if line else {"error": f"child exited {r.returncode} without a result"}
                except Exception as e:
                    log(f"[bench] batches_in_flight leg failed: {e!r}")
                    extra["batches_in_flight"] = {"error": repr(e)}
                stdout = b""
                try:                                         # segment with many chunks per call (both wave plans); pat text parsers
                    stdout = subprocess.run([sys.executable, os.path.abspath(__file__), "--segment-leg"], stdout=subprocess.PIPE, timeout=200).stdout
                except subprocess.TimeoutExpired as e:
                    stdout = e.stdout or b""
                    log("[bench] segment / pat-parse leg ran into its time limit")
                except Exception as e:
                    log(f"[bench] segment / pat-parse leg failed: {e!r}")
                for l in stdout.decode(errors="replace").splitlines():
                    if l.startswith("{"):
                        d = json.loads(l); extra[d.pop("key", "staged")] = d
                for k in ("segment_at_scale", "pat_parse"):
                    extra.setdefault(k, {"error": "no result"})
                # children, verified code first: text route (view -> SAM text -> tokenizer), direct route (BAM records feed the pileup
                # kernels in place); then ONE child for the staged configurations (two batches in flight; teams of G lanes per BGZF
                # block), which prints a line per configuration as it goes -- what it measured before a fault or its time limit is kept
                staged = os.environ.get("WGBS_BENCH_BAM_STAGED", "bam_device_direct_2streams:1::2,bam_device_direct_inflate_g8:1:g8:1,"
                                        "bam_device_direct_inflate_g16:1:g16:1,bam_device_direct_inflate_g8_2streams:1:g8:2")
                for cfgs, limit in [("bam_device:0::1", 150), ("bam_device_direct:1::1", 150)] + ([(staged, 200)] if staged else []):
                    stdout = b""
                    try:
                        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--bam-leg", tmp, "--bam-configs", cfgs, "--reads", str(n_rec), "--sam-bytes", str(text_bytes),
                                            "--peak", str(roof["peak"] if roof else 6650.0)], stdout=subprocess.PIPE, timeout=limit)
                        stdout = r.stdout
                        note = None if r.returncode == 0 else f"child exited {r.returncode}"
                    except subprocess.TimeoutExpired as e:
                        stdout, note = e.stdout or b"", f"child ran into its {limit} s limit"
                    except Exception as e:
                        note = repr(e)
                    seen = set()
                    for l in stdout.decode(errors="replace").splitlines():
                        if l.startswith("{"):
                            d = json.loads(l); seen.add(d.get("key")); extra[d.pop("key", "bam_device")] = d
                    for c in cfgs.split(","):
                        if c.split(":")[0] not in seen:
                            extra[c.split(":")[0]] = {"error": note or "no result"}
                            log(f"[bench] {c.split(':')[0]}: {note or 'no result'}")
            except Exception as e:
                log(f"[bench] bam_device leg failed: {e!r}")
                extra["bam_device"] = {"error": repr(e)}
            finally:
                subprocess.run(["rm", "-rf", tmp])

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "config": config, "clocks": clocks,
            "e2e": {"value": e2e, "unit": "reads/s", "h2d_bytes_per_step": text_bytes, "d2h_bytes_per_step": last["text_bytes"] + 2 * g.n_cpg,
                    "ms_per_step": ms_e2e / args.steps, "mode": "streamed: wgbs_prefetch uploads batch i+1 (pinned host -> HBM) while batch i is processed; all K uploads and read-backs inside the timed region",
                    "serial": {"value": total_rec * args.steps / (ms_serial / 1e3), "ms_per_step": ms_serial / args.steps,
                               "mode": "upload, process, read back one batch after the other (host pointers passed to wgbs_pileup_sam)"}},
            "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu, "extra": extra,
            "outputs": {"pat_text_bytes": last["text_bytes"], "stats": dict(zip(["lines", "pairs", "empty", "short", "invalid", "paired", "nanopore", "templates"], last["stats"]))},
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
