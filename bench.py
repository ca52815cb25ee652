#!/usr/bin/env python
"""bench.py -- bam2pat reads/s on BASELINE.json configs[1] (synthetic 150 bp PE WGBS reads, 1M records, chr19-sized CpG
index) on N B200s, next to the reference's own CPU pipeline; pat2beta / homog / segment sites/s beside it.

A "step" = one pass of the whole bam2pat hot path over one batch of 1M alignment records per GPU:
    [compressed BAM -> BGZF inflate -> record table -> view filters ->] pair mates -> CIGAR/CpG calls -> mate merge ->
    beta counts (+ NCCL reduce at N>1, + uint8 trim) -> sort/collapse -> pat text
`value`  : the part the reference's executables do (match_maker | patter | sort | uniq | awk): the batch's SAM text resident
           in HBM, outputs left in HBM; device timed (CUDA events, max over ranks).  `value_bam` beside it: the same from the
           COMPRESSED BAM bytes resident in HBM (inflate and record table included).
`e2e`    : the step through the public C ABI with HOST buffers: the compressed .bam bytes in pinned host memory in (what
           `wgbstools bam2pat X.bam` is given; 54 B per read over PCIe instead of 358 B of SAM text), pat text + .beta bytes in
           pinned host memory out, S batches in flight (one Context / stream / host thread each: chromosomes are independent units
           of work, the reference runs one worker per chromosome, bam2pat.py:343).  Every upload and read-back is inside the timed region.
`--impl reference`: the reference's unmodified executables (oracle/_ref) as P = host-core-count concurrent pipelines
           `match_maker | patter | sort -k2,2n -k3,3 | uniq -c | awk` over shards of the same batch cut where no template straddles.
`parity` : the GPU pat text and .beta bytes of the batch against the reference pipelines' output of the same batch.

Multi-GPU (weak scaling): every rank piles up its own 1M-record batch over the same chromosome index (the reference shards the
same way: one process per region, bam2pat.py:343); the only exchange is one NCCL reduce(sum) of the int32[nCpG,2] beta counts to
rank 0, before the non-linear uint8 trim.
"""
from __future__ import annotations

import argparse
import ctypes as C
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


def make_bam(sam: bytes) -> bytes:
    """the batch as a coordinate-sorted BGZF-compressed BAM (zlib level 6, 0xff00-byte blocks like htslib); forked workers: call it
    before CUDA is initialised in this process"""
    from wgbs_tools_b200 import bamio
    t = time.time()
    bam = bamio.sam_to_bam(sam, [(CHR, CHR_LEN)], procs=max(1, min(host_threads(), 32)))
    log(f"[bench] BAM of the batch: {len(bam) / 1e6:.1f} MB ({time.time() - t:.1f}s)")
    return bam


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def bind_to_gpu_numa_node(local: int):
    """run this rank on the cores next to its GPU (pinned buffers are then allocated on that NUMA node: at N = 8 every rank uploads
    through its own socket instead of all through one)"""
    try:
        import torch
        p = torch.cuda.get_device_properties(local)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        cl = open(f"/sys/bus/pci/devices/{bdf}/local_cpulist").read().strip()
        cpus = set()
        for part in cl.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return cl
    except Exception:
        pass
    return None


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
def safe_shards(lines: list, shards: int) -> list:
    """cut the (coordinate-sorted) lines into at most `shards` contiguous pieces at places no template straddles -- both mates of
    every pair stay in one piece, so the pieces together give exactly what ONE pipeline over the whole batch gives (match_maker
    pairs within its input only)"""
    n = len(lines)
    last = {}
    names = [l.split(b"\t", 1)[0] for l in lines]
    for i, q in enumerate(names):
        last[q] = i
    reach = np.fromiter((last[q] for q in names), dtype=np.int64, count=n)
    reach = np.maximum.accumulate(reach)                       # furthest line any template started so far extends to
    ok = np.flatnonzero(reach == np.arange(n)) + 1             # a cut AFTER line i is safe when nothing reaches past i
    cuts = [0]
    for s in range(1, shards):
        want = n * s // shards
        k = int(np.searchsorted(ok, want))
        if k < ok.size and ok[k] > cuts[-1] and ok[k] < n:
            cuts.append(int(ok[k]))
    cuts.append(n)
    return [(a, b) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]


def merge_pat_parts(parts: list) -> bytes:
    """the shard outputs as one collapsed pat text: equal (index, pattern) lines of different shards add their counts; order =
    `sort -k2,2n -k3,3` (C locale).  Test scaffolding of the comparison, not part of the timed reference step."""
    acc = {}
    for p in parts:
        for l in p.splitlines():
            c, i, pat, cnt = l.split(b"\t")
            k = (int(i), pat)
            acc[k] = acc.get(k, 0) + int(cnt)
    cb = CHR.encode()
    return b"".join(b"%s\t%d\t%s\t%d\n" % (cb, i, pat, acc[(i, pat)]) for i, pat in sorted(acc))


class ReferenceArm:
    """the reference pipelines over one batch: P shards, one `match_maker | patter | sort | uniq -c | awk` each, all concurrent"""

    def __init__(self, sam: bytes, shards: int, opt: bool):
        from oracle import harness as H
        self.H, self.opt = H, opt
        self.tmp = tempfile.mkdtemp(prefix="wgbsref_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        self.dpath = os.path.join(self.tmp, "CpG.bed")
        with open(self.dpath, "wb") as f:
            f.write(genome().dict_text())
        lines = sam.splitlines(keepends=True)
        self.n_rec = len(lines)
        self.paths = []
        for s, (a, b) in enumerate(safe_shards(lines, shards)):
            p = os.path.join(self.tmp, f"s{s}.sam")
            with open(p, "wb") as f:
                f.writelines(lines[a:b])
            self.paths.append(p)
        self.empty = os.path.join(self.tmp, "two.sam")
        with open(self.empty, "wb") as f:
            f.writelines(lines[:2])
        self.env = dict(os.environ); self.env["PATH"] = H.SHIM + os.pathsep + self.env.get("PATH", ""); self.env["LC_ALL"] = "C"
        self.cmd = (f"{H.tool('match_maker', opt)} < {{inp}} | {H.tool('patter', opt)} {self.dpath} {CHR} --min_cpg 1 --clip 0 2>/dev/null"
                    " | sort -k2,2n -k3,3 | uniq -c | awk -v OFS='\\t' '{{print $2,$3,$4,$1}}' > {out}")

    def _run(self, inputs: list) -> float:
        t0 = time.time()
        procs = [subprocess.Popen(self.cmd.format(inp=p, out=f"{self.tmp}/o{k}.pat"), shell=True, env=self.env, stderr=subprocess.DEVNULL) for k, p in enumerate(inputs)]
        rcs = [p.wait() for p in procs]
        if any(rcs):
            raise RuntimeError("reference pipeline failed")
        return time.time() - t0

    def step(self) -> float:
        """wall seconds of one pass over the whole batch"""
        return self._run(self.paths)

    def dictionary_load(self) -> float:
        """what every pipeline spends before its first read: patter loading the chromosome's CpG dictionary (patter.cpp:14-42) -- through
        the awk stand-in for `tabix` here (oracle/shim), which is slower than tabix.  The same number of concurrent pipelines on
        two-line inputs."""
        return self._run([self.empty] * len(self.paths))

    def output(self) -> bytes:
        return merge_pat_parts([open(f"{self.tmp}/o{k}.pat", "rb").read() for k in range(len(self.paths))])

    def close(self):
        subprocess.run(["rm", "-rf", self.tmp])


def reference_numbers(sam: bytes, steps: int, warmup: int, opt: bool, want_output: bool):
    """(seconds per step, seconds of dictionary loading inside it, pipelines, merged pat text or None)"""
    from oracle import harness as H
    if not H.have_ref():
        return None
    arm = ReferenceArm(sam, host_threads(), opt)
    try:
        for _ in range(warmup):
            arm.step()
        sec = float(np.mean([arm.step() for _ in range(steps)]))
        out = arm.output() if want_output else None
        load = float(np.mean([arm.dictionary_load() for _ in range(2)]))
        return sec, load, len(arm.paths), out
    finally:
        arm.close()


# ----------------------------------------------------------------------------------------------------------------------
# roofline bookkeeping: algorithmic bytes per launch of each hot kernel (DESIGN.md section "kernels")
# ----------------------------------------------------------------------------------------------------------------------
def algorithmic_bytes(kernel: str, w: dict):
    """bytes one launch must move at minimum.  w: the workload's sizes.  n for the sort kernels is not fixed
    (records when pairing, templates when collapsing): the larger is used, so the fraction is a lower bound."""
    kernel = kernel.strip("()").split("<")[0]            # the profiler reports template instances: nl_scan_k<8>
    n_rec, text_bytes, n_tmpl = w["n_rec"], w["text_bytes"], w["n_tmpl"]
    return {
        "nl_scan_k": text_bytes + 4 * n_rec,             # read the text ONCE; write one newline offset per line
        "sam_lines_k": int(w["seq_end_avg"] * n_rec) + 8 * n_rec + 49 * n_rec,   # each line up to the end of SEQ + 2 newline offsets in; 12 words + status out
        "rs_onesweep_k": 16 * n_rec, "rs_global_hist_k": 4 * n_rec,
        "pileup_measure_k": 36 * n_rec,
        # descriptors (44 B) + the CIGAR's sector + one 32-byte SEQ sector per candidate CpG (150 bp x N_CPG / CHR_LEN candidates per read)
        "pileup_call_k": int(n_rec * (44 + 32 + 32 * 150.0 * N_CPG / CHR_LEN)),
        "slots_init_k": 24 * (1 << (2 * n_rec - 1).bit_length()),
        "pair_insert_k": 12 * n_rec + 24 * n_rec,
        "pair_resolve_k": 4 * n_rec + 24 * n_rec + 32 * n_rec + 4 * n_rec,
        "merge_templates_k": 12 * n_rec + 16 * n_rec + 8 * n_rec,
        "line_write_k": 16 * n_tmpl + 8 * n_tmpl + w["out_text_bytes"],
        # BAM front end: compressed bytes in; literals (~1/16 of the output) + one 8-byte token per match out / tokens in, inflated bytes out, and read once more for the CRC
        "bgzf_team_decode_k": w.get("bam_bytes", 0) + w.get("inflated", 0) // 16 + 8 * (w.get("inflated", 0) // 16),
        "bgzf_resolve_k": 8 * (w.get("inflated", 0) // 16) + 2 * w.get("inflated", 0),
        "bam_records_k": 36 * n_rec + 57 * n_rec, "bam_pass_k": 36 * n_rec + 4 * n_rec,
    }.get(kernel)


def build_roofline(rep: dict, psteps: int, w: dict, peak: float, how: str) -> dict:
    """the `roofline` object of the bench line from the library's per-kernel profile of `psteps` steps ({kernel: (launches, ms)}):
    the dominant kernel = the one with the largest share of the step among those whose algorithmic bytes are defined"""
    tot = sum(v[1] for v in rep.values())
    top = sorted(rep.items(), key=lambda kv: -kv[1][1])
    per = []
    for k, (c, ms) in top[:12]:
        b = algorithmic_bytes(k, w)
        if b:
            per.append({"kernel": k, "launches_per_step": c // psteps, "avg_launch_ms": ms / c, "achieved": b / (ms / c / 1e3) / 1e9,
                        "frac": b / (ms / c / 1e3) / 1e9 / peak, "share_of_step": ms / tot, "algorithmic_bytes_per_launch": b})
    breakdown = {k: round(v[1] / psteps, 4) for k, v in top[:14]}
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
            "share_of_step": d["share_of_step"], "kernel_sum_ms_per_step": round(tot / psteps, 4), "per_kernel": per, "breakdown_ms_per_step": breakdown}


def hbm_peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(peaks_path):
        return float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------------------------------
# the other hot-path steps (BASELINE.json metric: "pat2beta CpG-sites/sec", homog, segment, the MM/ML pileup)
# ----------------------------------------------------------------------------------------------------------------------
def run_concurrently(cmds: list, env=None) -> float:
    t0 = time.time()
    procs = [subprocess.Popen(c, shell=True, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for c in cmds]
    for p in procs:
        p.wait()
    return time.time() - t0


def other_steps(ctx, torch, peak: float) -> dict:
    """pat2beta + homog on a sorted pat of 16M records over the chr19-sized index, segment on K = 10 betas x 480 000 sites, the MM/ML
    pileup on 150 000 tagged single-end reads: device time, roofline of the main kernel, and the reference executables on the same
    inputs as P concurrent processes (P = host cores, the reference's own `-@` default)."""
    from oracle import harness as H
    from wgbs_tools_b200 import synth
    from wgbs_tools_b200._lib import check, lib
    out = {}
    P = host_threads()
    ev = lambda: torch.cuda.Event(enable_timing=True)
    stream = torch.cuda.current_stream()
    tmp = tempfile.mkdtemp(prefix="wgbsoth_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)

    def kernel_ms(fn, reps=2):
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

    try:
        # ---- pat2beta + homog (records + symbols ~390 MB > L2)
        N, R = N_CPG, 16_000_000
        t0 = time.time()
        txt = synth.make_pat_text_fast(3, R, N, chrom=CHR)
        log(f"[bench] pat text: {R:,} records, {len(txt) / 1e6:.1f} MB ({time.time() - t0:.1f}s)")
        d_txt = ctx.upload(txt)
        mc = ctx.alloc(N * 8)
        Pt = ctx.pats_from_text(d_txt)
        sec_parse = dev_time(lambda: ctx.pats_from_text(d_txt).free())
        sec_p2b = dev_time(lambda: ctx.pat2beta(Pt, 1, N + 1, meth_cov=mc))
        words = Pt.pool_words
        p2b_bytes = R * 16 + words * 4 + N * 8          # idx,len,count,off + symbol words + one int32 pair per site written
        out["pat2beta"] = {"metric": "pat2beta_sites_per_sec", "value": N / (sec_parse + sec_p2b), "unit": "CpG-sites/s", "records_per_sec": R / (sec_parse + sec_p2b),
                           "workload": f"{R:,} pat records ({len(txt) / 1e6:.0f} MB text resident in HBM) over {N:,} sites", "parse_text_ms": sec_parse * 1e3, "accumulate_ms": sec_p2b * 1e3,
                           "kernels_ms": {"parse": kernel_ms(lambda: ctx.pats_from_text(d_txt).free()), "accumulate": kernel_ms(lambda: ctx.pat2beta(Pt, 1, N + 1, meth_cov=mc))},
                           "roofline": {"kernel": "pat2beta_k", "bound": "hbm", "achieved": p2b_bytes / sec_p2b / 1e9, "peak": peak, "unit": "GB/s",
                                        "frac": p2b_bytes / sec_p2b / 1e9 / peak, "algorithmic_bytes": p2b_bytes},
                           "parse_roofline": {"bound": "hbm", "achieved": (len(txt) + R * 16 + words * 4) / sec_parse / 1e9, "peak": peak, "unit": "GB/s",
                                              "frac": (len(txt) + R * 16 + words * 4) / sec_parse / 1e9 / peak}}
        blocks = synth.make_blocks(5, 1, N)
        rng = np.array([0, 0.334, 0.667, 1], np.float32)
        bs = ctx.upload(np.ascontiguousarray(blocks[:, 0])); be = ctx.upload(np.ascontiguousarray(blocks[:, 1]))
        d_rng = ctx.upload(rng); d_out = ctx.alloc(blocks.shape[0] * 12)
        hom = lambda: check(lib.wgbs_homog(ctx.h, Pt.h, bs.ptr, be.ptr, blocks.shape[0], d_rng.ptr, 3, 3, 0, d_out.ptr))
        sec_h = dev_time(hom)
        hom_bytes = R * 16 + words * 4 + blocks.shape[0] * 20
        out["homog"] = {"metric": "homog_sites_per_sec", "value": N / sec_h, "unit": "CpG-sites/s", "records_per_sec": R / sec_h, "blocks": int(blocks.shape[0]),
                        "workload": f"{R:,} pat records resident in HBM, {blocks.shape[0]:,} blocks", "ms": sec_h * 1e3, "kernels_ms": kernel_ms(hom),
                        "roofline": {"kernel": "homog_k", "bound": "hbm", "achieved": hom_bytes / sec_h / 1e9, "peak": peak, "unit": "GB/s", "frac": hom_bytes / sec_h / 1e9 / peak,
                                     "algorithmic_bytes": hom_bytes}}
        if H.have_ref():
            # P slices of the same text, one stdin2beta / homog process each, all at once
            cuts = [0] + [txt.index(b"\n", len(txt) * k // P) + 1 for k in range(1, P)] + [len(txt)]
            for k in range(P):
                open(f"{tmp}/p{k}.pat", "wb").write(txt[cuts[k]:cuts[k + 1]])
            bp = f"{tmp}/blocks.bed"; open(bp, "wb").write(synth.blocks_text(CHR, blocks))
            c1 = run_concurrently([f"{H.tool('stdin2beta')} 1 {N + 1} < {tmp}/p{k}.pat" for k in range(P)])
            c2 = run_concurrently([f"{H.tool('homog')} -b {bp} -r 0,0.334,0.667,1 -l 3 < {tmp}/p{k}.pat" for k in range(P)])
            out["pat2beta"]["cpu_baseline"] = {"value": N / c1, "unit": "CpG-sites/s", "records_per_sec": R / c1, "cores": P, "kind": "reference",
                                               "sample": f"the whole text as {P} slices, one stdin2beta process each, concurrently"}
            out["homog"]["cpu_baseline"] = {"value": N / c2, "unit": "CpG-sites/s", "records_per_sec": R / c2, "cores": P, "kind": "reference",
                                            "sample": f"the whole text as {P} slices, one homog process each (all blocks), concurrently"}
        Pt.free()
        for b in (d_txt, mc, bs, be, d_rng, d_out):
            b.free()
    except Exception as e:
        out["pat2beta"] = out.get("pat2beta") or {"error": repr(e)}
        out.setdefault("homog", {"error": repr(e)})
    try:
        # ---- segment: K betas x S sites in 60000-site chunks (segment.py defaults: max_cpg 1000, max_bp 2000, pcount 15)
        K, S = 10, 480_000
        betas = synth.make_betas(9, K, S)
        loci = genome().loci[:S]
        dbet = [ctx.upload(b) for b in betas]; dd = ctx.upload(loci)
        chunks = [(s, min(60_000, S - s)) for s in range(0, S, 60_000)]
        res = ctx.segment(dbet, dd, chunks, 1000, 2000, 15); torch.cuda.synchronize()
        t0 = time.time(); res = ctx.segment(dbet, dd, chunks, 1000, 2000, 15); torch.cuda.synchronize(); sec_s = time.time() - t0
        l64 = loci.astype(np.int64); e_idx = np.arange(S)
        lo_i = np.maximum(np.maximum((e_idx // 60_000) * 60_000, e_idx + 1 - 1000), np.searchsorted(l64, l64 - 2000, side="left"))
        cells = int((e_idx - lo_i + 1).sum())
        hbm_min = 2 * K * S + 8 * S
        out["segment"] = {"metric": "segment_sites_per_sec", "value": S / sec_s, "unit": "CpG-sites/s", "workload": f"K = {K} betas x {S:,} sites, {len(chunks)} chunks of 60 000 in one call",
                          "ms": sec_s * 1e3, "blocks": int(sum(len(r) - 1 for r in res)), "timing": "host wall clock around the C-ABI call (betas resident in HBM, borders read back)",
                          "cost_cells": cells, "cell_terms_per_sec": cells * K / sec_s,
                          "kernels_ms": kernel_ms(lambda: ctx.segment(dbet, dd, chunks, 1000, 2000, 15), reps=1),
                          "roofline": {"bound": "arithmetic (fp32 divide + log2f + fp64 log2 per term) and the sequential DP chain per chunk; HBM minimum shown for scale",
                                       "hbm_min_bytes": hbm_min, "achieved": hbm_min / sec_s / 1e9, "peak": peak, "unit": "GB/s", "frac": hbm_min / sec_s / 1e9 / peak}}
        if H.have_ref():
            paths = []
            for i, b in enumerate(betas):
                p = f"{tmp}/b{i}.beta"; b.tofile(p); paths.append(p)
            nproc = min(P, len(chunks))
            for k, (s, n) in enumerate(chunks[:nproc]):
                open(f"{tmp}/d{k}.txt", "wb").write(b"".join(b"%d\n" % x for x in loci[s:s + n].tolist()))
            c3 = run_concurrently([f"{H.tool('segmentor')} {' '.join(paths)} -s {s} -n {n} -max_cpg 1000 -ps 15 -max_bp 2000 < {tmp}/d{k}.txt > {tmp}/seg{k}.txt"
                                   for k, (s, n) in enumerate(chunks[:nproc])])
            same = all(np.array_equal(np.array(open(f"{tmp}/seg{k}.txt").read().split(), dtype=np.int64), np.asarray(res[k], dtype=np.int64)) for k in range(nproc))
            out["segment"]["cpu_baseline"] = {"value": sum(n for _, n in chunks[:nproc]) / c3, "unit": "CpG-sites/s", "cores": nproc, "kind": "reference",
                                              "sample": f"{nproc} chunks of 60 000 sites, one segmentor process each, concurrently", "identical_borders": bool(same)}
        for b in dbet + [dd]:
            b.free()
    except Exception as e:
        out["segment"] = {"error": repr(e)}
    try:
        # ---- the MM/ML mode of the pileup (BASELINE config 5's flavour: single-end reads with MM:Z / ML:B:C tags, 5mC + 5hmC calls)
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
        Pn.collapse(); Pn.free()
        out["pileup_mm_ml"] = {"metric": "bam2pat_mm_ml_reads_per_sec", "value": n_np / sec_np, "unit": "reads/s", "records": n_np, "ms": sec_np * 1e3, "nanopore_mode": int(st_np["nanopore"]),
                               "templates": int(st_np["templates"]), "what": "tokenize + MM/ML decode + calls + collapse, SAM text resident in HBM", "kernels_ms": kernel_ms(np_step)}
        if H.have_ref():
            sub = npsam[: npsam.index(b"\n", len(npsam) // 8) + 1]
            dp = H.write_tmp(genome().dict_text(), ".CpG.bed")
            sub2 = npsam[: npsam.index(b"\n", len(npsam) // 4) + 1]
            t0 = time.time(); ro, _ = H.ref_patter(sub, dp, CHR, False, nanopore=True); c_np = time.time() - t0
            t0 = time.time(); H.ref_patter(sub2, dp, CHR, False, nanopore=True); c_np2 = time.time() - t0
            os.remove(dp)
            gsub, _ = ctx.pileup_sam(ix_np, sub)
            gsub.collapse()
            per_read = max(c_np2 - c_np, 1e-9) / max(sub2.count(10) - sub.count(10), 1)      # two sample sizes: the dictionary load drops out
            out["pileup_mm_ml"]["cpu_baseline"] = {"value": 1.0 / per_read, "unit": "reads/s", "cores": 1, "kind": "reference",
                                                   "sample": f"patter --nanopore on {sub.count(10):,} and {sub2.count(10):,} records; rate from the difference",
                                                   "identical_pat": bool(gsub.to_text(CHR) == H.ref_collapse(ro))}
            gsub.free()
        d_np.free(); ix_np.free()
    except Exception as e:
        out["pileup_mm_ml"] = {"error": repr(e)}
    subprocess.run(["rm", "-rf", tmp])
    return out


# ----------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=1_000_000)
    ap.add_argument("--streams", type=int, default=8, help="batches in flight per GPU in the end-to-end leg: S Contexts (own stream each) on S host threads [8]")
    ap.add_argument("--no-extras", dest="no_extras", action="store_true", help="skip the pat2beta / homog / segment / MM-ML side measurements and the CPU baseline")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    workload = f"bam2pat synthetic 150bp PE WGBS, {args.reads:,} records per GPU, {CHR} index ({N_CPG:,} CpGs)"
    config = {"workload": workload, "records_per_gpu": args.reads, "read_len": 150, "paired": True, "n_cpg": N_CPG,
              "sharding": "reads (one batch per GPU), beta counts NCCL-reduced" if args.gpus > 1 else "single GPU",
              "l2": "inputs larger than L2 (SAM batch ~345 MB, inflated BAM ~275 MB > 126 MB)"}

    # ------------------------------------------------------------------------------------------------------------------
    if args.impl == "reference":
        if rank != 0:
            return
        sam = make_batch(args.reads, 1000)
        n_rec = sam.count(10)
        r = reference_numbers(sam, args.steps, args.warmup, opt=True, want_output=False)
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref executables missing"}))
            return
        sec, load, nsh, _ = r
        work = max(sec - load, 1e-9)
        val = n_rec / work
        cores = host_threads()
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": val, "unit": "reads/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": work * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": val, "unit": "reads/s", "cores": cores, "kind": "reference", "built_with_O2": True,
                             "sample": f"whole batch ({n_rec:,} records) per step as {nsh} concurrent pipelines (match_maker|patter|sort|uniq|awk built with -O2; the reference's own "
                                       "setup.py builds patter without -O), shards cut where no template straddles; the time every pipeline spends loading the CpG dictionary "
                                       f"({load:.2f} s through the awk stand-in for tabix, of {sec:.2f} s per step) is measured separately and taken out",
                             "step_s_with_dictionary_load": sec, "dictionary_load_s": load},
            "e2e": {"value": val, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    # ------------------------------------------------------------------------------------------------------------------
    g = genome()
    sam = make_batch(args.reads, 1000 + rank)
    bam = make_bam(sam)                                   # (forks worker processes: before CUDA is initialised here)
    import torch
    import torch.distributed as dist
    from wgbs_tools_b200._lib import PileupOpts, ViewOpts, check, lib
    from wgbs_tools_b200.api import Context

    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    S = max(1, args.streams)
    groups = []
    if world > 1:
        import datetime
        # NCCL prints its version banner on STDOUT at communicator creation whenever NCCL_DEBUG >= VERSION; rank 0's stdout
        # must carry exactly one JSON line, so stdout points at stderr while the communicators come up.
        sys.stdout.flush()
        keep = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=180))
            w = torch.zeros(1, device="cuda")
            dist.all_reduce(w)
            # one communicator per batch in flight: every worker thread reduces on its own, so the threads of a rank need no common order
            for _ in range(S):
                gr = dist.new_group(backend="nccl")
                dist.all_reduce(w, group=gr)
                groups.append(gr)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(keep, 1)
            os.close(keep)
    n_rec = sam.count(10)
    text_bytes = len(sam)
    n_cpg = g.n_cpg
    # real (non-default) torch streams are handed to the library, so that torch's CUDA events, the NCCL work and every kernel of
    # ours are ordered on the same streams (the legacy default stream has handle 0 = "create your own")
    streams = [torch.cuda.Stream() for _ in range(S)]
    main_stream = streams[0]
    torch.cuda.set_stream(main_stream)
    ctxs = [Context(local, stream=st.cuda_stream) for st in streams]
    ctx = ctxs[0]
    ix = ctx.load_index(g.loci, 1)

    h_sam = torch.frombuffer(bytearray(sam), dtype=torch.uint8).pin_memory()
    d_sam = h_sam.cuda(non_blocking=False)
    h_bam = torch.frombuffer(bytearray(bam), dtype=torch.uint8).pin_memory()
    d_bam = torch.zeros(len(bam) + 256, dtype=torch.uint8, device="cuda")
    d_bam[:len(bam)].copy_(h_bam)
    bix = C.c_void_p()
    check(lib.wgbs_bgzf_index_build(h_bam.data_ptr(), len(bam), C.byref(bix)))        # the file's block table (a .gzi): built once, outside every timed region
    mc = [torch.zeros((n_cpg, 2), dtype=torch.int32, device="cuda") for _ in range(S)]
    d_text = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
    d_beta = torch.empty((n_cpg, 2), dtype=torch.uint8, device="cuda")
    h_text = [torch.empty(64 << 20, dtype=torch.uint8).pin_memory() for _ in range(S)]
    h_beta = [torch.empty((n_cpg, 2), dtype=torch.uint8).pin_memory() for _ in range(S)]
    last = [{} for _ in range(S)]
    torch.cuda.synchronize()

    def finish(w: int, h, st, tout, bout, reduce: bool):
        """pat2beta (+ reduce) + collapse + pat text + trim of a piled-up batch on worker w"""
        c = ctxs[w]
        check(lib.wgbs_pat2beta(c.h, h, 1, n_cpg + 1, mc[w].data_ptr(), 1))
        work = None
        if world > 1 and reduce:                              # the one exchange step: int32[N,2] beta counts over NVLink,
            work = dist.reduce(mc[w], dst=0, op=dist.ReduceOp.SUM, group=groups[w], async_op=True)   # overlapped with collapse + formatting
        check(lib.wgbs_collapse(c.h, h))
        n = C.c_size_t()
        check(lib.wgbs_pats_format(c.h, h, CHR.encode(), tout.data_ptr(), tout.numel(), C.byref(n)))
        if work is not None:
            work.wait()                                       # this worker's stream waits for the NCCL stream
        if rank == 0:
            check(lib.wgbs_trim(c.h, mc[w].data_ptr(), n_cpg, 8, bout.data_ptr()))
        last[w].update(text_bytes=n.value, stats=[int(x) for x in st])
        lib.wgbs_pats_free(c.h, h)

    def step_sam(w: int = 0, reduce: bool = True):
        """device-resident: SAM text in HBM -> pat text + .beta in HBM"""
        o = PileupOpts(1, 0, -1, 0, 0, 0.67, b"C")
        h = C.c_void_p(); st = (C.c_uint64 * 8)()
        check(lib.wgbs_pileup_sam(ctxs[w].h, ix.h, d_sam.data_ptr(), text_bytes, C.addressof(o), C.byref(h), C.addressof(st)))
        finish(w, h, st, d_text, d_beta, reduce)

    def step_bam(w: int, host: bool, reduce: bool = True):
        """from the compressed BAM: pinned host bytes -> host outputs (e2e), or resident bytes + block table -> HBM outputs"""
        c = ctxs[w]
        B = C.c_void_p()
        if host:
            check(lib.wgbs_dbam_open(c.h, h_bam.data_ptr(), len(bam), C.byref(B)))
        else:
            check(lib.wgbs_dbam_open_indexed(c.h, d_bam.data_ptr(), len(bam), bix, C.byref(B)))
        last[w]["inflated"] = int(lib.wgbs_dbam_inflated_bytes(B))
        vo = ViewOpts(); vo.refid = 0
        o = PileupOpts(1, 0, -1, 0, 0, 0.67, b"C")
        h = C.c_void_p(); st = (C.c_uint64 * 8)()
        check(lib.wgbs_pileup_dbam(c.h, ix.h, B, C.byref(vo), C.addressof(o), C.byref(h), C.addressof(st), None))
        lib.wgbs_dbam_close(c.h, B)
        finish(w, h, st, h_text[w] if host else d_text, h_beta[w] if host else d_beta, reduce)

    def run_workers(fn, k_per_worker: int, nworkers: int):
        """fn(w) k times on each of the first nworkers workers, one host thread each"""
        if nworkers == 1:
            for _ in range(k_per_worker):
                fn(0)
            return
        errs = []

        def work(w):
            try:
                torch.cuda.set_device(local)
                torch.cuda.set_stream(streams[w])
                for _ in range(k_per_worker):
                    fn(w)
            except Exception as e:
                errs.append(repr(e))
        th = [threading.Thread(target=work, args=(w,)) for w in range(nworkers)]
        [t.start() for t in th]; [t.join() for t in th]
        if errs:
            raise SystemExit(f"worker failed: {errs[0]}")

    def timed(fn, steps: int, warmup: int, nworkers: int):
        """(device ms, host wall ms, kernel launches) of steps x nworkers passes: a start event every worker stream waits for, an end
        event that waits for every worker stream; max over ranks"""
        run_workers(fn, warmup, nworkers)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        l0 = sum(c.launches for c in ctxs)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(main_stream)
        for st in streams[1:nworkers]:
            st.wait_event(e0)
        run_workers(fn, steps, nworkers)
        for st in streams[1:nworkers]:
            evt = torch.cuda.Event(); evt.record(st); main_stream.wait_event(evt)
        e1.record(main_stream)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1), wall], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0].item()), float(t[1].item()), sum(c.launches for c in ctxs) - l0

    # nvidia-smi samples every 100 ms; one timed region lasts tens of ms, so the sampler spans all of them (warm-ups included:
    # the GPU is under the same load throughout)
    cs = ClockSampler(local)
    cs.start()
    ms_dev, _, launches = timed(lambda w: step_sam(w), args.steps, args.warmup, 1)
    ms_bam, _, _ = timed(lambda w: step_bam(w, False), args.steps, args.warmup, 1)
    ms_serial, wall_serial, _ = timed(lambda w: step_bam(w, True), args.steps, args.warmup, 1)
    # steps per worker: per * S >= K batches go through, and at least 6 per worker -- a 12-batch region lasts ~45 ms, and at N > 1 one
    # hiccup of a rank moved the whole number by 15 % (2 GPUs: 406 M reads/s with 3 steps per worker, 474 M with 10)
    per = max(6, (args.steps + S - 1) // S)
    e2e_reduce = os.environ.get("WGBS_BENCH_E2E_REDUCE", "1") != "0"    # diagnostic: 0 = the in-flight leg without its per-step reduce (how much of the N > 1 step it is)
    ms_e2e, wall_e2e, _ = timed(lambda w: step_bam(w, True, e2e_reduce), per, args.warmup, S)
    n_e2e = per * S
    if ms_dev + ms_e2e < 1500:                       # keep the GPU busy long enough for a few clock samples
        t_end = time.time() + 1.0
        while time.time() < t_end:                   # time-bounded => rank-local work only: NO collective in here
            step_sam(0, reduce=False)
        torch.cuda.synchronize()
    clocks = cs.stop()
    nrec_t = torch.tensor([n_rec], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(nrec_t)
    total_rec = int(nrec_t.item())
    value = total_rec * args.steps / (ms_dev / 1e3)
    e2e = total_rec * n_e2e / (ms_e2e / 1e3)

    # outputs of the three routes on rank 0's batch must be the same bytes
    out_text = out_beta = None
    routes_same = None
    if rank == 0:
        step_sam(0, reduce=False); torch.cuda.synchronize()
        a = (d_text[:last[0]["text_bytes"]].cpu().numpy().tobytes(), None)
        step_bam(0, False, reduce=False); torch.cuda.synchronize()
        b = (d_text[:last[0]["text_bytes"]].cpu().numpy().tobytes(), None)
        step_bam(0, True, reduce=False); torch.cuda.synchronize()
        out_text = h_text[0][:last[0]["text_bytes"]].numpy().tobytes()
        routes_same = bool(a[0] == b[0] == out_text)
        if world == 1:
            out_beta = h_beta[0].numpy().tobytes()

    # per-kernel breakdown (separate profiled steps: one event pair per launch)
    roof = roof_bam = None
    peak, how = hbm_peak()
    if rank == 0:
        head = sam[:2_000_000].splitlines()[:5000]
        w = {"n_rec": n_rec, "text_bytes": text_bytes, "n_tmpl": last[0]["stats"][7], "out_text_bytes": last[0]["text_bytes"], "bam_bytes": len(bam),
             "inflated": last[0].get("inflated", 0), "seq_end_avg": float(np.mean([len(b"\t".join(l.split(b"\t")[:10])) + 1 for l in head]))}
        for name, fn in (("sam", lambda: step_sam(0, reduce=False)), ("bam", lambda: step_bam(0, False, reduce=False))):
            ctx.prof(True)
            for _ in range(3):
                fn()
            rep = ctx.prof_report()
            ctx.prof(False)
            r = build_roofline(rep, 3, w, peak, how)
            log(f"[bench] kernel breakdown, {name} route (device ms per step): " + ", ".join(f"{k} {v}" for k, v in list(r["breakdown_ms_per_step"].items())[:10]))
            if name == "sam":
                roof = r
            else:
                roof_bam = r

    cpu = parity = None
    if rank == 0 and args.gpus == 1 and not args.no_extras:
        cores = host_threads()
        try:
            r = reference_numbers(sam, 1, 0, opt=True, want_output=True)
        except Exception as e:  # the baseline must not kill the bench line
            log(f"[bench] cpu baseline failed: {e}")
            r = None
        if r:
            sec, load, nsh, ref_text = r
            work = max(sec - load, 1e-9)
            shipped = None
            try:                                     # the reference as its own setup.py builds it (patter without -O)
                r2 = reference_numbers(sam, 1, 0, opt=False, want_output=False)
                if r2:
                    shipped = {"value": n_rec / max(r2[0] - r2[1], 1e-9), "unit": "reads/s", "step_s_with_dictionary_load": r2[0], "dictionary_load_s": r2[1]}
            except Exception as e:
                log(f"[bench] as-shipped cpu baseline failed: {e}")
            cpu = {"value": n_rec / work, "unit": "reads/s", "cores": cores, "kind": "reference", "built_with_O2": True, "as_shipped_no_O": shipped,
                   "sample": f"whole batch ({n_rec:,} records) once, as {nsh} concurrent pipelines of the reference executables (match_maker|patter|sort|uniq|awk, -O2), shards "
                             f"cut where no template straddles; dictionary loading ({load:.2f} s of {sec:.2f} s, through the awk stand-in for tabix) measured separately and taken out",
                   "step_s_with_dictionary_load": sec, "dictionary_load_s": load}
            # parity on the same batch: pat text, and the .beta bytes the reference's own stdin2beta + trim_to_uint8 arithmetic gives for it
            from oracle import harness as H
            try:
                ref_beta = H.ref_trim(H.ref_stdin2beta(ref_text, 1, n_cpg + 1)).tobytes()
                parity = {"pat_identical": bool(ref_text == out_text), "beta_identical": bool(ref_beta == out_beta), "pat_bytes": len(ref_text),
                          "against": "the reference pipelines' merged output of this batch (timed run above); .beta from the reference's stdin2beta on that text + trim_to_uint8"}
            except Exception as e:
                parity = {"error": repr(e)}

    other = None
    if rank == 0 and args.gpus == 1 and not args.no_extras:
        try:
            other = other_steps(ctx, torch, peak)
        except Exception as e:
            log(f"[bench] other steps failed: {e!r}")
            other = {"error": repr(e)}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "config": config, "clocks": clocks,
            "value_input": "SAM text of the batch resident in HBM (what match_maker | patter consume, SURVEY 8a); outputs left in HBM",
            "value_bam": {"value": total_rec * args.steps / (ms_bam / 1e3), "unit": "reads/s", "ms_per_step": ms_bam / args.steps,
                          "input": "compressed BAM bytes resident in HBM + the file's BGZF block table; inflate and record table included"},
            "e2e": {"value": e2e, "unit": "reads/s", "h2d_bytes_per_step": len(bam), "d2h_bytes_per_step": last[0]["text_bytes"] + 2 * n_cpg,
                    "ms_per_step": ms_e2e / n_e2e, "host_wall_ms_per_step": wall_e2e / n_e2e, "batches_in_flight": S, "steps_timed": n_e2e,
                    "input": "compressed .bam bytes in pinned host memory (wgbs_dbam_open + wgbs_pileup_dbam + wgbs_pat2beta + wgbs_collapse + wgbs_pats_format + wgbs_trim); "
                             "every upload and read-back inside the timed region",
                    "serial": {"value": total_rec * args.steps / (ms_serial / 1e3), "ms_per_step": ms_serial / args.steps, "host_wall_ms_per_step": wall_serial / args.steps,
                               "mode": "one batch after the other on one stream"}},
            "gpu_launches": launches, "roofline": roof, "roofline_bam_route": roof_bam, "cpu_baseline": cpu, "parity": parity, "routes_identical": routes_same,
            "numa_binding": numa,
            "outputs": {"pat_text_bytes": last[0]["text_bytes"], "stats": dict(zip(["lines", "pairs", "empty", "short", "invalid", "paired", "nanopore", "templates"], last[0]["stats"]))},
        }
        if other:
            out.update({k: v for k, v in other.items()})
        print(json.dumps(out))
    lib.wgbs_bgzf_index_free(bix)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
