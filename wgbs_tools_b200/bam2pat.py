"""bam2pat -- `wgbstools bam2pat` (reference src/python/bam2pat.py): per chromosome,
    samtools view ... | [match_maker |] patter ... | sort | uniq -c | awk | bgzip     (bam2pat.py:144-209, 88-111)
becomes ONE wgbs_pileup_sam + wgbs_collapse + wgbs_pats_format per chromosome shard; parts are BGZF-compressed on host
threads and concatenated in chromosome order (bam2pat.py:398-422), and the beta file comes from the same device-resident
templates (no second pass over the pat text).

Input: a coordinate-sorted .bam (decoded by the library's own BGZF/BAM readers -- on the device, csrc/bamdev.cu, or on host
threads, csrc/bam.cu; samtools is not needed) or SAM text (.sam, or '-' for stdin: what `samtools view BAM` prints)."""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

from .genome import IllegalArgumentError, extend_region
from .samfilter import filter_sam, load_bed_intervals, parse_region_str

MAPQ = 10
FLAGS_FILTER = 1796
FLAGS_FILTER_NANOPORE = 3844


def default_threads() -> int:
    """the reference's default for -@ (utils_wgbs.py:250-260): SLURM_JOB_CPUS_PER_NODE, else all CPUs, else 8 -- capped, because
    here the threads only (de)compress BGZF blocks"""
    try:
        n = int(os.environ["SLURM_JOB_CPUS_PER_NODE"]) if "SLURM_JOB_CPUS_PER_NODE" in os.environ else (len(os.sched_getaffinity(0)) or 8)
    except Exception:
        n = 8
    return max(1, min(n, 64))


def split_sam_by_chrom(sam: bytes) -> dict[str, bytes]:
    """contiguous runs of equal RNAME in a coordinate-sorted SAM (header lines dropped)"""
    out: dict[str, list[bytes]] = {}
    for l in sam.splitlines(keepends=True):
        if l.startswith(b"@") or not l.strip():
            continue
        c = l.split(b"\t", 3)[2].decode()
        out.setdefault(c, []).append(l)
    return {c: b"".join(v) for c, v in out.items()}


def template_windows(weight: int, limit: int, lo: int, hi: int):
    """None when one call takes the whole region; else 0-based half-open windows on the template key max(POS, PNEXT)
    (wgbs_view_opts.key_beg / key_end) that partition its records into ~equal shares of at most `limit` units (records of a
    .bam, bytes of SAM text): a 30x chromosome is tens of GB of SAM text, one call takes < 4 GiB.  The first window starts at
    0 and the last one is open-ended, so every record falls into exactly one window."""
    if limit <= 0 or weight <= limit or hi <= lo:
        return None
    k = -(-weight // limit)
    edges = sorted({lo + (hi - lo) * i // k for i in range(1, k)} - {lo, hi})
    bounds = [0] + [e for e in edges if e > 0] + [1 << 40]
    return list(zip(bounds[:-1], bounds[1:]))


def merge_pat_texts(ctx, chrom: str, texts: list[bytes], long: bool) -> bytes:
    """the collapsed pat texts of the windows of one chromosome -> one sorted, collapsed text (`sort -k2,2n -k3,3 | uniq -c`
    over their union: counts of identical (index, pattern) add up)"""
    if long:                                             # no uniq in --long: a merge by (index, pattern, read name)
        lines = [l for t in texts for l in t.splitlines(keepends=True)]
        lines.sort(key=lambda l: (lambda f: (int(f[1]), f[2], f[4]))(l.rstrip(b"\n").split(b"\t")))
        return b"".join(lines)
    P = ctx.pats_from_text(b"".join(texts))
    P.collapse()
    txt = P.to_text(chrom)
    P.free()
    return txt


class ChromPile:
    """one chromosome / region being piled up: every add() is one pileup call (the whole region, a template window of it, or
    the share of it found in one part of a streamed file); beta counts go into mc_buf (device) as they come, the collapsed pat
    texts are merged in finish().  patter is given the dictionary of the region extended by MAX_READ_SIZE (bam2pat.py:190):
    CpGs outside it are not called."""

    def __init__(self, ctx, ref, region: str, args, mc_buf):
        self.ctx, self.ref, self.args, self.mc_buf = ctx, ref, args, mc_buf
        self.chrom, beg, end = parse_region_str(extend_region(region))
        loci, first = ref.chrom_loci(self.chrom)
        if end > 0:
            lo = int(np.searchsorted(loci, beg, side="left")); hi = int(np.searchsorted(loci, end, side="right"))
            loci, first = loci[lo:hi], first + lo
        self.ix = ctx.load_index(loci, first)
        self.kw = dict(min_cpg=args.min_cpg, clip=args.clip, paired=-1, nanopore=args.nanopore, np_thresh=args.np_thresh, cpc_call=args.cpc_call,
                       combine_mods=args.combine_mods, mbias=args.mbias, keep_names=args.long)
        self.texts = []; self.st = None

    def add(self, s):
        """s: SAM text (bytes), or a callable(index, **pileup options) -> (Pats | None, stats) that views and piles up on the
        device (device-decoded .bam: wgbs_pileup_dbam)"""
        if not callable(s) and len(s) == 0:
            return
        args = self.args
        P, st_w = s(self.ix, **self.kw) if callable(s) else self.ctx.pileup_sam(self.ix, s, **self.kw)
        if P is None or st_w["lines"] == 0:
            if P is not None:
                P.free()
            return
        if self.mc_buf is not None:
            self.ctx.pat2beta(P, 1, self.ref.nr_sites + 1, meth_cov=self.mc_buf, zero_first=False)
        P.collapse(long=args.long)                      # --long: `sort | awk '{print $1,$2,$3,1,$4}'`, no uniq (bam2pat.py:102-103)
        self.texts.append(P.to_text(self.chrom, long=args.long))
        P.free()
        if self.st is None:
            self.st = st_w
            # patter decides paired / MM-ML mode from the FIRST line of the chromosome (patter.cpp:324-350): later calls inherit it
            self.kw.update(paired=st_w["paired"], nanopore=bool(st_w["nanopore"]))
        else:
            for k, v in st_w.items():
                if k not in ("paired", "nanopore"):
                    self.st[k] = self.st[k] + v         # counters and the M-bias tables add up

    def finish(self):
        """(pat text | None, stats)"""
        self.ix.free()
        st, args, chrom = self.st, self.args, self.chrom
        if st is None:
            return None, {"lines": 0}
        txt = self.texts[0] if len(self.texts) == 1 else merge_pat_texts(self.ctx, chrom, self.texts, args.long)
        self.texts = []
        pe = f"({st['pairs']:,} pairs). " if st["paired"] else ""
        good = st["lines"] - st["empty"] - st["invalid"]
        succ = int((1.0 - st["invalid"] / st["lines"]) * 100.0) if st["lines"] else 0
        short = f"{st['short']:,} with too few CpGs. " if args.min_cpg > 1 else ""
        print(f"[ patter ] [ {chrom} ] finished {st['lines']:,} lines. {pe}{good:,} good, {st['empty']:,} empty, {short}"
              f"{st['invalid']:,} invalid. (success {succ}%)", file=sys.stderr)          # patter.cpp:298-316
        return txt, st


def proc_chr(ctx, ref, region: str, get, args, mc_buf, windows=None):
    """one chromosome / region: returns (pat text bytes | None, stats); adds its beta counts into mc_buf (device).
    get(window): the alignments of the region [restricted to a template window] as SAM text (bytes), or a
    callable(index, **pileup options) -> (Pats | None, stats).  windows: template_windows() of the region, or None for one call."""
    pile = ChromPile(ctx, ref, region, args, mc_buf)
    for w in (windows or [None]):
        pile.add(get(w))
    return pile.finish()


class _Source:
    """the alignments of one input: a .bam (native reader) or SAM text (what `samtools view -h BAM` prints)"""

    def __init__(self, path: str, threads: int, ctx=None, decode: str = "host"):
        self.bam = None; self.sam = None; self.on_device = False; self.stream = None
        if path.endswith(".bam") and decode == "stream":
            # a file too large to hold inflated (host or device): read as a sequence of parts (bamio.stream_parts).  The head of
            # the file (header, reference list, the first records for the paired-end / MM-tag peeks) comes from its first blocks.
            from .bamio import BamPart, bgzf_block_table
            coff, csize, usize = bgzf_block_table(path)
            k = int(np.searchsorted(np.cumsum(usize), 4 << 20)) + 1
            with open(path, "rb") as f:
                self.bam = BamPart(f.read(int(coff[:k][-1] + csize[:k][-1])), threads=threads)
            self.stream = dict(path=path, threads=threads, budget=int(os.environ.get("WGBS_STREAM_BYTES", 2 << 30)),   # inflated bytes per part: its SAM text stays < 4 GiB
                               table=(coff, csize, usize), first=None)
            self.header = self.bam.header
            self.chroms = set(self.bam.refs)
            return
        if path.endswith(".bam"):
            from .bamio import BamFile, DeviceBam
            # device: the compressed bytes cross PCIe, BGZF inflate + record filtering + SAM formatting run in HBM (csrc/bamdev.cu);
            # auto: device unless the inflated file does not fit in device memory (then the host reader decodes it)
            self.on_device = decode in ("device", "auto")
            # auto: a file whose inflated bytes (~5x the compressed ones) cannot fit beside the working set goes straight to streaming
            # (under torchrun every rank would hold the whole file: the limit is shared out, and a streamed file is read by block range)
            if decode == "auto" and os.path.getsize(path) > int(os.environ.get("WGBS_DEVICE_BAM_MAX", 20 << 30)) // max(int(os.environ.get("WORLD_SIZE", "1")), 1):
                self.__init__(path, threads, ctx, "stream")
                return
            if self.on_device:
                from ._lib import WgbsError
                try:
                    self.bam = DeviceBam(ctx, path)
                except WgbsError as e:
                    if decode != "auto" or not any(m in str(e).lower() for m in ("does not fit in device memory", "out of memory", "cannot pin")):
                        raise
                    print(f"[wt bam2pat] {e}; reading the file in parts (--bam_decode stream)", file=sys.stderr)
                    self.__init__(path, threads, ctx, "stream")
                    return
            if not self.on_device:
                self.bam = BamFile(path, threads)
            self.header = self.bam.header
            self.chroms = set(self.bam.refs)                 # `samtools idxstats | cut -f1` lists every @SQ (bam2pat.py:59)
        elif path.endswith(".cram"):
            raise IllegalArgumentError("CRAM input is not supported: convert to BAM first")
        else:
            raw = sys.stdin.buffer.read() if path == "-" else open(path, "rb").read()
            self.header = b"".join(l for l in raw.splitlines(keepends=True) if l.startswith(b"@")).decode(errors="replace")
            self.sam = split_sam_by_chrom(raw)
            self.chroms = set(self.sam)
            self._first = next((l for l in raw.splitlines(keepends=True) if l.strip() and not l.startswith(b"@")), b"")

    def head(self, n: int) -> bytes:
        """samtools view BAM | head -n  (no filters)"""
        if self.bam is not None:
            return self.bam.view(None, max_records=n)
        if n == 1:
            return self._first
        return filter_sam(b"".join(self.sam.values()), max_records=n)

    def view(self, region: str, **kw):
        """SAM text of one region: bytes, or -- device decode -- a DevBuf the caller frees (len() == number of bytes)"""
        chrom, beg, end = parse_region_str(region)
        if self.bam is not None:
            if chrom not in self.bam.refs:
                return b""
            if self.on_device:
                return self.bam.view_dev(chrom, beg=beg, end=end, **kw)
            return self.bam.view(chrom, beg=beg, end=end, **kw)
        return filter_sam(self.sam.get(chrom, b""), chrom=chrom, beg=beg, end=end, **kw)

    def getter(self, region: str, **view_kw):
        """get(window) for proc_chr: the alignments of `region` that pass the `samtools view` filters, optionally restricted to a
        template window -- SAM text (bytes), or for a device-decoded .bam a callable(index, **pileup options) -> (Pats | None,
        stats) that views and piles up without leaving the device"""
        chrom, beg, end = parse_region_str(region)

        def get(window):
            if self.bam is not None:
                if chrom not in self.bam.refs:
                    return b""
                if self.on_device:
                    return lambda ix, **kw: self.bam.pileup(ix, chrom, view=dict(beg=beg, end=end, key_window=window, **view_kw), ctx=ix.ctx, **kw)
                return self.bam.view(chrom, beg=beg, end=end, key_window=window, as_array=True, **view_kw)
            return filter_sam(self.sam.get(chrom, b""), chrom=chrom, beg=beg, end=end, key_window=window, **view_kw)
        return get

    def weight(self, chrom: str) -> int:
        """how much work a chromosome is (records in a .bam, bytes of SAM text): the LPT weights of the multi-GPU split"""
        if self.stream is not None:                          # the compressed bytes between the chromosome's first block and the next one's
            if chrom not in self.bam.refs:
                return 0
            coff, csize, _ = self.stream["table"]
            lo, hi = self.chrom_blocks()[self.bam.refs.index(chrom)]
            return int(coff[hi - 1] + csize[hi - 1] - coff[lo]) if hi > lo else 0
        if self.bam is not None:
            return self.bam.nrecords(chrom) if chrom in self.bam.refs else 0
        return len(self.sam.get(chrom, b""))

    def chrom_blocks(self) -> list[tuple[int, int]]:
        """streamed .bam: the BGZF block range of every reference (bamio.chrom_block_ranges), found once by probing"""
        if self.stream["first"] is None:
            from .bamio import chrom_block_ranges
            self.stream["first"] = chrom_block_ranges(self.stream["path"], self.stream["table"], len(self.bam.refs))
        return self.stream["first"]

    def block_runs(self, chroms) -> list[tuple[int, int]] | None:
        """streamed .bam: the BGZF block ranges that hold the records of `chroms` (merged where they touch), or None when that is
        (nearly) the whole file anyway"""
        refs = self.bam.refs; nb = self.stream["table"][0].size
        want = sorted(refs.index(c) for c in chroms if c in refs)
        if len(want) >= len(refs):
            return None
        ranges = self.chrom_blocks()
        runs: list[list[int]] = []
        for c in want:
            lo, hi = ranges[c]
            if hi <= lo:
                continue
            if runs and lo <= runs[-1][1]:
                runs[-1][1] = max(runs[-1][1], hi)
            else:
                runs.append([lo, hi])
        return [(a, b) for a, b in runs]

    def close(self):
        if self.bam is not None:
            self.bam.close()


def is_sorted_header(header: str) -> bool:
    """bam2pat.py:228-234: an @HD line without 'coordinate' means the input is not coordinate sorted"""
    first = header.split("\n", 1)[0]
    return not (first.startswith("@HD") and "coordinate" not in first)


def detect_nanopore(src: "_Source") -> bool:
    """bam2pat.py:248-266: @RG PL:ONT in the header, or MM:Z:/Mm:Z: tags within the first 200 reads"""
    if "\tPL:ONT" in src.header:
        return True
    return any(b"\tMM:Z:" in l or b"\tMm:Z:" in l for l in src.head(200).splitlines())


def strand_flags(paired: bool, top: bool, bottom: bool):
    """the awk FLAG filters of bam2pat.py:135-144; both switches together leave nothing (the awk stages are chained)"""
    if top and bottom:
        return None
    if top:
        return (147, 99) if paired else (0,)
    if bottom:
        return (83, 163) if paired else (16,)
    return ()


def add_args(p):
    p.add_argument("bam", nargs="+", help="coordinate-sorted .bam, SAM text (.sam), or '-' for SAM on stdin")
    p.add_argument("-s", "--sites", help="a CpG index range, of the form: 'start-end'")
    p.add_argument("-r", "--region", help="genomic region of the form 'chr1:10,000-10,500'")
    p.add_argument("--genome", help="Genome reference name")
    p.add_argument("--include_flags", type=int, help="flags to include (samtools view -f) [3 for PE, None for SE]")
    p.add_argument("-F", "--exclude_flags", type=int, default=FLAGS_FILTER, help=f"flags to exclude (samtools view -F) [{FLAGS_FILTER}]")
    p.add_argument("-q", "--mapq", type=int, default=MAPQ, help=f"Minimal mapping quality [{MAPQ}]")
    p.add_argument("-rg", "--read_group", help="filter reads by read group (samtools view -r)")
    p.add_argument("--top_strand", action="store_true", help="Consider only top strand reads")
    p.add_argument("--bottom_strand", action="store_true", help="Consider only bottom strand reads")
    p.add_argument("--out_dir", "-o", default=".")
    p.add_argument("--min_cpg", type=int, default=1, help="Reads covering less than MIN_CPG sites are removed [1]")
    p.add_argument("--debug", "-d", action="store_true")
    p.add_argument("--force", "-f", action="store_true", help="overwrite existing files if exists")
    p.add_argument("--verbose", "-v", action="store_true")
    p.add_argument("--clip", type=int, default=0, help="Clip for each read the first and last CLIP characters [0]")
    p.add_argument("--long", action="store_true", help="Use long format for pat file (add read name to each line)")
    p.add_argument("-@", "--threads", type=int, default=default_threads(),
                   help="host threads for BGZF inflate/deflate (default: all available CPUs, capped at 64; utils_wgbs.py:250-260)")
    p.add_argument("--gpu_streams", type=int, default=1, help="chromosomes in flight on the GPU (one Context / stream / host thread each) [1]")
    p.add_argument("--bam_decode", choices=["auto", "host", "device", "stream"], default=os.environ.get("WGBS_BAM_DECODE", "auto"),
                   help="where the .bam is decoded: on the GPU (compressed bytes over PCIe, one warp per BGZF block), on host threads (zlib), "
                        "auto = GPU unless the inflated file does not fit in device memory, or stream = read the file as a sequence of "
                        "parts of WGBS_STREAM_BYTES inflated bytes (files larger than memory) [auto]")
    p.add_argument("--no_beta", action="store_true", help="Do not generate a beta file")
    p.add_argument("-l", "--lbeta", action="store_true", help="Use lbeta file (uint16) instead of beta (uint8)")
    p.add_argument("-T", "--temp_dir", help="accepted for CLI compatibility (the collapse is a device sort: no temp files)")
    lists = p.add_mutually_exclusive_group()
    lists.add_argument("--blacklist", nargs="?", const=True, default=False, help="bed file. Ignore reads overlapping this bed file")
    lists.add_argument("-L", "--whitelist", nargs="?", const=True, default=False, help="bed file. Consider only reads overlapping this bed file")
    p.add_argument("--mbias", "-mb", action="store_true", help="write the M-bias tables (<name>.mbias/<name>.mbias.{OT,OB}.txt); plots are out of scope")
    p.add_argument("--blueprint", "-bp", action="store_true", help="legacy filter: not supported")
    p.add_argument("--nanopore", "-np", action="store_true", help="Input has MM/ML modification tags. Auto-detected. Sets -q 0 and -F 3844")
    p.add_argument("--cpc_call", default="C", choices=["C", "H", "."])
    p.add_argument("--np_thresh", type=float, default=0.67)
    p.add_argument("--combine_mods", action="store_true")
    return p


def main(argv=None):
    from .api import Context
    from .genome import GenomeRef, GenomicRegion
    from .patio import bgzf_compress
    a = add_args(argparse.ArgumentParser(description="Run the WGBS pipeline to generate pat & beta files out of an input alignment file")).parse_args(argv)
    if not os.path.isdir(a.out_dir):
        raise IllegalArgumentError(f"Invalid output dir: {a.out_dir}")
    if not 0 < a.np_thresh < 1:
        raise IllegalArgumentError("Invalid np_thresh range: must be in range (0,1)")
    if a.blueprint:
        raise IllegalArgumentError("--blueprint (legacy bisulfite-conversion filter) is not supported")
    ref = GenomeRef(a.genome)
    gr = GenomicRegion(ref, region=a.region, sites=a.sites)
    # black/white lists (bam2pat.py:287-304): a bare flag means the genome's own list
    lists = None
    for flag, fname, excl in ((a.blacklist, "blacklist.bed", True), (a.whitelist, "whitelist.bed", False)):
        if flag:
            path = os.path.join(ref.dir, fname) if flag is True else flag
            if not os.path.isfile(path):
                raise IllegalArgumentError(f"Invalid file: {path}")
            lists = (load_bed_intervals(path), excl)
    empty_iv = (np.zeros(0, np.int64), np.zeros(0, np.int64))
    # under torchrun the chromosomes are dealt to the ranks (one GPU each) the way the reference deals them to its worker
    # pool (bam2pat.py:319-346); pat parts come back in chromosome order, the beta counts through ONE reduce
    from . import dist as wd
    rank, world, local = wd.init_from_env()
    with Context(local) as ctx:
        for path in a.bam:
            if rank == 0:
                print(f"[wt bam2pat] bam: {path}", file=sys.stderr)
            if path != "-" and not os.path.isfile(path):
                print(f"[wt bam2pat] Invalid bam: {path}\n[wt bam2pat] Skipping {path}", file=sys.stderr)
                continue
            name = "stdin" if path == "-" else os.path.splitext(os.path.basename(path))[0]
            pat_path = os.path.join(a.out_dir, name + (f".{a.read_group}" if a.read_group else "") + ".pat.gz")   # bam2pat.py:406-407
            if os.path.exists(pat_path) and not a.force:
                print(f"File {pat_path} already exists. Skipping it. Use -f to overwrite", file=sys.stderr)
                continue
            src = _Source(path, a.threads, ctx, a.bam_decode)
            if not is_sorted_header(src.header):
                print(f"[wt bam2pat] WARNING: based on the @HD, bam file is not sorted: {path}\n[wt bam2pat] Skipping {path}", file=sys.stderr)
                src.close(); continue
            first = src.head(1)
            if not first.strip():
                print("[wt bam2pat] Empty bam file", file=sys.stderr)
                src.close(); continue
            paired = bool(int(first.split(b"\t", 2)[1]) & 1)                    # is_pair_end (bam2pat.py:262-267)
            nanopore = a.nanopore
            if not nanopore and detect_nanopore(src):
                print("[wt bam2pat] Auto-detected modification-aware BAM — enabling --nanopore mode", file=sys.stderr)
                nanopore = True
            run = argparse.Namespace(**vars(a)); run.nanopore = nanopore
            ex = FLAGS_FILTER_NANOPORE if nanopore else a.exclude_flags
            mapq = 0 if nanopore else a.mapq
            inc = a.include_flags if a.include_flags is not None else (3 if paired else None)
            feq = strand_flags(paired, a.top_strand, a.bottom_strand)
            if gr.region_str:
                regions = [gr.region_str]
            else:
                regions = [c for c in ref.chroms if c in src.chroms]              # intersect, in chromosome_order (bam2pat.py:68-80)
                if not regions:
                    print("[wt bam2pat] Failed retrieving valid chromosome names. Perhaps you are using a wrong genome reference.", file=sys.stderr)
                    raise IllegalArgumentError("Failed")
            mine = set(range(len(regions)))
            if world > 1:
                weights = [src.weight(r.split(":")[0]) for r in regions]
                mine = set(wd.lpt_assign(weights, world)[rank])
            if a.no_beta:
                mc = None
            elif world > 1:
                import torch
                mc = torch.zeros((ref.nr_sites, 2), dtype=torch.int32, device=f"cuda:{local}")     # NCCL reduces it in place
                torch.cuda.current_stream(local).synchronize()       # the fill ran on torch's stream; the Context's stream adds into mc
            else:
                mc = ctx.alloc(ref.nr_sites * 8)
                ctx.pat2beta(ctx.pats_from_text(b""), 1, ref.nr_sites + 1, meth_cov=mc, zero_first=True)
            parts = []
            mb_total = None

            def view_kw(chrom: str) -> dict:
                kw = dict(mapq=mapq, exclude_flags=ex, include_flags=inc, read_group=a.read_group)
                if lists is not None:
                    kw.update(intervals=lists[0].get(chrom, empty_iv), exclude_intervals=lists[1])
                return kw

            if src.stream is not None and feq is not None:
                # one pass over the file, part by part; a chromosome's ChromPile lives from its first part to its last
                from .bamio import BamPart, DeviceBamPart, stream_parts
                by_chrom = {r.split(":")[0]: (ri, r) for ri, r in enumerate(regions) if ri in mine}
                piles: dict[str, ChromPile] = {}; finished: set[str] = set()
                # parts decoded on host threads, or (WGBS_STREAM_BACKEND=device) uploaded compressed and decoded in HBM
                on_dev = os.environ.get("WGBS_STREAM_BACKEND", "host") == "device"
                if on_dev:
                    opener = lambda data, refs, lens, first: DeviceBamPart(ctx, data, refs, lens, first)
                else:
                    opener = lambda data, refs, lens, first: BamPart(data, refs, lens, first, threads=src.stream["threads"])
                # only the block ranges of the chromosomes asked for are read (a region, or this rank's share of a multi-GPU run)
                runs = src.block_runs(by_chrom) if by_chrom else []
                import itertools
                stream = itertools.chain.from_iterable(
                    stream_parts(path, opener, lambda c: dict(flag_eq=feq, **view_kw(c)), src.stream["budget"], blocks=r, refs0=src.bam.refs, table=src.stream["table"])
                    for r in ([None] if runs is None else runs))

                def close_pile(chrom, mb):
                    txt, st = piles.pop(chrom).finish()
                    if txt is not None:
                        if a.mbias and "mbias" in st:
                            mb = st["mbias"].astype(np.int64) if mb is None else mb + st["mbias"]
                        if txt:
                            parts.append((by_chrom[chrom][0], bgzf_compress(txt, a.threads)))
                    return mb

                for part, chrom, win, done in stream:
                    if chrom not in by_chrom or (chrom not in piles and chrom in finished):
                        continue
                    ri, region = by_chrom[chrom]
                    _, beg, end = parse_region_str(region)
                    if chrom not in piles:
                        piles[chrom] = ChromPile(ctx, ref, region, run, mc)
                    vkw = dict(beg=beg, end=end, key_window=win, flag_eq=feq, **view_kw(chrom))
                    if on_dev:
                        piles[chrom].add(lambda ix, _p=part, _c=chrom, _v=vkw, **kw: _p.pileup(ix, _c, view=_v, **kw))
                    else:
                        piles[chrom].add(part.view(chrom, as_array=True, **vkw))
                    if done:
                        finished.add(chrom)
                        mb_total = close_pile(chrom, mb_total)
                for chrom in list(piles):                                      # (a chromosome is always closed by its last part; belt and braces)
                    mb_total = close_pile(chrom, mb_total)
                regions = []                                                   # all done above
            def do_region(cx, ri: int, region: str):
                """one chromosome / region on Context cx: (ri, compressed part | None, stats | None)"""
                chrom = region.split(":")[0]
                if feq is None:
                    return ri, None, None
                # a chromosome too large for one call (< 4 GiB of SAM text / BAM stream) is piled up in template windows
                _, rb, re_ = parse_region_str(extend_region(region)) if ":" in region else (chrom, 0, 0)
                limit = int(os.environ.get("WGBS_CHUNK_RECORDS", 6_000_000)) if src.bam is not None else int(os.environ.get("WGBS_CHUNK_BYTES", 2 << 30))
                wins = template_windows(src.weight(chrom), limit, max(rb - 1, 0) if re_ > 0 else 0, re_ if re_ > 0 else ref.chrom_length(chrom))
                if wins and a.verbose:
                    print(f"[wt bam2pat] {region}: {len(wins)} template windows", file=sys.stderr)
                txt, st = proc_chr(cx, ref, region, src.getter(region, flag_eq=feq, **view_kw(chrom)), run, mc, wins)
                if txt is None:
                    return ri, None, None
                return ri, (bgzf_compress(txt, a.threads) if txt else None), st

            todo = [(ri, region) for ri, region in enumerate(regions) if ri in mine]
            S = max(1, min(int(os.environ.get("WGBS_GPU_STREAMS", a.gpu_streams)), len(todo) or 1))
            if S == 1:
                results = (do_region(ctx, ri, region) for ri, region in todo)
            else:
                # S chromosomes in flight on this GPU: one Context (own stream, own scratch) per worker thread, like the reference's
                # pool of chromosome workers (bam2pat.py:343); the beta counts of all of them add into the one device array
                import threading
                from concurrent.futures import ThreadPoolExecutor
                ctx.sync()                                                      # the zeroed counters are there before any worker adds to them
                tls = threading.local(); made = []

                def on_worker(item):
                    if not hasattr(tls, "ctx"):
                        tls.ctx = Context(local); made.append(tls.ctx)
                    return do_region(tls.ctx, *item)
                with ThreadPoolExecutor(S) as pool:
                    results = list(pool.map(on_worker, todo))
                for c in made:
                    c.sync(); c.close()
            for ri, part, st in results:
                if st is None:
                    if a.verbose:
                        print(f"[wt bam2pat] Skipping region {regions[ri]}, no reads found", file=sys.stderr)
                    continue
                if a.mbias and "mbias" in st:
                    mb_total = st["mbias"].astype(np.int64) if mb_total is None else mb_total + st["mbias"]   # mbias_merge (bam2pat.py:375-395)
                if part:
                    parts.append((ri, part))
            src.close()
            if world > 1:
                ctx.sync()
                if mc is not None:
                    wd.reduce_counts(mc, 0)
                if a.mbias:
                    mb_total = wd.reduce_np(np.zeros((2, 2, 1000, 2), np.int64) if mb_total is None else mb_total, 0)
                parts = [p for lst in wd.gather_parts(parts, 0) or [] for p in lst]
                if rank != 0:
                    continue
            parts = [b for _, b in sorted(parts, key=lambda t: t[0])]
            if not parts:
                print("[wt bam2pat] No reads found. No pat file is generated", file=sys.stderr)
                if mc is not None and hasattr(mc, "free"):
                    mc.free()
                continue
            with open(pat_path, "wb") as f:
                f.write(b"".join(parts))                                       # `cat parts` (bam2pat.py:408)
            from .csi import index_pat
            index_pat(pat_path)                                                # Indxer(pat_path).run(): X.pat.gz.csi (bam2pat.py:415)
            print(f"[wt bam2pat] generated {pat_path}", file=sys.stderr)
            if a.mbias and mb_total is not None:
                mdir = os.path.join(a.out_dir, name) + ".mbias"
                os.makedirs(mdir, exist_ok=True)
                for si, x in enumerate(("OT", "OB")):
                    with open(os.path.join(mdir, name) + f".mbias.{x}.txt", "w") as f:
                        f.write("r1m1\tr1u1\tr2m2\tr2u2\n")
                        for pos in range(1000):
                            t = mb_total[si, :, pos, :]
                            f.write(f"{t[0, 0]}\t{t[0, 1]}\t{t[1, 0]}\t{t[1, 1]}\n")
            if mc is not None:
                beta = ctx.trim(mc, ref.nr_sites, 16 if a.lbeta else 8)
                bp = pat_path[:-len(".pat.gz")] + (".lbeta" if a.lbeta else ".beta")
                beta.tofile(bp)
                print(f"[wt bam2pat] generated {bp}", file=sys.stderr)
                if hasattr(mc, "free"):
                    mc.free()


if __name__ == "__main__":
    main()
