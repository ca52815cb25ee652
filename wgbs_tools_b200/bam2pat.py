"""bam2pat -- `wgbstools bam2pat` (reference src/python/bam2pat.py): per chromosome,
    samtools view ... | [match_maker |] patter ... | sort | uniq -c | awk | bgzip     (bam2pat.py:144-209, 88-111)
becomes ONE wgbs_pileup_sam + wgbs_collapse + wgbs_pats_format per chromosome shard; parts are BGZF-compressed on host
threads and concatenated in chromosome order (bam2pat.py:398-422), and the beta file comes from the same device-resident
templates (no second pass over the pat text).

Input: a coordinate-sorted .bam (decoded by the library's own BGZF/BAM reader, csrc/bam.cu -- samtools is not needed) or
SAM text (.sam, or '-' for stdin: what `samtools view BAM` prints)."""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

MAPQ = 10
FLAGS_FILTER = 1796
FLAGS_FILTER_NANOPORE = 3844


def split_sam_by_chrom(sam: bytes) -> dict[str, bytes]:
    """contiguous runs of equal RNAME in a coordinate-sorted SAM (header lines dropped)"""
    out: dict[str, list[bytes]] = {}
    for l in sam.splitlines(keepends=True):
        if l.startswith(b"@") or not l.strip():
            continue
        c = l.split(b"\t", 3)[2].decode()
        out.setdefault(c, []).append(l)
    return {c: b"".join(v) for c, v in out.items()}


def filter_sam(sam: bytes, mapq: int, exclude_flags: int, include_flags: int | None) -> bytes:
    """samtools view -q MAPQ -F EXCL [-f INCL] on SAM text (bam2pat.py:165)"""
    keep = []
    for l in sam.splitlines(keepends=True):
        t = l.split(b"\t", 5)
        if len(t) < 5:
            keep.append(l); continue
        f, q = int(t[1]), int(t[4])
        if q < mapq or (f & exclude_flags) or (include_flags and (f & include_flags) != include_flags):
            continue
        keep.append(l)
    return b"".join(keep)


def proc_chr(ctx, ref, chrom: str, sam: bytes, args, mc_buf):
    """one chromosome: returns (pat text bytes, stats); adds the chromosome's beta counts into mc_buf (device)"""
    loci, first = ref.chrom_loci(chrom)
    ix = ctx.load_index(loci, first)
    P, st = ctx.pileup_sam(ix, sam, min_cpg=args.min_cpg, clip=args.clip, paired=-1, nanopore=args.nanopore,
                           np_thresh=args.np_thresh, cpc_call=args.cpc_call, combine_mods=args.combine_mods, mbias=args.mbias,
                           keep_names=args.long)
    if mc_buf is not None:
        ctx.pat2beta(P, 1, ref.nr_sites + 1, meth_cov=mc_buf, zero_first=False)
    P.collapse(long=args.long)                          # --long: `sort | awk '{print $1,$2,$3,1,$4}'`, no uniq (bam2pat.py:102-103)
    txt = P.to_text(chrom, long=args.long)
    P.free(); ix.free()
    pe = f"({st['pairs']:,} pairs). " if st["paired"] else ""
    good = st["lines"] - st["empty"] - st["invalid"]
    succ = int((1.0 - st["invalid"] / st["lines"]) * 100.0) if st["lines"] else 0
    short = f"{st['short']:,} with too few CpGs. " if args.min_cpg > 1 else ""
    print(f"[ patter ] [ {chrom} ] finished {st['lines']:,} lines. {pe}{good:,} good, {st['empty']:,} empty, {short}"
          f"{st['invalid']:,} invalid. (success {succ}%)", file=sys.stderr)          # patter.cpp:298-316
    return txt, st


def main(argv=None):
    from .api import Context
    from .genome import GenomeRef
    from .patio import bgzf_compress
    p = argparse.ArgumentParser(description="Run the WGBS pipeline to generate pat & beta files out of an input alignment file")
    p.add_argument("bam", nargs="+", help="coordinate-sorted .bam, SAM text (.sam), or '-' for SAM on stdin")
    p.add_argument("-s", "--sites"); p.add_argument("-r", "--region"); p.add_argument("--genome")
    p.add_argument("--include_flags", type=int); p.add_argument("-F", "--exclude_flags", type=int, default=FLAGS_FILTER)
    p.add_argument("-q", "--mapq", type=int, default=MAPQ)
    p.add_argument("--out_dir", "-o", default="."); p.add_argument("--min_cpg", type=int, default=1)
    p.add_argument("--force", "-f", action="store_true"); p.add_argument("--verbose", "-v", action="store_true")
    p.add_argument("--clip", type=int, default=0); p.add_argument("-@", "--threads", type=int, default=8)
    p.add_argument("--long", action="store_true", help="Use long format for pat file (add read name to each line)")
    p.add_argument("--no_beta", action="store_true"); p.add_argument("-l", "--lbeta", action="store_true")
    p.add_argument("--nanopore", "-np", action="store_true"); p.add_argument("--cpc_call", default="C", choices=["C", "H", "."])
    p.add_argument("--np_thresh", type=float, default=0.67); p.add_argument("--combine_mods", action="store_true")
    p.add_argument("--mbias", "-mb", action="store_true", help="write the M-bias tables (<name>.mbias/<name>.mbias.{OT,OB}.txt); plots are out of scope")
    a = p.parse_args(argv)
    if not 0 < a.np_thresh < 1:
        raise ValueError("Invalid np_thresh range: must be in range (0,1)")
    ref = GenomeRef(a.genome)
    with Context(0) as ctx:
        for path in a.bam:
            name = "stdin" if path == "-" else os.path.splitext(os.path.basename(path))[0]
            pat_path = os.path.join(a.out_dir, name + ".pat.gz")
            if os.path.exists(pat_path) and not a.force:
                print(f"File {pat_path} already exists. Skipping it. Use -f to overwrite", file=sys.stderr)
                continue
            ex = FLAGS_FILTER_NANOPORE if a.nanopore else a.exclude_flags
            bam = None
            if path.endswith(".bam"):
                from .bamio import BamFile
                bam = BamFile(path, a.threads)
                by_chrom = {c: None for c in bam.refs if bam.nrecords(c)}
            else:
                sam = sys.stdin.buffer.read() if path == "-" else open(path, "rb").read()
                by_chrom = split_sam_by_chrom(sam)
            mc = None if a.no_beta else ctx.alloc(ref.nr_sites * 8)
            if mc is not None:
                ctx.pat2beta(ctx.pats_from_text(b""), 1, ref.nr_sites + 1, meth_cov=mc, zero_first=True)
            parts = []
            mb_total = None
            for chrom in [c for c in ref.chroms if c in by_chrom]:           # chromosome_order (init_genome.py:263-275)
                if bam is not None:
                    head = bam.view(chrom, beg=1, end=1 << 29)[:4096]         # is_pair_end: FLAG of the first read (bam2pat.py:262-267)
                    first_flag = int(head.split(b"\t", 2)[1]) if head else 0
                    inc = a.include_flags if a.include_flags is not None else (3 if first_flag & 1 else None)
                    s = bam.view(chrom, 0 if a.nanopore else a.mapq, ex, inc or 0)
                else:
                    s = by_chrom[chrom]
                    first_flag = int(s.split(b"\t", 2)[1])
                    inc = a.include_flags if a.include_flags is not None else (3 if first_flag & 1 else None)
                    s = filter_sam(s, 0 if a.nanopore else a.mapq, ex, inc)
                if not s:
                    continue
                txt, st = proc_chr(ctx, ref, chrom, s, a, mc)
                if a.mbias and "mbias" in st:
                    mb_total = st["mbias"].astype(np.int64) if mb_total is None else mb_total + st["mbias"]   # mbias_merge (bam2pat.py:375-395)
                if txt:
                    parts.append(bgzf_compress(txt, a.threads))
            if not parts:
                print("[wt bam2pat] No reads found. No pat file is generated", file=sys.stderr)
                continue
            with open(pat_path, "wb") as f:
                f.write(b"".join(parts))                                       # `cat parts` (bam2pat.py:408)
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
                bp = os.path.join(a.out_dir, name + (".lbeta" if a.lbeta else ".beta"))
                beta.tofile(bp)
                print(f"[wt bam2pat] generated {bp}", file=sys.stderr)
                mc.free()


if __name__ == "__main__":
    main()
