"""homog -- host side of `wgbstools homog` (reference src/python/homog.py): thresholds, block ordering, output
scaling, around one wgbs_homog call per pat file."""
from __future__ import annotations

import argparse
import gzip
import os
import struct
import sys

import numpy as np


class IllegalArgumentError(ValueError):
    pass


def rate_edges(rlen: int, thresholds: str | None = None) -> tuple[str, np.ndarray]:
    """The `-r` string the reference builds (homog.py:96-104) and the float32 edges `homog` parses from it
    (istream >> float, reference patter_utils.cpp:70-81 = strtof)."""
    if thresholds:
        s = f"0,{thresholds},1"
    else:
        th1 = round(1 - (rlen - 1) / rlen, 3) + 0.001
        th2 = round((rlen - 1) / rlen, 3)
        s = f"0,{th1},{th2},1"
    edges = np.array([np.float32(float(x)) for x in s.split(",")], np.float32)   # float(x) -> nearest double -> nearest float == strtof for these short decimals
    return s, edges


def parse_range(range_str: str) -> np.ndarray:
    """homog.cpp:322-342 parse_range: monotone, within [0,1], starts at 0, ends at 1"""
    v = np.array([np.float32(float(x)) for x in range_str.split(",") if x != ""], np.float32)
    if v.size < 2 or (np.diff(v) <= 0).any() or (v < 0).any() or (v > 1).any() or v[0] > 0 or v[-1] < 1:
        raise IllegalArgumentError("Invalid range")
    return v


def trim_uxm_to_uint8(data: np.ndarray, nr_bits: int = 8) -> np.ndarray:
    """homog.py:48-58 (float64 row / rowmax * max, truncated)"""
    data = data.astype(np.int64).copy()
    dtype = np.uint16 if nr_bits == 16 else np.uint8
    max_val = 2 ** nr_bits - 1
    big = np.argwhere(data.max(axis=1) > max_val).flatten()
    data[big, :] = data[big, :] / data.max(axis=1)[big][:, None] * max_val
    return data.astype(dtype)


def blocks_are_sorted(starts, ends) -> bool:
    s = np.asarray(starts); e = np.asarray(ends)
    ds = np.diff(s)
    return bool(((ds > 0) | ((ds == 0) & (np.diff(e) >= 0))).all())


def sort_blocks_order(lines: list[bytes]) -> np.ndarray:
    """order in which `homog --sort_blocks` processes (and prints) blocks: `sort -k4,4n -k5,5n` with the whole line as
    last resort (homog.cpp:36-39), C locale."""
    keys = []
    for i, l in enumerate(lines):
        t = l.rstrip(b"\n").split(b"\t")
        keys.append((int(t[3]), int(t[4]), l.rstrip(b"\n"), i))
    return np.array([k[3] for k in sorted(keys)], np.int64)


def restore_block_order(counts_sorted: np.ndarray, starts: np.ndarray) -> np.ndarray:
    """homog.py:113-118: the wrapper maps the sorted output back with a stable argsort of startCpG only."""
    sorted_starts = np.asarray(starts).argsort(kind="stable")
    inv_order = np.argsort(sorted_starts, kind="stable")
    return counts_sorted[inv_order]


def load_blocks(path: str):
    """(lines, chr, start, end, startCpG, endCpG) of a blocks file; '#' comments and a `chr` header are skipped
    (homog.cpp:70-84)."""
    op = gzip.open if path.endswith(".gz") else open
    lines, cols = [], []
    with op(path, "rb") as f:
        for l in f:
            s = l.rstrip(b"\n")
            if not s or s.startswith(b"#"):
                continue
            t = s.split(b"\t")
            if len(t) < 5:
                raise IllegalArgumentError("Invalid block format")
            if not lines and t[0] == b"chr":
                continue
            lines.append(l if l.endswith(b"\n") else l + b"\n")
            cols.append((t[0], int(t[1]), int(t[2]), int(t[3]), int(t[4])))
    return lines, cols


def homog_counts(ctx, pats, lines, cols, edges: np.ndarray, rlen: int, inclusive: bool) -> np.ndarray:
    """int32[B, nbins] in the ORIGINAL block order (what homog.py merges next to the block columns)."""
    starts = np.array([c[3] for c in cols], np.int64); ends = np.array([c[4] for c in cols], np.int64)
    if (ends - starts <= 0).any():
        raise IllegalArgumentError("Invalid blocks file: Some blocks are empty (startCpG==endCpG)")
    need_sort = not blocks_are_sorted(starts, ends)
    order = sort_blocks_order(lines) if need_sort else np.arange(len(lines))
    blocks = np.stack([starts[order], ends[order]], axis=1).astype(np.int32)
    counts = ctx.homog(pats, blocks, edges, rlen, inclusive)
    return restore_block_order(counts, starts) if need_sort else counts


def parse_args(argv=None):
    p = argparse.ArgumentParser(description="count the number of U,X,M reads for each block for each pat file")
    p.add_argument("input_files", nargs="+"); p.add_argument("-b", "--blocks_file", required=True)
    g = p.add_mutually_exclusive_group(); g.add_argument("-o", "--out_dir"); g.add_argument("-p", "--prefix")
    p.add_argument("--force", "-f", action="store_true"); p.add_argument("--inclusive", action="store_true")
    p.add_argument("--verbose", "-v", action="store_true"); p.add_argument("--binary", action="store_true")
    p.add_argument("--genome"); p.add_argument("--nr_bits", type=int, default=8)
    p.add_argument("--thresholds", "-t"); p.add_argument("--rlen", "-l", type=int, default=3); p.add_argument("--debug", "-d", action="store_true")
    p.add_argument("--pat_decode", choices=["auto", "host", "device"], default=os.environ.get("WGBS_PAT_DECODE", "auto"),
                   help="where X.pat.gz is inflated: on the GPU when it is BGZF (one warp per block), or on the host (gzip) [auto]")
    return p.parse_args(argv)


def main(argv=None):
    from .api import Context
    from .patio import read_pat_text
    args = parse_args(argv)
    if args.nr_bits not in (8, 16):
        raise IllegalArgumentError("nr_bits must be in {8, 16}")
    if args.rlen < 2:
        raise IllegalArgumentError("rlen must be >= 2")
    if args.thresholds is not None:
        th = args.thresholds.split(",")
        if len(th) != 2 or not 1 > float(th[1]) > float(th[0]) > 0:
            raise IllegalArgumentError("Invalid thresholds")
    elif args.rlen == 2:
        raise IllegalArgumentError("for rlen==2, --thresholds must be specified")
    _, edges = rate_edges(args.rlen, args.thresholds)
    lines, cols = load_blocks(args.blocks_file)
    outdir = os.path.dirname(args.prefix) if args.prefix else (args.out_dir or ".")
    os.makedirs(outdir or ".", exist_ok=True)
    from . import dist as wd
    rank, world, local = wd.init_from_env()                            # under torchrun: records sharded over the ranks, ONE reduce of the bins
    with Context(local) as ctx:
        for pat in sorted(args.input_files):
            name = os.path.basename(pat)
            for suf in (".pat.gz", ".pat"):
                if name.endswith(suf):
                    name = name[: -len(suf)]
            prefix = args.prefix or os.path.join(outdir, name)
            opath = prefix + ".uxm" + ("" if args.binary else ".bed.gz")
            if os.path.exists(opath) and not args.force:
                print(f"[ wt homog ] skipping {name}. Use -f to overwrite", file=sys.stderr)
                continue
            dtext = None
            small = len(lines) <= 5000                                 # homog.py:107-110: few blocks -> `wgbstools cview pat -L blocks | homog`
            if small:
                # The reference then feeds homog with cview's output: the reads that start at most 100 CpGs before a block (extend_blocks.sh +
                # tabix -R) and overlap one, whole (no --strict), sorted and collapsed.  A read that starts earlier and still reaches into
                # a block is NOT counted on this path (it is with > 5000 blocks): mirrored, so both paths give the reference's numbers.
                from .genome import GenomeRef
                from .view import load_cview_blocks, view_pat
                ref = GenomeRef(args.genome)
                vtext = view_pat(ctx, ref, wd.shard_lines(read_pat_text(pat), rank, world), blocks=load_cview_blocks(args.blocks_file), prefilter=True)
            if not small and world == 1 and args.pat_decode in ("auto", "device"):
                from .patio import read_pat_device
                dtext = read_pat_device(ctx, pat)                  # BGZF: only the compressed bytes cross PCIe; None: plain gzip / text
            shard = None
            if not small and world > 1 and args.pat_decode in ("auto", "device"):
                from .patio import read_pat_device_shard
                shard = read_pat_device_shard(ctx, pat, rank, world)   # this rank's BGZF blocks only, inflated in HBM; None: not BGZF
                if shard is not None:
                    dtext = shard[1]
            from .patio import pat_pieces
            counts = None                                              # bins are sums over records: a text of any size goes piece by piece
            for piece in pat_pieces(ctx, vtext if small else dtext if dtext is not None else wd.shard_lines(read_pat_text(pat), rank, world)):
                P = ctx.pats_from_text(piece)
                c = homog_counts(ctx, P, lines, cols, edges, args.rlen, args.inclusive)
                P.free()
                counts = c if counts is None else counts + c
            if shard is not None:
                shard[0].free()
            elif dtext is not None:
                dtext.free()
            if counts is None:                                         # (a rank whose share holds no line)
                counts = np.zeros((len(lines), len(edges) - 1), np.int32)
            counts = wd.reduce_np(counts, 0)                           # bins are sums over records: exact under any record split
            if rank != 0:
                continue
            if args.binary:
                trim_uxm_to_uint8(counts, args.nr_bits).tofile(opath)
            else:
                fmt = b"\t".join([b"%d"] * counts.shape[1]) + b"\n"
                text = b"".join(b"\t".join(l.rstrip(b"\n").split(b"\t")[:5]) + b"\t" + fmt % tuple(row) for l, row in zip(lines, counts.tolist()))
                with gzip.open(opath, "wb", compresslevel=6) as f:     # (one write of the whole table; the text is what counts, not the gzip level)
                    f.write(text)


if __name__ == "__main__":
    main()
