"""segment -- host side of `wgbstools segment` (reference src/python/segment.py): chunking, batched GPU solves,
pairwise stitching.  Byte-identical blocks need the reference's exact chunk / patch schedule, so this module mirrors
break_to_chunks (:84-135), segment_process (:41-59), merge_df_list (:157-165), stitch_2_dfs (:199-232), merge2 (:243-246),
increase_patch (:249-252) -- but every round of independent DPs (all chunks; all stitching patches of one tree level and
one patch-size iteration) goes to the GPU as ONE wgbs_segment call instead of one `segmentor` process per DP.

`solve(list of (start, end)) -> list of absolute border arrays` is injected, so the same logic runs against the CUDA
library (product) or, in tests, against the oracle."""
from __future__ import annotations

import argparse
import sys

import numpy as np

DEF_CHUNK = 60000


class IllegalArgumentError(ValueError):
    pass


def break_to_chunks(regions, step: int):
    """regions: iterable of (startCpG, endCpG).  Returns (tags, starts, ends) like segment.py:124-135."""
    tags, starts, ends = [], [], []
    for start, end in regions:
        bords = list(range(start, end, step)) + [end]
        tags += [f"{start}-{end}"] * (len(bords) - 1)
        starts += bords[:-1]
        ends += bords[1:]
    return tags, starts, ends


def _is_2_overlap(b1, b2) -> bool:
    return bool(np.intersect1d(b1, b2).size)


def merge2(b1, b2):
    """segment.py:243-246"""
    nr_from_df1 = int(np.argmax(np.isin(b1, b2)))          # first border of b1 that b2 also has
    skip_from_df2 = int(np.searchsorted(b2, b1[nr_from_df1]))
    return np.concatenate([b1[:nr_from_df1 + 1], b2[skip_from_df2 + 1:]]).copy()


def increase_patch(pre_size: int, maxval: int) -> int:
    if pre_size == maxval:
        return maxval + 1
    return int(min(pre_size * 2, maxval))


def _solve_sites(solve, sites):
    """segment_process for many site ranges: ranges of a single site short-circuit (segment.py:45-46)."""
    out = [None] * len(sites)
    todo = []
    for i, (s, e) in enumerate(sites):
        assert e - s > 0, f"trying to segment an empty interval {(s, e)}"
        if e - s == 1:
            out[i] = np.array([s, e])
        else:
            todo.append(i)
    if todo:
        res = solve([sites[i] for i in todo])
        for i, r in zip(todo, res):
            out[i] = np.asarray(r)
    return out


def stitch_level(pairs, solve):
    """stitch_2_dfs (segment.py:199-232) for all pairs of one tree level; patch DPs of one size iteration are batched."""
    n = len(pairs)
    done = [None] * n
    state = []
    for b1, b2 in pairs:
        if b1[-1] != b2[0]:
            raise IllegalArgumentError("[wt segment] Patch stitching Failed! patches are not supposed to be merged")
        n1, n2 = int(b1[-1] - b1[0]), int(b2[-1] - b2[0])
        state.append([min(50, n1), min(50, n2), n1, n2])
    active = list(range(n))
    while active:
        for i in active:
            p1, p2, n1, n2 = state[i]
            if not (p1 <= n1 and p2 <= n2):
                raise IllegalArgumentError("[wt segment] Patch stitching Failed! Try increasing chunk size (--chunk_size flag)")
        sites = [(int(pairs[i][0][-1]) - state[i][0], int(pairs[i][0][-1]) + state[i][1]) for i in active]
        patches = _solve_sites(solve, sites)
        nxt = []
        for i, patch in zip(active, patches):
            b1, b2 = pairs[i]
            o1, o2 = _is_2_overlap(b1, patch), _is_2_overlap(patch, b2)
            if o1 and o2:
                done[i] = merge2(merge2(b1, patch), b2)
            else:
                if not o1:
                    state[i][0] = increase_patch(state[i][0], state[i][2])
                if not o2:
                    state[i][1] = increase_patch(state[i][1], state[i][3])
                nxt.append(i)
        active = nxt
    return done


def merge_df_list(dflist, solve):
    """segment.py:157-165"""
    while len(dflist) > 1:
        pairs = [(dflist[i - 1], dflist[i]) for i in range(1, len(dflist), 2)]
        arr = stitch_level(pairs, solve)
        last = [dflist[-1]] if len(dflist) % 2 else []
        dflist = arr + last
    return dflist[0]


def segment_regions(regions, solve, chunk_size: int = DEF_CHUNK):
    """regions: list of (startCpG, endCpG).  Returns int64[B,2] blocks (startCpG, endCpG) before the min_cpg filter,
    region by region in input order (the reference concatenates per-tag results, then sorts by startCpG in dump_result)."""
    tags, starts, ends = break_to_chunks(regions, chunk_size)
    arr = _solve_sites(solve, list(zip(starts, ends)))
    blocks = []
    seen = []
    for t in tags:
        if t not in seen:
            seen.append(t)
    for tag in seen:
        carr = [arr[i] for i in range(len(arr)) if tags[i] == tag]
        merged = merge_df_list(carr, solve)
        blocks.append(np.stack([merged[:-1], merged[1:]], axis=1))
    if not blocks:
        return np.zeros((0, 2), np.int64)
    df = np.concatenate(blocks).astype(np.int64)
    return df[np.argsort(df[:, 0], kind="stable")]


def filter_min_cpg(blocks: np.ndarray, min_cpg: int) -> np.ndarray:
    """dump_result (segment.py:175): keep blocks with endCpG - startCpG > min_cpg - 1"""
    return blocks[(blocks[:, 1] - blocks[:, 0]) > min_cpg - 1]


def effective_max_cpg(max_cpg: int, max_bp: int) -> int:
    m = min(max_cpg, max_bp // 2)          # segment.py:65
    assert m > 1
    return m


def add_loci(blocks: np.ndarray, chrom_of, loci_of) -> list[bytes]:
    """`chr \\t start \\t end \\t startCpG \\t endCpG` lines: start = locus(startCpG), end = locus(endCpG-1)+1
    (reference src/cpg2bed/add_loci.cpp:51-54)."""
    out = []
    for s, e in blocks.tolist():
        out.append(b"%s\t%d\t%d\t%d\t%d\n" % (chrom_of(s).encode(), loci_of(s), loci_of(e - 1) + 1, s, e))
    return out


class GpuSolver:
    """solve() backed by wgbs_segment: betas stay resident in HBM across all rounds."""

    def __init__(self, ctx, betas, dists, max_cpg: int, max_bp: int, pcount: float, first_site: int = 1):
        self.ctx, self.max_cpg, self.max_bp, self.pcount, self.first = ctx, max_cpg, max_bp, pcount, first_site
        self.betas = [ctx.upload(np.ascontiguousarray(b, np.uint8)) for b in betas]
        self.dists = ctx.upload(np.ascontiguousarray(dists, np.uint32))

    def __call__(self, sites):
        chunks = [(s - self.first, e - s) for s, e in sites]
        res = self.ctx.segment(self.betas, self.dists, chunks, self.max_cpg, self.max_bp, self.pcount)
        return [r + s for r, (s, _) in zip(res, sites)]

    def close(self):
        for b in self.betas:
            b.free()
        self.dists.free()


def parse_args(argv=None):
    p = argparse.ArgumentParser(description="Segment the genome, or a subset region, to homogenously methylated blocks.")
    p.add_argument("-s", "--sites"); p.add_argument("-r", "--region"); p.add_argument("-L", "--bed_file"); p.add_argument("--genome")
    g = p.add_mutually_exclusive_group(required=True)
    g.add_argument("--betas", nargs="+"); g.add_argument("--beta_file", "-F")
    p.add_argument("-c", "--chunk_size", type=int, default=DEF_CHUNK)
    p.add_argument("-p", "--pcount", type=float, default=15)
    p.add_argument("--min_cpg", type=int, default=1)
    p.add_argument("--max_cpg", type=int, default=1000)
    p.add_argument("--max_bp", type=int, default=2000)
    p.add_argument("-o", "--out_path", default=None)
    p.add_argument("-@", "--threads", type=int, default=1, help="accepted for CLI compatibility; the GPU batches all DPs")
    return p.parse_args(argv)


def main(argv=None):
    from .api import Context
    from .genome import GenomeRef, parse_region
    args = parse_args(argv)
    betas = args.betas
    if args.beta_file:
        betas = [b.strip() for b in open(args.beta_file) if b.strip() and not b.startswith("#")]
        if not betas:
            raise IllegalArgumentError(f"no beta files found in file {args.beta_file}")
    ref = GenomeRef(args.genome)
    regions = parse_region(ref, args)
    max_cpg = effective_max_cpg(args.max_cpg, args.max_bp)
    if args.chunk_size < args.max_cpg:
        print("[wt segment] WARNING: chunk_size is small compared to max_cpg and/or max_bp.", file=sys.stderr)
    n = ref.nr_sites
    arrs = []
    for b in betas:
        a = np.fromfile(b, np.uint8).reshape(-1, 2)
        if a.shape[0] != n:
            raise IllegalArgumentError(f"[wt segment] ERROR: current genome reference ({ref.name}) does not match the input beta file ({b}).")
        arrs.append(a)
    from . import dist as wd
    rank, world, local = wd.init_from_env()                            # under torchrun: the DPs of every round are dealt round-robin to the GPUs
    with Context(local) as ctx:
        solver = GpuSolver(ctx, arrs, ref.all_loci(), max_cpg, args.max_bp, args.pcount)
        blocks = filter_min_cpg(segment_regions(regions, wd.DistSolver(solver) if world > 1 else solver, args.chunk_size), args.min_cpg)
        solver.close()
    if rank != 0:
        return
    print(f"[wt segment] found {blocks.shape[0]:,} blocks", file=sys.stderr)
    lines = add_loci(blocks, ref.chrom_of_site, ref.locus_of_site)
    out = sys.stdout.buffer if args.out_path in (None, "-") else open(args.out_path, "wb")
    out.writelines(lines)


if __name__ == "__main__":
    main()
