"""pat2beta -- `wgbstools pat2beta X.pat.gz` (reference src/python/pat2beta.py): pat text -> GPU -> uint8/uint16 beta."""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np


def pat2beta(ctx, pat_path: str, out_dir: str, nr_sites: int, lbeta: bool = False, force: bool = True, decode: str = "auto"):
    from .patio import read_pat_text, splitextgz
    suff = ".lbeta" if lbeta else ".beta"
    out_beta = os.path.join(out_dir, splitextgz(os.path.basename(pat_path)) + suff)
    if os.path.exists(out_beta) and not force:                     # delete_or_skip (utils_wgbs.py:435-454)
        print(f"File {out_beta} already exists. Skipping it. Use -f to overwrite", file=sys.stderr)
        return None
    from . import dist as wd
    rank, world = wd.world()
    dtext = None
    if world == 1 and decode in ("auto", "device"):
        from .patio import read_pat_device
        dtext = read_pat_device(ctx, pat_path)                     # BGZF inflated in HBM; None: not a BGZF file
    from .patio import pat_pieces

    def counts(text):
        """device int32[nr_sites, 2] of a pat text of any size (pieces of < 4 GiB, accumulated)"""
        mc = None
        for piece in pat_pieces(ctx, text):
            P = ctx.pats_from_text(piece)
            mc = ctx.pat2beta(P, 1, nr_sites + 1, meth_cov=mc, zero_first=mc is None)
            P.free()
        return mc

    shard = None
    if world > 1 and decode in ("auto", "device"):
        from .patio import read_pat_device_shard
        shard = read_pat_device_shard(ctx, pat_path, rank, world)   # this rank's BGZF blocks only, inflated in HBM; None: not BGZF
    if shard is not None:                                           # record-sharded: ONE reduce of the int32 counts, trim afterwards (it is non-linear)
        buf, mine = shard
        if len(mine):
            d = counts(mine); mc = d.to_host(np.int32).reshape(-1, 2); d.free()
        else:
            mc = np.zeros((nr_sites, 2), np.int32)
        buf.free()
        mc = wd.reduce_np(mc, 0)
        if rank != 0:
            return None
        beta = ctx.trim(mc, nr_sites, 16 if lbeta else 8)
        beta.tofile(out_beta)
        return out_beta
    if dtext is not None:
        mc = counts(dtext)
        beta = ctx.trim(mc, nr_sites, 16 if lbeta else 8)
        mc.free(); dtext.free()
        beta.tofile(out_beta)
        return out_beta
    text = wd.shard_lines(read_pat_text(pat_path), rank, world)
    big = len(text) > int(os.environ.get("WGBS_PAT_CHUNK_BYTES", 2 << 30))
    if world == 1 and not big:
        beta = ctx.pat2beta_text(text, 1, nr_sites + 1, 16 if lbeta else 8)   # `stdin2beta 1 N+1` + trim_to_uint8 (pat2beta.py:32-37)
    elif world == 1:
        mc = counts(text)
        beta = ctx.trim(mc, nr_sites, 16 if lbeta else 8)
        mc.free()
    else:                                                           # record-sharded: ONE reduce of the int32 counts, trim afterwards (it is non-linear)
        if big:
            d = counts(text); mc = d.to_host(np.int32).reshape(-1, 2); d.free()
        else:
            _, mc = ctx.pat2beta_text(text, 1, nr_sites + 1, 8, want_counts=True)
        mc = wd.reduce_np(mc, 0)
        if rank != 0:
            return None
        beta = ctx.trim(mc, nr_sites, 16 if lbeta else 8)
    beta.tofile(out_beta)
    return out_beta


def main(argv=None):
    from .api import Context
    from .genome import GenomeRef
    p = argparse.ArgumentParser(description="Generate a beta file from a pat file")
    p.add_argument("pat_paths", nargs="+"); p.add_argument("-f", "--force", action="store_true")
    p.add_argument("-o", "--out_dir", default="."); p.add_argument("-l", "--lbeta", action="store_true")
    p.add_argument("--genome"); p.add_argument("-@", "--threads", type=int, default=1)
    p.add_argument("--pat_decode", choices=["auto", "host", "device"], default=os.environ.get("WGBS_PAT_DECODE", "auto"),
                   help="where X.pat.gz is inflated: on the GPU when it is BGZF (compressed bytes over PCIe, one warp per block), or on the host (gzip) [auto]")
    a = p.parse_args(argv)
    ref = GenomeRef(a.genome)
    from . import dist as wd
    _, _, local = wd.init_from_env()
    with Context(local) as ctx:
        for pat in a.pat_paths:
            pat2beta(ctx, pat, a.out_dir, ref.nr_sites, a.lbeta, a.force, a.pat_decode)


if __name__ == "__main__":
    main()
