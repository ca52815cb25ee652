"""beta_to_blocks -- `wgbstools beta_to_blocks` (reference src/python/beta_to_blocks.py): collapse beta files to blocks:
per block the sums of (meth, cover) over its sites (np.add.reduceat), trimmed to uint8 (`.bin`) or uint16 (`.lbeta`),
optionally a bedGraph.  One wgbs_beta_to_blocks call per beta file."""
from __future__ import annotations

import argparse
import gzip
import os
import sys

import numpy as np

from .genome import IllegalArgumentError


def load_blocks(path: str):
    """(chr, start, end) text columns and int64 startCpG/endCpG (NA -> -1) of a 5+ column blocks file; `#` comments and
    a header row are skipped (beta_to_blocks.py:50-91)"""
    op = gzip.open if path.endswith(".gz") else open
    head, s, e = [], [], []
    with op(path, "rt") as f:
        for l in f:
            if not l.strip() or l.startswith("#"):
                continue
            t = l.rstrip("\r\n").split("\t")
            if len(t) < 5:
                raise IllegalArgumentError(f"Invalid blocks file: {path}. less than 5 columns.\n"
                                           f"Run wgbstools convert -L {path} -o OUTPUT_REGION_FILE to add the CpG columns")
            if not head and not t[1].isdigit():
                continue                                            # header row
            head.append(tuple(t[:3]))
            na = t[3] in ("NA", "NaN", "") or t[4] in ("NA", "NaN", "")
            s.append(-1 if na else int(t[3])); e.append(-1 if na else int(t[4]))
    s = np.array(s, np.int64); e = np.array(e, np.int64)
    ok = s >= 0
    if not np.all(e[ok] - s[ok] >= 0):
        raise IllegalArgumentError(f"Invalid CpG columns in blocks file {path}")
    return head, s, e


def is_block_file_nice(s: np.ndarray, e: np.ndarray) -> tuple[bool, str]:
    """beta_to_blocks.py:23-47"""
    if np.any(s < 0):
        return False, "Some blocks are empty (NA)"
    if not np.all(e - s > 0):
        return False, "Some blocks are empty (startCpG==endCpG)"
    if not np.all(np.diff(s) >= 0):
        return False, "startCpG is not monotonically increasing"
    if not np.all(np.diff(e) >= 0):
        return False, "endCpG is not monotonically increasing"
    if np.unique(np.stack([s, e], 1), axis=0).shape[0] != s.size:
        return False, "Some blocks are duplicated"
    if not np.all(s[1:] - e[:-1] >= 0):
        return False, "Some blocks overlap"
    return True, ""


def main(argv=None):
    from .api import Context
    p = argparse.ArgumentParser(description="Collapse beta file to blocks binary file, of the same beta format")
    p.add_argument("input_files", nargs="+", help="one or more beta files")
    p.add_argument("-b", "--blocks_file", required=True); p.add_argument("-o", "--out_dir", default=".")
    p.add_argument("-l", "--lbeta", action="store_true", help="Use lbeta file (uint16) instead of bin (uint8)")
    p.add_argument("--bedGraph", action="store_true", help="output a text file in addition to binary file")
    p.add_argument("--force", "-f", action="store_true"); p.add_argument("--debug", "-d", action="store_true")
    p.add_argument("-@", "--threads", type=int, default=1, help="accepted for CLI compatibility")
    a = p.parse_args(argv)
    if not os.path.isdir(a.out_dir):
        raise IllegalArgumentError(f"Invalid output dir: {a.out_dir}")
    head, s, e = load_blocks(a.blocks_file)
    nice, msg = is_block_file_nice(s, e)
    if not nice:
        print("[ wt beta_to_blocks ]", msg, file=sys.stderr)
    bs = np.where(s < 0, 1, s); be = np.where(s < 0, 1, e)          # NA blocks sum to (0, 0) (slow_method)
    suff = ".lbeta" if a.lbeta else ".bin"
    with Context(0) as ctx:
        for beta in a.input_files:
            name = os.path.splitext(os.path.basename(beta))[0]
            prefix = os.path.join(a.out_dir, name)
            if os.path.isfile(prefix + suff) and not a.force:
                print(f"[ wt beta_to_blocks ] Skipping {beta}. Use -f flag to overwrite", file=sys.stderr)
                continue
            ext = os.path.splitext(beta)[1]
            if ext not in (".beta", ".lbeta", ".bin") or not os.path.isfile(beta):
                raise IllegalArgumentError(f"Invalid beta file:\n{beta}")
            data = np.fromfile(beta, np.uint16 if ext == ".lbeta" else np.uint8).reshape(-1, 2)
            table, sums = ctx.beta_to_blocks(data, bs, be, 16 if a.lbeta else 8, want_sums=True)
            table.tofile(prefix + suff)
            print("[ wt beta_to_blocks ]", prefix + suff, file=sys.stderr)
            if a.bedGraph:                                          # chr start end beta(%.2f, -1 if no coverage) coverage
                with open(prefix + ".bedGraph", "w") as f:
                    for (c, st, en), (m, v) in zip(head, sums.tolist()):
                        f.write(f"{c}\t{st}\t{en}\t{'-1' if v == 0 else '%.2f' % (m / v)}\t{v}\n")


if __name__ == "__main__":
    main()
