"""beta_to_table -- `wgbstools beta_to_table` (reference src/python/beta_to_table.py:72-107 get_table, :110-165): one row per
block, one column per sample (or per group of samples, averaged): meth / cover over the block's sites, NA where the block has fewer
than --min_cov observations.  The per-block sums are wgbs_beta_to_blocks on the device (np.add.reduceat in the reference,
beta_to_blocks.py:101-126); division, group means and the text are host formatting."""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

from .beta_to_blocks import load_blocks
from .genome import IllegalArgumentError


def load_groups(groups_file: str | None, betas: list[str]) -> list[tuple[str, str, str]]:
    """(fname, group, full path) rows in file order (beta_to_table.py:39-57, dmb.py:24-38,41-75): the first csv column names the
    sample, `group` its group, rows with `include` == False are dropped; without a file every beta is its own group"""
    rows = []
    if groups_file is not None:
        if not os.path.isfile(groups_file):
            raise IllegalArgumentError(f"Invalid file: {groups_file}")
        import csv
        with open(groups_file, newline="") as f:
            lines = [l for l in f if not l.startswith("#")]
        r = list(csv.reader(lines))
        head = r[0]
        if "group" not in head:
            raise IllegalArgumentError('gropus file must have a column named "group"')
        gi = head.index("group"); ii = head.index("include") if "include" in head else None
        for t in r[1:]:
            if not t or len(t) <= gi:
                continue
            if ii is not None:
                if t[ii] not in ("True", "False"):
                    raise IllegalArgumentError("Invalid group file. Include column must be boolean")
                if t[ii] == "False":
                    continue
            if t[0] == "" or t[gi] == "":
                continue                                            # dropna
            rows.append((t[0], t[gi]))
    else:
        seen = set()
        for b in betas:
            if b not in seen:
                seen.add(b)
                name = os.path.basename(b)
                for suf in (".beta", ".lbeta", ".bin"):
                    if name.endswith(suf):
                        name = name[: -len(suf)]
                rows.append((name, name))
    suff = ".lbeta" if betas[0].endswith(".lbeta") else ".beta"
    out, missing = [], []
    for fname, group in rows:
        hits = [b for b in betas if os.path.basename(b) == fname + suff]
        if not hits:
            missing.append(fname)
        else:
            out.append((fname, group, hits[0]))
    if missing:
        raise IllegalArgumentError(f"{len(missing)} prefixes from groups file were not found in input bins: {missing[:5]}")
    return out


def beta2vec(sums: np.ndarray, min_cov: int) -> np.ndarray:
    """utils_wgbs.py:270-274: meth / cover where cover >= min_cov, NaN elsewhere"""
    cond = sums[:, 1] >= min_cov
    vec = np.full(sums.shape[0], np.nan)
    np.divide(sums[:, 0], sums[:, 1], where=cond, out=vec)
    return vec


def table_text(ctx, blocks_path: str, betas: list[str], groups_file: str | None, min_cov: int, digits: int, chunk_size: int | None = None) -> str:
    head, s, e = load_blocks(blocks_path)
    rows = load_groups(groups_file, betas)
    bs = np.where(s < 0, 1, s); be = np.where(s < 0, 1, e)          # NA blocks sum to (0, 0)
    vecs = {}
    for fname, _, path in rows:
        if path in vecs:
            continue
        ext = os.path.splitext(path)[1]
        data = np.fromfile(path, np.uint16 if ext == ".lbeta" else np.uint8).reshape(-1, 2)
        _, sums = ctx.beta_to_blocks(data, bs, be, 8, want_sums=True)
        vecs[path] = beta2vec(sums.astype(np.float64), min_cov)
    groups = []
    for _, g, _ in rows:
        if g not in groups:
            groups.append(g)
    import warnings
    cols = []
    with warnings.catch_warnings():
        warnings.filterwarnings("ignore", category=RuntimeWarning)
        for g in groups:
            cols.append(np.nanmean(np.stack([vecs[p] for _, gg, p in rows if gg == g]), axis=0))
    fmt = f"%.{digits}f"
    all_int = bool(np.all(s >= 0))
    out = ["\t".join(["chr", "start", "end", "startCpG", "endCpG"] + groups) + "\n"]
    n = len(head)
    step = chunk_size or n or 1
    for a in range(0, n, step):                                     # the reference prints chunk by chunk (header once): same text
        for i in range(a, min(n, a + step)):
            c, st, en = head[i]
            cp = [str(int(s[i])) if s[i] >= 0 else "NA", str(int(e[i])) if s[i] >= 0 else "NA"]
            vals = ["NA" if np.isnan(col[i]) else fmt % col[i] for col in cols]
            out.append("\t".join([c, st, en] + cp + vals) + "\n")
    _ = all_int
    return "".join(out)


def main(argv=None):
    from .api import Context
    p = argparse.ArgumentParser(description="build a text table from beta files; optionally collapse samples with a groups file")
    p.add_argument("blocks", help="Blocks file with no header and with >= 5 columns")
    p.add_argument("--output", "-o", help="specify output path for the table [Default is stdout]")
    p.add_argument("--groups_file", "-g", help="groups csv file with at least 2 columns: name, group. beta files belong to the same group are averaged")
    p.add_argument("--betas", nargs="+", required=True, help="beta files")
    p.add_argument("--verbose", "-v", action="store_true")
    p.add_argument("-c", "--min_cov", type=int, default=4, help="blocks with less than MIN_COV site observations are considered as missing. [4]")
    p.add_argument("--digits", type=int, default=2, help="float percision (number of digits) [2]")
    p.add_argument("--chunk_size", type=int, default=200000, help="Number of blocks to load on each step [200000]")
    p.add_argument("-@", "--threads", type=int, default=1, help="accepted for CLI compatibility")
    a = p.parse_args(argv)
    if not os.path.isfile(a.blocks):
        raise IllegalArgumentError(f"Invalid file: {a.blocks}")
    with Context(0) as ctx:
        txt = table_text(ctx, a.blocks, a.betas, a.groups_file, a.min_cov, a.digits, a.chunk_size)
    if a.output is None:
        sys.stdout.write(txt)
    else:
        with open(a.output, "w") as f:
            f.write(txt)


if __name__ == "__main__":
    main()
