"""index -- `wgbstools index` for pat files (reference src/python/index.py:83-139): BGZF-compress a plain .pat if needed and
write the CSI index next to the .pat.gz (`tabix -Cf -b 2 -e 2 -m 12`).  Host-side format work; see csi.py."""
from __future__ import annotations

import argparse
import os
import sys

from .genome import IllegalArgumentError


def index_file(path: str, force: bool = False, threads: int = 8) -> str:
    from .csi import index_pat
    from .patio import bgzf_compress
    if path.endswith(".pat"):
        gz = path + ".gz"
        if os.path.exists(gz) and not force:
            raise IllegalArgumentError(f"File {gz} already exists. Use -f to overwrite")
        with open(path, "rb") as f, open(gz, "wb") as o:
            o.write(bgzf_compress(f.read(), threads))
        path = gz
    if not path.endswith(".pat.gz"):
        raise IllegalArgumentError(f"Unknown input format: {path} (only pat files are indexed here)")
    if os.path.exists(path + ".csi") and not force:
        print(f"[wt index] {path}.csi already exists. Use -f to overwrite", file=sys.stderr)
        return path + ".csi"
    return index_pat(path)


def main(argv=None):
    p = argparse.ArgumentParser(description="bgzip and index pat files")
    p.add_argument("input_files", nargs="+"); p.add_argument("-f", "--force", action="store_true")
    p.add_argument("-@", "--threads", type=int, default=8)
    a = p.parse_args(argv)
    for f in a.input_files:
        if not os.path.isfile(f):
            raise IllegalArgumentError(f"Invalid file: {f}")
        index_file(f, a.force, a.threads)


if __name__ == "__main__":
    main()
