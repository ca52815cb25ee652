"""Build libwgbs_b200.so (sm_100a only) in-tree with nvcc.  `python -m wgbs_tools_b200.build [-v]`."""
from __future__ import annotations

import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "libwgbs_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC,-O2",
         "--fmad=false",   # bit-exactness: no silent FMA contraction anywhere; kernels that want FMA say __fma_rn
         "-Xcudafe", "--diag_suppress=177"]


def _newer(src: str, dst: str, deps) -> bool:
    if not os.path.exists(dst):
        return True
    t = os.path.getmtime(dst)
    return any(os.path.getmtime(d) > t for d in [src, *deps])


def build(verbose: bool = False, force: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdrs = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    objs = [os.path.join(OBJ, os.path.basename(s)[:-3] + ".o") for s in srcs]

    def cc(pair):
        s, o = pair
        if not force and not _newer(s, o, hdrs):
            return
        cmd = [NVCC, *FLAGS, "-c", s, "-o", o] + (["-Xptxas", "-v"] if verbose else [])
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {s}:\n{r.stdout}")
        if verbose:
            sys.stderr.write(r.stdout)

    with ThreadPoolExecutor(max_workers=min(8, len(srcs) or 1)) as ex:
        list(ex.map(cc, zip(srcs, objs)))
    if force or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", *objs, "-o", LIB, "-lcudart_static", "-lpthread", "-ldl", "-lrt", "-lz"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}")
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
