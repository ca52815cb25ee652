"""TEST INFRASTRUCTURE ONLY: run the compiled reference executables (oracle/_ref/*) and the C restatement
(oracle/_ref/liboracle_port.so) on in-memory bytes.  Never imported by wgbs_tools_b200/."""
from __future__ import annotations

import ctypes
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
SHIM = os.path.join(HERE, "shim")


def have_ref() -> bool:
    return all(os.access(os.path.join(REF, t), os.X_OK) for t in ("patter", "match_maker", "stdin2beta", "homog", "segmentor"))


def have_port() -> bool:
    return os.path.isfile(os.path.join(REF, "liboracle_port.so"))


def build() -> None:
    subprocess.check_call(["make", "-s", "-C", HERE, "-j4"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def _env():
    e = dict(os.environ)
    e["PATH"] = SHIM + os.pathsep + e.get("PATH", "")
    e["LC_ALL"] = "C"
    return e


def tool(name: str, opt: bool = False) -> str:
    return os.path.join(REF, name + (".O2" if opt else ""))


# ------------------------------------------------------------------------------------------------------------
# reference executables
# ------------------------------------------------------------------------------------------------------------

def ref_patter(sam: bytes, dict_path: str, region: str, paired: bool, min_cpg: int = 1, clip: int = 0,
               nanopore: bool = False, np_thresh: float | None = None, cpc_call: str | None = None,
               combine_mods: bool = False, opt: bool = False, mbias: str | None = None, long: bool = False):
    """`[match_maker |] patter DICT REGION ...`  (reference bam2pat.py:186-204). returns (stdout, stderr)."""
    cmd = ""
    if paired:
        cmd += f"{tool('match_maker', opt)} | "
    cmd += f"{tool('patter', opt)} {dict_path} {region} --min_cpg {min_cpg} --clip {clip}"
    if nanopore:
        cmd += " --nanopore"
        if np_thresh is not None:
            cmd += f" --np_thresh {np_thresh}"
    if cpc_call is not None:
        cmd += f" --cpc_call {cpc_call}"
    if combine_mods:
        cmd += " --combine_mods"
    if mbias:
        cmd += f" --mbias {mbias}"
    if long:
        cmd += " --long"
    p = subprocess.run(cmd, shell=True, input=sam, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=_env())
    return p.stdout, p.stderr


def read_mbias(prefix: str) -> np.ndarray:
    """<prefix>.OT.txt / .OB.txt written by `patter --mbias` (patter.cpp:50-72) -> int32[2][2][1000][2] = [OT|OB][mate][pos][meth|unmeth]"""
    out = np.zeros((2, 2, 1000, 2), np.int32)
    for s, x in enumerate(("OT", "OB")):
        rows = np.loadtxt(f"{prefix}.{x}.txt", skiprows=1, dtype=np.int64).reshape(1000, 4)
        out[s, 0, :, 0] = rows[:, 0]; out[s, 0, :, 1] = rows[:, 1]; out[s, 1, :, 0] = rows[:, 2]; out[s, 1, :, 1] = rows[:, 3]
    return out


def ref_collapse(txt: bytes) -> bytes:
    """reference bam2pat.py:99-106 (without bgzip)."""
    cmd = "sort -k2,2n -k3,3 | uniq -c | awk -v OFS='\\t' '{print $2,$3,$4,$1}'"
    return subprocess.run(cmd, shell=True, input=txt, stdout=subprocess.PIPE, env=_env(), check=True).stdout


def ref_collapse_long(txt: bytes) -> bytes:
    """reference bam2pat.py:102-103 (--long), without bgzip"""
    cmd = "sort -k2,2n -k3,3 | awk -v OFS='\\t' '{print $1,$2,$3,1,$4}'"
    return subprocess.run(cmd, shell=True, input=txt, stdout=subprocess.PIPE, env=_env(), check=True).stdout


def ref_stdin2beta(pat: bytes, start: int, end: int, opt: bool = False) -> np.ndarray:
    """reference pat2beta.py:32-35. returns int64[n,2] (empty array when the tool printed nothing)."""
    p = subprocess.run([tool("stdin2beta", opt), str(start), str(end)], input=pat, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE)
    return np.array(p.stdout.split(), dtype=np.int64).reshape(-1, 2)


def ref_trim(arr: np.ndarray, lbeta: bool = False) -> np.ndarray:
    """verbatim arithmetic of reference utils_wgbs.py:277-290 (numpy float64)."""
    data = arr.astype(np.int64).copy()
    nr_bits = 16 if lbeta else 8
    dtype = np.uint16 if lbeta else np.uint8
    max_val = 2 ** nr_bits - 1
    big = np.argwhere(data[:, 1] > max_val).flatten()
    data[:, 0][big] = data[big][:, 0] / data[big][:, 1] * max_val
    data[:, 1][big] = max_val
    return data.astype(dtype)


def ref_homog(pat: bytes, blocks_path: str, range_str: str, min_len: int, inclusive: bool = False,
              chrom: str | None = None, sort_blocks: bool = False, opt: bool = False) -> np.ndarray:
    cmd = [tool("homog", opt), "-b", blocks_path, "-r", range_str, "-l", str(min_len)]
    if inclusive:
        cmd.append("--inclusive")
    if chrom:
        cmd += ["--chrom", chrom]
    if sort_blocks:
        cmd.append("--sort_blocks")
    p = subprocess.run(cmd, input=pat, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=_env())
    rows = [l.split(b"\t") for l in p.stdout.splitlines()]
    return np.array(rows, dtype=np.int64).reshape(len(rows), -1)


def ref_segmentor(beta_paths, start0: int, n: int, max_cpg: int, max_bp: int, pseudo: float, dists: np.ndarray,
                  opt: bool = False) -> np.ndarray:
    cmd = [tool("segmentor", opt), *beta_paths, "-s", str(start0), "-n", str(n), "-max_cpg", str(max_cpg),
           "-ps", repr(float(pseudo)) if pseudo != int(pseudo) else str(int(pseudo)), "-max_bp", str(max_bp)]
    inp = b"".join(b"%d\n" % d for d in dists.tolist())
    p = subprocess.run(cmd, input=inp, stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True)
    return np.array(p.stdout.split(), dtype=np.int64)


def write_tmp(data: bytes, suffix: str = "") -> str:
    fd, path = tempfile.mkstemp(suffix=suffix, dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    with os.fdopen(fd, "wb") as f:
        f.write(data)
    return path


# ------------------------------------------------------------------------------------------------------------
# the C restatement
# ------------------------------------------------------------------------------------------------------------
_port = None


def port():
    global _port
    if _port is None:
        L = ctypes.CDLL(os.path.join(REF, "liboracle_port.so"))
        L.port_patter.restype = ctypes.c_long
        L.port_match_maker.restype = ctypes.c_long
        L.port_collapse.restype = ctypes.c_long
        _port = L
    return _port


def _take(L, ptr, n) -> bytes:
    b = ctypes.string_at(ptr, n)
    L.port_free(ptr)
    return b


def port_match_maker(sam: bytes) -> bytes:
    L = port(); out = ctypes.c_void_p()
    n = L.port_match_maker(sam, ctypes.c_long(len(sam)), ctypes.byref(out))
    return _take(L, out, n)


def port_patter(sam: bytes, loci: np.ndarray, idx: np.ndarray, min_cpg=1, clip=0, nanopore=False, np_thresh=0.67,
                cpc_call="C", combine_mods=False):
    """returns (pat lines as patter prints them, stats[lines,pairs,empty,short,invalid,is_pe])."""
    L = port(); out = ctypes.c_void_p(); stats = (ctypes.c_long * 6)()
    loci = np.ascontiguousarray(loci, np.int32); idx = np.ascontiguousarray(idx, np.int32)
    n = L.port_patter(sam, ctypes.c_long(len(sam)), loci.ctypes.data_as(ctypes.c_void_p), idx.ctypes.data_as(ctypes.c_void_p),
                      ctypes.c_int(loci.size), ctypes.c_int(min_cpg), ctypes.c_int(clip), ctypes.c_int(int(nanopore)),
                      ctypes.c_float(np_thresh), ctypes.c_char(cpc_call.encode()), ctypes.c_int(int(combine_mods)),
                      ctypes.byref(out), stats)
    if n < 0:
        raise RuntimeError("port_patter: fatal (first line)")
    return _take(L, out, n), list(stats)


def port_collapse(txt: bytes) -> bytes:
    L = port(); out = ctypes.c_void_p()
    n = L.port_collapse(txt, ctypes.c_long(len(txt)), ctypes.byref(out))
    return _take(L, out, n)


def port_pat2beta(pat: bytes, start: int, end: int) -> np.ndarray:
    L = port(); mc = np.zeros((end - start, 2), np.int32)
    rc = L.port_pat2beta(pat, ctypes.c_long(len(pat)), ctypes.c_int(start), ctypes.c_int(end), mc.ctypes.data_as(ctypes.c_void_p))
    if rc:
        raise ValueError("failed calculating beta")
    return mc


def port_trim(mc: np.ndarray, nbits: int = 8) -> np.ndarray:
    L = port(); a = np.ascontiguousarray(mc, np.int64); out = np.zeros(a.shape, np.uint8 if nbits == 8 else np.uint16)
    L.port_trim(a.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(a.shape[0]), ctypes.c_int(nbits), out.ctypes.data_as(ctypes.c_void_p))
    return out


def port_homog(pat: bytes, blocks: np.ndarray, rng: np.ndarray, min_cpgs: int, inclusive: bool = False) -> np.ndarray:
    L = port(); bs = np.ascontiguousarray(blocks[:, 0], np.int32); be = np.ascontiguousarray(blocks[:, 1], np.int32)
    r = np.ascontiguousarray(rng, np.float32); nb = r.size - 1
    out = np.zeros((bs.size, nb), np.int32)
    rc = L.port_homog(pat, ctypes.c_long(len(pat)), bs.ctypes.data_as(ctypes.c_void_p), be.ctypes.data_as(ctypes.c_void_p),
                      ctypes.c_long(bs.size), r.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(nb), ctypes.c_int(min_cpgs),
                      ctypes.c_int(int(inclusive)), out.ctypes.data_as(ctypes.c_void_p))
    if rc:
        raise ValueError("failed calculating homog")
    return out


def port_segment(betas, dists: np.ndarray, max_cpg: int, max_bp: int, pseudo: float) -> np.ndarray:
    L = port(); K = len(betas); n = betas[0].shape[0]
    arrs = [np.ascontiguousarray(b, np.uint8) for b in betas]
    ptrs = (ctypes.c_void_p * K)(*[a.ctypes.data for a in arrs])
    d = np.ascontiguousarray(dists, np.uint32); out = np.zeros(n + 2, np.int32)
    L.port_segment.restype = ctypes.c_int
    nb = L.port_segment(ptrs, ctypes.c_int(K), d.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(n), ctypes.c_int(max_cpg),
                        ctypes.c_uint32(max_bp), ctypes.c_float(pseudo), out.ctypes.data_as(ctypes.c_void_p))
    return out[:nb].astype(np.int64)


# ------------------------------------------------------------------------------------------------------------
# cview (reference src/cview/cview.cpp) + the `| sort -k2,2n -k3,3 | collapse_pat.pl -` tail of cview.py
# ------------------------------------------------------------------------------------------------------------
def have_cview() -> bool:
    return os.access(os.path.join(REF, "cview"), os.X_OK)


def ref_cview(pat: bytes, *, sites: tuple[int, int] | None = None, blocks_path: str | None = None, strict=False, strip=False,
              no_gaps=False, min_cpgs: int = 1) -> bytes:
    """`cview --sites "s\\te" | --blocks_path F [--strict] [--strip] [--no_gaps] [--min_cpgs N]` on pat text"""
    cmd = [tool("cview")]
    cmd += ["--sites", f"{sites[0]}\t{sites[1]}"] if sites is not None else ["--blocks_path", blocks_path]
    cmd += ["--strict"] * bool(strict) + ["--strip"] * bool(strip) + ["--no_gaps"] * bool(no_gaps)
    if min_cpgs != 1:
        cmd += ["--min_cpgs", str(min_cpgs)]
    return subprocess.run(cmd, input=pat, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, env=_env()).stdout


def port_collapse_pat(text: bytes) -> bytes:
    """restatement of reference src/collapse_pat.pl: adjacent lines equal in every column but the 4th are merged,
    their counts summed (lines whose summed count is 0 vanish)"""
    out = []; prev = None; count = 0
    for line in text.splitlines():
        w = line.split(b"\t")
        if prev is not None and len(w) <= len(prev) and all(w[i] == prev[i] for i in range(len(w)) if i != 3):
            count += int(w[3])
        else:
            if prev is not None and count > 0:
                prev[3] = b"%d" % count; out.append(b"\t".join(prev))
            count = int(w[3])
        prev = w
    if prev is not None and count > 0:
        prev[3] = b"%d" % count; out.append(b"\t".join(prev))
    return b"".join(l + b"\n" for l in out)


def ref_collapse_pat(text: bytes) -> bytes | None:
    """the reference's own perl script, where /root/reference is mounted (this container only)"""
    script = "/root/reference/src/collapse_pat.pl"
    if not os.path.isfile(script):
        return None
    f = write_tmp(text, ".pat")
    return subprocess.run(["perl", script, f], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, env=_env()).stdout


def sort_pat(text: bytes) -> bytes:
    return subprocess.run(["sort", "-k2,2n", "-k3,3"], input=text, stdout=subprocess.PIPE, env=_env()).stdout
