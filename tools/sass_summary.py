#!/usr/bin/env python
"""SASS evidence for profiles/: per kernel of libwgbs_b200.so (sm_100a), the instruction count and a histogram of the mnemonics that
say how it touches memory and synchronises (LDG / STG widths, LDS / STS, LDGSTS = cp.async, UBLKCP = TMA bulk copy, ATOM / RED, SHFL, VOTE,
REDUX, BAR, WARPSYNC, MUFU, DFMA ...).   python tools/sass_summary.py > profiles/r2_sass_mnemonics.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "wgbs_tools_b200", "libwgbs_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True, check=True).stdout
arch = set(re.findall(r"arch = (sm_\w+)", txt))
print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)}   architectures: {', '.join(sorted(arch))}")
KEEP = ("LDG", "STG", "LDS", "STS", "LDGSTS", "UBLKCP", "UTMA", "ATOM", "RED", "SHFL", "VOTE", "REDUX", "MATCH", "BAR", "WARPSYNC", "MUFU", "DFMA", "DADD", "DMUL", "FFMA", "POPC", "FLO", "BREV",
        "LDL", "STL", "LDC", "SYNCS", "DEPBAR", "LDGDEPBAR", "CCTL", "MEMBAR", "FENCE", "PRMT", "SHF", "LOP3", "IMAD", "BRA", "BSSY")
cur = None; counts = {}
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1); counts[cur] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
    if m and cur:
        op, mods = m.group(1), m.group(2)
        counts[cur]["_total"] += 1
        if op in ("LDG", "STG", "LDS", "STS", "LDGSTS"):
            w = re.search(r"\.(U8|S8|U16|S16|64|128)", mods)
            counts[cur][op + ("." + w.group(1) if w else ".32")] += 1
        elif op.startswith(KEEP):
            counts[cur][op] += 1
def demangle(n):
    r = subprocess.run(["c++filt", n], stdout=subprocess.PIPE, text=True).stdout.strip()
    r = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", r)
    return r.split("(")[0]
for k in sorted(counts, key=lambda k: demangle(k)):
    c = counts[k]
    tot = c.pop("_total", 0)
    if not tot:
        continue
    print(f"{demangle(k)}  [{tot} instructions]")
    print("    " + "  ".join(f"{op}:{n}" for op, n in sorted(c.items(), key=lambda kv: (-kv[1], kv[0]))[:22]))
