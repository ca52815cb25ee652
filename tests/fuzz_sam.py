"""Seeded generator of hostile SAM text for the fuzz tests (CPU: port vs reference; GPU: kernels vs port)."""
import numpy as np

from wgbs_tools_b200 import synth


def fuzz_sam(g, seed: int, n: int = 600, np_mode: bool = False) -> bytes:
    rng = np.random.default_rng(seed)
    base = (synth.make_np_sam(g, n, seed) if np_mode else synth.make_sam(g, n, seed, paired=False)).splitlines()
    out = []
    for l in base:
        t = l.split(b"\t")
        r = rng.random()
        k = int(rng.integers(0, 24))
        if r < 0.45:
            pass                                           # leave valid
        elif k == 0:
            t[5] = bytes(rng.choice(list(b"0123456789MIDNSHP=X*B"), size=int(rng.integers(0, 12))).tolist())   # random CIGAR
        elif k == 1:
            t = t[: int(rng.integers(0, 11))]              # too few fields
        elif k == 2:
            t[9] = t[9][: int(rng.integers(0, len(t[9]) + 1))]   # SEQ shorter than CIGAR
        elif k == 3:
            t[9] = t[9] + b"ACGT" * int(rng.integers(1, 5))      # SEQ longer than CIGAR
        elif k == 4:
            t[3] = [b"0", b"-5", b"abc", b"", b"99999999999", b"1", b" 12", b"+7", b"12x"][int(rng.integers(0, 9))]   # POS oddities
        elif k == 5:
            t[1] = [b"x", b"", b"-1", b"65535", b"4294967296", b"16 "][int(rng.integers(0, 6))]                      # FLAG oddities
        elif k == 6:
            t[9] = b"*"
        elif k == 7:
            t[9] = t[9].lower()
        elif k == 8:
            t[5] = b"%dM" % int(rng.integers(0, 400))
        elif k == 9:
            t[5] = b"5H" + t[5] + b"3H"
        elif k == 10:
            t[5] = b"%dS%dM%dN%dM" % tuple(int(x) for x in rng.integers(1, 60, size=4))
        elif k == 11:
            t = t + [b""] if rng.random() < 0.5 else t[:11] + [b""]          # trailing tab
        elif k == 12:
            t[0] = b""                                     # empty QNAME
        elif k == 13:
            t[9] = bytes(rng.choice(list(b"ACGTNRYacgt=."), size=len(t[9])).tolist())
        elif k == 14:
            t[5] = t[5].replace(b"M", b"=", 1) if rng.random() < 0.5 else t[5].replace(b"M", b"X", 1)
        elif k == 15:
            t[5] = b"2147483648M"
        elif k == 16:
            t[5] = b"10M 5M"
        elif k == 17:
            t[3] = b"%d" % int(g.loci[-1] - rng.integers(0, 100))            # runs past the last CpG
        elif k == 18:
            t[3] = b"1"
        elif k == 19 and np_mode and len(t) > 11:
            tags = [b"MM:Z:C+m?,1,,2;", b"MM:Z:C+m,;", b"MM:Z:C+m?;C+h?;", b"MM:Z:C+m?,0,0", b"MM:Z:;;C+m.,0;", b"MM:Z:C+m?,99999;",
                    b"MM:Z:C+m?,0,1;C+h?,0,1;\tML:B:C,1,2,3", b"MM:Z:C+h.,0;\tML:B:C", b"MM:Z:C+m?,0,1;\tML:B:C,255", b"MM:Z:C+C?,0;C+m?,0;\tML:B:C,5,250",
                    b"Mm:Z:C+m?,0,1;\tMl:B:C,200,10", b"MM:Z:C+m?,0\tMM:Z:C+m?,1;\tML:B:C,250"]
            t = t[:12] + [tags[int(rng.integers(0, len(tags)))]]
        elif k == 20:
            out.append(b"")                                # blank line
        elif k == 21:
            t[5] = b"0M" + t[5]
        elif k == 22:
            t[5] = t[5] + b"7"                             # trailing digits without an op
        elif k == 23:
            t[5] = b"3I" + t[5][:-1] + b"M2D"
        out.append(b"\t".join(t))
    return b"\n".join(out) + b"\n"
