"""TEST INFRASTRUCTURE: a Python model of pat_tiles_k (wgbs_tools_b200/csrc/pat.cu, the staged two-pass tile parser for pat text):
tile / span ownership of lines, line ordering through the per-tile prefixes, the mask-based field search across spans and
tiles -- against a straightforward parser, with tiny tiles so that every boundary case occurs.  The kernel was written
without GPU access; this pins its algorithm (tests/test_host_logic.py runs it)."""
import random
def model(text: bytes, PS_T=4, SPAN=8):
    n = len(text); TILE = PS_T*SPAN
    ntiles = (n + TILE - 1)//TILE
    def run(PASS, tlo=None, two=None, out=None):
        tiles_l, tiles_w = [], []
        for tile in range(ntiles):
            t0 = tile*TILE; t1 = min(t0+TILE, n)
            sm = bytearray(TILE); sm[:t1-t0] = text[t0:t1]
            nlm = []; tbm = []
            for tid in range(PS_T):
                nl = tb = 0
                for b in range(SPAN):
                    c = sm[tid*SPAN+b]
                    if c == 10: nl |= 1 << b
                    if c == 9: tb |= 1 << b
                nlm.append(nl); tbm.append(tb)
            def byte(p): return sm[p-t0] if t0 <= p < t1 else text[p]
            def nxt(m, c, p):
                if t0 <= p < t1:
                    k = (p-t0)//SPAN
                    wd = m[k] & ~((1 << ((p-t0) % SPAN)) - 1)
                    ks = (t1-t0+SPAN-1)//SPAN
                    while not wd and k+1 < ks:
                        k += 1; wd = m[k]
                    if wd: return t0 + k*SPAN + ((wd & -wd).bit_length()-1)
                    p = t1
                while p < n and text[p] != c: p += 1
                return p
            def parse_int(s, e):
                while s < e and (byte(s) == 32 or 9 <= byte(s) <= 13): s += 1
                neg = False
                if s < e and byte(s) in (43, 45): neg = byte(s) == 45; s += 1
                if s >= e or not (48 <= byte(s) <= 57): return None
                v = 0
                while s < e and 48 <= byte(s) <= 57:
                    v = v*10 + byte(s)-48
                    if v > 0x80000000: return None
                    s += 1
                if neg: v = -v
                if v > 0x7fffffff or v < -0x80000000: return None
                return v
            def line(s, full):
                r = dict(idx=0, len=0, cnt=0, ps=s, err=0)
                e = nxt(nlm, 10, s)
                if e == s: return r
                tab = []; p = s
                while len(tab) < 4:
                    t = nxt(tbm, 9, p)
                    if t >= e: break
                    tab.append(t); p = t+1
                if len(tab) < 3: r['err'] = 1; return r
                r['len'] = tab[2]-tab[1]-1; r['ps'] = tab[1]+1
                if not full: return r
                cend = tab[3] if len(tab) >= 4 else e
                vi = parse_int(tab[0]+1, tab[1]); vc = parse_int(tab[2]+1, cend)
                if vi is None or vc is None: r['err'] = 2; return r
                r['idx'] = vi & 0xffffffff; r['cnt'] = vc & 0xffffffff
                return r
            my = []
            for tid in range(PS_T):
                span0 = t0 + tid*SPAN
                starts = [0] if (tile == 0 and tid == 0 and n > 0) else []
                m = nlm[tid]
                while m:
                    b = (m & -m).bit_length()-1; m &= m-1
                    s = span0+b+1
                    if s < n: starts.append(s)
                my.append(starts)
            cl = [len(x) for x in my]; cw = [sum((line(s, False)['len']+15) >> 4 for s in x) for x in my]
            if PASS == 0:
                tiles_l.append(sum(cl)); tiles_w.append(sum(cw)); continue
            for tid in range(PS_T):
                ln = tlo[tile] + sum(cl[:tid]); o = two[tile] + sum(cw[:tid])
                for s in my[tid]:
                    r = line(s, True)
                    if r['err']: out['err'] |= r['err']
                    L = 0 if r['err'] else r['len']
                    out['rec'][ln] = (r['idx'], L, r['cnt'], o)
                    for b in range(0, L, 16):
                        out['pool'][o + b//16] = bytes(byte(r['ps']+b+k) for k in range(min(16, L-b)))
                    o += (r['len']+15) >> 4; ln += 1
        return tiles_l, tiles_w
    tl, tw = run(0)
    tlo = [0]; two = [0]
    for a in tl: tlo.append(tlo[-1]+a)
    for a in tw: two.append(two[-1]+a)
    out = dict(err=0, rec=[None]*tlo[-1], pool=[None]*two[-1])
    run(1, tlo, two, out)
    return out, tlo[-1], two[-1]

def straight(text: bytes):
    n = len(text)
    lines = text.split(b"\n")
    if text.endswith(b"\n") or n == 0: lines = lines[:-1]
    rec = []; pool = []; err = 0
    for l in lines:
        if not l: rec.append((0, 0, 0, len(pool))); continue
        f = l.split(b"\t")
        if len(f) < 4: err |= 1; rec.append((0, 0, 0, len(pool))); continue
        def pi(x):
            x = x.lstrip(b" \t\n\v\f\r")
            import re
            mm = re.match(rb"[+-]?\d+", x)
            if not mm: return None
            v = int(mm.group())
            return v if -0x80000000 <= v <= 0x7fffffff else None
        # note: tabs beyond the 4th belong to extra columns: count field is f[3]
        vi, vc = pi(f[1]), pi(f[3])
        if vi is None or vc is None:
            err |= 2; o = len(pool); pool += [None]*((len(f[2])+15)//16); rec.append((0, 0, 0, o)); continue
        o = len(pool)
        for b in range(0, len(f[2]), 16): pool.append(f[2][b:b+16])
        rec.append((vi & 0xffffffff, len(f[2]), vc & 0xffffffff, o))
    return rec, pool, err

def rnd_text():
    parts = []
    for _ in range(random.randint(0, 30)):
        k = random.random()
        if k < 0.1: parts.append(b"")
        elif k < 0.15: parts.append(b"chr1\t5")
        elif k < 0.2: parts.append(b"chr1\tx\tCT\t1")
        else:
            pat = bytes(random.choice(b"CT.H") for _ in range(random.randint(1, 40)))
            l = b"chr%d\t%d\t%s\t%d" % (random.randint(1, 22), random.randint(1, 10**random.randint(1, 8)), pat, random.randint(1, 300))
            if random.random() < 0.2: l += b"\textra\tcols"
            parts.append(l)
    t = b"\n".join(parts)
    if random.random() < 0.7 and parts: t += b"\n"
    return t

def run(iters=3000, seed=1):
    random.seed(seed)
    bad = 0
    for it in range(iters):
        t = rnd_text()
        T, S = random.choice([(4, 8), (2, 16), (8, 8), (3, 8)])
        out, nl, nw = model(t, T, S)
        rec, pool, err = straight(t)
        ok = (nl == len(rec)) and out['err'] == err
        if err == 0:
            ok = ok and out['rec'] == rec and out['pool'] == pool and nw == len(pool)
        bad += not ok
    return bad


if __name__ == "__main__":
    print("bad", run())
