"""How quickly does a DEFLATE decoder started at a wrong bit offset fall into step?  Pure-Python walker over the BGZF blocks of a synthetic BAM
(zlib level 6): 20 random start offsets per block, bits / symbols until the trail meets the true symbol boundaries.  The numbers behind
SYNC_BITS in wgbs_tools_b200/csrc/inflate3_core.cuh.  python tools/deflate_sync_stats.py"""
import sys, zlib, struct, random
sys.path.insert(0, '/root/repo')
import numpy as np
from wgbs_tools_b200 import synth, bamio

g = synth.make_genome(7, "chrT", 2_000_000)
sam = synth.make_sam(g, 20000, 3, paired=True)
bam = bamio.sam_to_bam(sam, [("chrT", 2_000_000)])
print(len(sam), len(bam))
# BGZF blocks
blocks = []
p = 0
while p < len(bam):
    xlen = struct.unpack_from('<H', bam, p + 10)[0]
    bsize = struct.unpack_from('<H', bam, p + 16)[0] + 1
    blocks.append(bam[p + 12 + xlen: p + bsize - 8])
    p += bsize
print(len(blocks), 'blocks')

LBASE = [3,4,5,6,7,8,9,10,11,13,15,17,19,23,27,31,35,43,51,59,67,83,99,115,131,163,195,227,258]
LEXT = [0,0,0,0,0,0,0,0,1,1,1,1,2,2,2,2,3,3,3,3,4,4,4,4,5,5,5,5,0]
DEXT = [0,0,0,0,1,1,2,2,3,3,4,4,5,5,6,6,7,7,8,8,9,9,10,10,11,11,12,12,13,13]

def build(lens):
    # canonical: returns dict (len, code) -> sym ; decode via bit-by-bit
    maxl = max(lens)
    cnt = [0] * (maxl + 1)
    for l in lens: cnt[l] += 1
    cnt[0] = 0
    code = 0; nxt = [0] * (maxl + 2)
    for l in range(1, maxl + 1):
        code = (code + cnt[l - 1]) << 1
        nxt[l] = code
    d = {}
    for s, l in enumerate(lens):
        if l:
            d[(l, nxt[l])] = s; nxt[l] += 1
    return d, maxl

class BR:
    def __init__(self, data, pos=0):
        self.v = int.from_bytes(data, 'little'); self.n = len(data) * 8; self.pos = pos
    def bits(self, n):
        r = (self.v >> self.pos) & ((1 << n) - 1); self.pos += n; return r
    def sym(self, tab):
        d, maxl = tab
        code = 0
        for l in range(1, maxl + 1):
            code = (code << 1) | ((self.v >> self.pos) & 1); self.pos += 1
            if (l, code) in d: return d[(l, code)]
        return -1

def walk(br, lt, dt, limit_syms=None, stop=None):
    """yield (pos) of litlen symbol starts"""
    trail = []
    while br.pos < br.n:
        trail.append(br.pos)
        if stop is not None and br.pos in stop: return trail, True
        s = br.sym(lt)
        if s < 0: return trail, False
        if s == 256: return trail, 'eob'
        if s > 256:
            if s > 285: return trail, False
            br.bits(LEXT[s - 257])
            ds = br.sym(dt)
            if ds < 0 or ds > 29: return trail, False
            br.bits(DEXT[ds])
        if limit_syms and len(trail) >= limit_syms: return trail, None
    return trail, False

random.seed(1)
stats = []
nblk_deflate = []
for bi, data in enumerate(blocks[1:40]):
    br = BR(data)
    nb = 0
    while True:
        last = br.bits(1); typ = br.bits(2); nb += 1
        if typ == 0:
            br.pos = (br.pos + 7) & ~7; ln = br.bits(16); br.bits(16); br.pos += 8 * ln
        else:
            if typ == 1:
                ll = [8]*144 + [9]*112 + [7]*24 + [8]*8; dl = [5]*30
            else:
                hlit = br.bits(5) + 257; hdist = br.bits(5) + 1; hclen = br.bits(4) + 4
                order = [16,17,18,0,8,7,9,6,10,5,11,4,12,3,13,2,14,1,15]
                cl = [0]*19
                for i in range(hclen): cl[order[i]] = br.bits(3)
                ct = build(cl)
                lens = []
                while len(lens) < hlit + hdist:
                    s = br.sym(ct)
                    if s < 16: lens.append(s)
                    elif s == 16: lens += [lens[-1]] * (3 + br.bits(2))
                    elif s == 17: lens += [0] * (3 + br.bits(3))
                    else: lens += [0] * (11 + br.bits(7))
                ll = lens[:hlit]; dl = lens[hlit:]
            lt = build(ll); dt = build(dl)
            start = br.pos
            trail, res = walk(br, lt, dt)
            assert res == 'eob', res
            tset = set(trail)
            end = br.pos
            # speculative starts
            for k in range(20):
                p0 = random.randrange(start + 64, max(start + 65, end - 2000))
                b2 = BR(data, p0)
                t2, r2 = walk(b2, lt, dt, limit_syms=3000, stop=tset)
                stats.append((len(t2) - 1, (t2[-1] - p0) if r2 is True else -1, r2))
            if bi < 5: print('block', bi, 'deflate block', nb, 'bits', end - start, 'syms', len(trail), 'maxlen', lt[1], dt[1])
        if last: break
    nblk_deflate.append(nb)
print('deflate blocks per bgzf block', np.bincount(nblk_deflate))
ok = [s for s in stats if s[2] is True]
print('synced', len(ok), 'of', len(stats))
syms = np.array([s[0] for s in ok]); bits = np.array([s[1] for s in ok])
for q in (50, 90, 99, 100): print(q, np.percentile(syms, q), np.percentile(bits, q))
print([s for s in stats if s[2] is not True][:10])
