// pat_core.cuh -- the per-line logic of the two-pass tile parser for pat text (pat_tiles_k in pat.cu): field search on the
// newline / tab masks of a tile, integer fields, the record of a line.  Host + device: tests/pat_core_check.cpp builds it with
// g++ and runs the kernel's two passes sequentially over real text, so the code the device runs is pinned without a GPU.
#pragma once
#include <stdint.h>

#include "pats.cuh"

#if defined(__CUDACC__)
#define WGBS_HD __host__ __device__ __forceinline__
#else
#define WGBS_HD inline
#endif

WGBS_HD int pat_ctz64(unsigned long long v) {      // index of the lowest set bit (v != 0)
#if defined(__CUDA_ARCH__)
    return __ffsll((long long)v) - 1;
#else
    return __builtin_ctzll(v);
#endif
}

constexpr int PS_T = 256, PS_SPAN = 64, PS_TILE = PS_T * PS_SPAN;

struct PatTile {
    const char *g; uint32_t n;                   // the whole text
    const unsigned char *sm; uint32_t t0, t1;    // bytes [t0, t1) of it are resident in sm
    const unsigned long long *nl, *tab;          // per 64-byte span of the tile: bit b set iff byte b of the span is '\n' / '\t'
    WGBS_HD uint32_t byte(uint32_t p) const { return (p >= t0 && p < t1) ? sm[p - t0] : (uint32_t)(uint8_t)g[p]; }
    // first position >= p holding the character the masks m[] stand for (c); n when there is none
    WGBS_HD uint32_t next(const unsigned long long *m, uint32_t c, uint32_t p) const {
        if (p >= t0 && p < t1) {
            uint32_t k = (p - t0) >> 6;
            unsigned long long wd = m[k] & (~0ull << ((p - t0) & 63));
            const uint32_t ks = (t1 - t0 + 63) >> 6;
            while (!wd && ++k < ks) wd = m[k];
            if (wd) return t0 + (k << 6) + (uint32_t)pat_ctz64(wd);
            p = t1;
        }
        while (p < n && (uint32_t)(uint8_t)g[p] != c) p++;
        return p;
    }
};
// parse_int_field on a tile accessor (std::stoi semantics, see above)
WGBS_HD bool parse_int_tile(const PatTile &t, uint32_t s, uint32_t e, int32_t *out) {
    while (s < e) { const uint32_t c = t.byte(s); if (c == ' ' || (c >= 9 && c <= 13)) s++; else break; }
    bool neg = false;
    if (s < e) { const uint32_t c = t.byte(s); if (c == '+' || c == '-') { neg = c == '-'; s++; } }
    if (s >= e) return false;
    uint32_t c = t.byte(s);
    if (c < '0' || c > '9') return false;
    int64_t v = 0;
    while (true) {
        v = v * 10 + (int64_t)(c - '0'); if (v > 0x80000000LL) return false;
        if (++s >= e) break;
        c = t.byte(s); if (c < '0' || c > '9') break;
    }
    if (neg) v = -v;
    if (v > 0x7fffffffLL || v < -0x80000000LL) return false;
    *out = (int32_t)v;
    return true;
}
struct PatRec { uint32_t idx, len, cnt, ps, err; };   // len: length of the pattern field (also when err == 2); err as pat_lines_k
// the line starting at s (s < n).  full = false: only len / ps / err 1 (what the pool offsets need)
WGBS_HD PatRec pat_line(const PatTile &T, uint32_t s, bool full) {
    PatRec r; r.idx = 0; r.len = 0; r.cnt = 0; r.ps = s; r.err = 0;
    const uint32_t e = T.next(T.nl, '\n', s);
    if (e == s) return r;                                            // empty line: a record of zeros
    uint32_t tab[4]; int nt = 0; uint32_t p = s;
    while (nt < 4) { const uint32_t t = T.next(T.tab, '\t', p); if (t >= e) break; tab[nt++] = t; p = t + 1; }
    if (nt < 3) { r.err = 1; return r; }
    r.len = tab[2] - tab[1] - 1; r.ps = tab[1] + 1;
    if (!full) return r;
    const uint32_t cend = nt >= 4 ? tab[3] : e;
    int32_t vi = 0, vc = 0;
    if (!parse_int_tile(T, tab[0] + 1, tab[1], &vi) || !parse_int_tile(T, tab[2] + 1, cend, &vc)) { r.err = 2; return r; }
    r.idx = (uint32_t)vi; r.cnt = (uint32_t)vc;
    return r;
}

