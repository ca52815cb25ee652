// inflate_core.cuh -- raw DEFLATE (RFC 1951) decoder for ONE BGZF block, written for a warp: lane 0 walks the Huffman
// stream and queues up to one symbol per lane; then the whole warp materialises the queue (literals in parallel, LZ77
// matches in stream order, each copied cooperatively).  The output buffer is the LZ77 window: a BGZF block is an
// independent deflate stream of <= 64 KiB (SAM spec 4.1), so distances never leave the block's own output.
//
// The same code compiles as plain host C++ with a one-lane policy (tests/bamdev_core_check.cpp pins it against zlib
// without a GPU); the device policy is the 32 lanes of a warp.  Replaces the zlib inflate samtools/htslib run on the
// host for the reference's `samtools view` stage (reference src/python/bam2pat.py:165).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define WGBS_HD __host__ __device__ __forceinline__
#else
#define WGBS_HD inline
#endif

namespace dflate {

constexpr int LBITS = 10;   // literal/length codes up to this length decode with one table probe
constexpr int DBITS = 8;    // distance codes

enum : int { OK = 0, E_INPUT = -1 /* ran out of input */, E_BTYPE = -2, E_STORED = -3, E_CODES = -4 /* bad code lengths */,
             E_SYMBOL = -5 /* invalid code / symbol */, E_DIST = -6 /* distance before the start of the block */,
             E_OUTPUT = -7 /* more output than ISIZE */, E_SHORT = -8 /* stream ended before ISIZE bytes */,
             E_CRC = -9 /* CRC32 of the inflated bytes differs from the block trailer */ };

// per-warp working set (shared memory on the device): 3.7 KB
struct Scratch {
    uint16_t lt[1 << LBITS];     // fast tables: symbol | code length << 9 (0: longer than the table, or unassigned)
    uint16_t dt[1 << DBITS];
    uint16_t lsym[288], dsym[32];   // symbols in canonical order + per-length counts, for the long codes
    uint16_t lcnt[16], dcnt[16];
    uint16_t lfirst[16], loffs[16], dfirst[16], doffs[16];   // first canonical code / first index in *sym per code length
    uint32_t q[32];              // decoded symbols: literal byte, or 0x80000000 | (dist-1) << 9 | length
    uint8_t lens[320];
};

// ---- lane policies ---------------------------------------------------------------------------------------------------
struct OneLane {
    static constexpr int N = 1;
    WGBS_HD int id() const { return 0; }
    WGBS_HD void sync() const {}
    WGBS_HD uint32_t shfl(uint32_t v, int) const { return v; }
    WGBS_HD uint32_t ballot(bool p) const { return p ? 1u : 0u; }
    WGBS_HD uint32_t exscan(uint32_t, uint32_t *total, uint32_t v_total) const { *total = v_total; return 0; }
};
#if defined(__CUDACC__)
struct WarpLanes {
    static constexpr int N = 32;
    __device__ __forceinline__ int id() const { return (int)(threadIdx.x & 31); }
    __device__ __forceinline__ void sync() const { __syncwarp(); }
    __device__ __forceinline__ uint32_t shfl(uint32_t v, int src) const { return __shfl_sync(0xffffffffu, v, src); }
    __device__ __forceinline__ uint32_t ballot(bool p) const { return __ballot_sync(0xffffffffu, p); }
    // exclusive prefix sum of v over the lanes; *total = sum
    __device__ __forceinline__ uint32_t exscan(uint32_t v, uint32_t *total, uint32_t) const {
        uint32_t x = v;
        const int l = id();
        for (int d = 1; d < 32; d <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, d); if (l >= d) x += y; }
        *total = __shfl_sync(0xffffffffu, x, 31);
        return x - v;
    }
};
// G consecutive lanes of a warp (G = 4, 8, 16) as an independent team WITH collectives: 32 / G teams share one warp's
// instruction stream, each on its own item.  Every collective names only the team's own lanes in its mask, so teams may sit
// at different points of the program (independent thread scheduling); where they happen to be at the same point, one
// issued instruction serves all of them -- that is the point: the decoder's per-symbol work is uniform over the lanes of a
// team, so a whole warp per BGZF block spends 32 lanes on what G do just as well.
template <int G>
struct SubWarp {
    static_assert(G == 4 || G == 8 || G == 16 || G == 32, "team size");
    static constexpr int N = G;
    __device__ __forceinline__ unsigned lane() const { return threadIdx.x & 31u; }
    __device__ __forceinline__ unsigned shift() const { return lane() & ~(unsigned)(G - 1); }
    __device__ __forceinline__ unsigned mask() const { return (G == 32 ? 0xffffffffu : ((1u << (G & 31)) - 1u)) << shift(); }
    __device__ __forceinline__ int id() const { return (int)(threadIdx.x & (G - 1)); }
    __device__ __forceinline__ void sync() const { __syncwarp(mask()); }
    __device__ __forceinline__ uint32_t shfl(uint32_t v, int src) const { return __shfl_sync(mask(), v, src, G); }
    __device__ __forceinline__ uint32_t ballot(bool p) const { return (__ballot_sync(mask(), p) >> shift()) & (G == 32 ? 0xffffffffu : ((1u << (G & 31)) - 1u)); }
    __device__ __forceinline__ uint32_t exscan(uint32_t v, uint32_t *total, uint32_t) const {
        uint32_t x = v;
        const int l = id();
        const unsigned m = mask();
#pragma unroll
        for (int d = 1; d < G; d <<= 1) { uint32_t y = __shfl_up_sync(m, x, d, G); if (l >= d) x += y; }
        *total = __shfl_sync(m, x, G - 1, G);
        return x - v;
    }
};
// G consecutive lanes working on one item (no collectives: only id() / N are meaningful)
template <int G>
struct LaneGroup {
    static constexpr int N = G;
    __device__ __forceinline__ int id() const { return (int)(threadIdx.x & (G - 1)); }
};
#endif

WGBS_HD int lowest_bit(uint32_t m) {             // index of the lowest set bit (m != 0)
#if defined(__CUDA_ARCH__)
    return __ffs((int)m) - 1;
#else
    return __builtin_ctz(m);
#endif
}
WGBS_HD uint32_t brev32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __brev(v);
#else
    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1); v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
    v = ((v >> 4) & 0x0f0f0f0fu) | ((v & 0x0f0f0f0fu) << 4); v = ((v >> 8) & 0x00ff00ffu) | ((v & 0x00ff00ffu) << 8);
    return (v >> 16) | (v << 16);
#endif
}

// ---- canonical Huffman tables ----------------------------------------------------------------------------------------
WGBS_HD uint32_t bitrev(uint32_t c, int n) { uint32_t r = 0; for (int i = 0; i < n; i++) { r = (r << 1) | (c & 1); c >>= 1; } return r; }

// lens[0..n) -> cnt[16], sym[] (canonical order), fast[1<<fbits].  Over-subscribed sets are rejected; incomplete sets are
// rejected too, except (single_ok: literal/length and distance sets, not the code-length code) when no code is longer
// than one bit -- the cases zlib's inftrees.c accepts; an unused slot decodes to "invalid code" if the stream reaches it.
WGBS_HD int build_table(const uint8_t *lens, int n, uint16_t *cnt, uint16_t *sym, uint16_t *fast, int fbits, bool single_ok,
                        uint16_t *first_out = nullptr, uint16_t *offs_out = nullptr) {
    for (int i = 0; i < 16; i++) cnt[i] = 0;
    for (int i = 0; i < n; i++) cnt[lens[i]]++;
    for (int i = 0; i < (1 << fbits); i++) fast[i] = 0;
    int left = 1, maxlen = 0;
    for (int l = 1; l <= 15; l++) { left <<= 1; left -= cnt[l]; if (left < 0) return E_CODES; if (cnt[l]) maxlen = l; }
    if (left > 0 && (maxlen > 1 || !single_ok)) return E_CODES;
    uint16_t offs[16], code[16];
    offs[1] = 0; code[1] = 0;
    for (int l = 1; l < 15; l++) { offs[l + 1] = (uint16_t)(offs[l] + cnt[l]); code[l + 1] = (uint16_t)((code[l] + cnt[l]) << 1); }
    if (first_out) for (int l = 1; l <= 15; l++) { first_out[l] = code[l]; offs_out[l] = offs[l]; }
    for (int s = 0; s < n; s++) {
        const int l = lens[s];
        if (!l) continue;
        const uint32_t c = code[l]++;
        sym[offs[l]++] = (uint16_t)s;
        if (l <= fbits) {
            const uint16_t e = (uint16_t)(s | (l << 9));
            for (uint32_t k = bitrev(c, l); k < (1u << fbits); k += (1u << l)) fast[k] = e;
        }
    }
    return OK;
}

// one bit at a time (codes longer than the fast table): returns the symbol or -1; *nbits = code length
WGBS_HD int slow_decode(uint64_t bb, const uint16_t *cnt, const uint16_t *sym, int *nbits) {
    int code = 0, first = 0, index = 0;
    for (int len = 1; len <= 15; len++) {
        code |= (int)((bb >> (len - 1)) & 1);
        const int count = cnt[len];
        if (code - count < first) { *nbits = len; return sym[index + (code - first)]; }
        index += count; first += count; first <<= 1; code <<= 1;
    }
    return -1;
}

// The same for codes longer than the fast table, without the bit-by-bit walk: the codes of length L are the numbers
// first[L] .. first[L]+cnt[L]-1, and the L-bit prefix of any longer code is larger than all of them, so the first L for
// which the (bit-reversed) next L stream bits fall into that range is the code length.  v: next 32 stream bits, LSB first.
WGBS_HD int long_decode(uint32_t v, const uint16_t *cnt, const uint16_t *first, const uint16_t *offs, const uint16_t *sym, int fbits, int *nbits) {
    const uint32_t r = brev32(v);
    for (int L = fbits + 1; L <= 15; L++) {
        const uint32_t d = (r >> (32 - L)) - first[L];
        if (d < cnt[L]) { *nbits = L; return sym[offs[L] + d]; }
    }
    return -1;
}

// ---- bit reader (lane 0 only) ------------------------------------------------------------------------------------------
struct Bits {
    const uint8_t *src; uint32_t p, end;
    uint64_t bb; int bc;
    WGBS_HD void init(const uint8_t *s, uint32_t n) { src = s; p = 0; end = n; bb = 0; bc = 0; }
    // at least 32 valid bits afterwards unless the input is exhausted (then the missing bits read as 0 and `take` notices)
    WGBS_HD void refill() {
        if (bc > 32) return;
        if (((uintptr_t)(src + p) & 3) == 0 && p + 4 <= end) {
            const uint32_t w = *(const uint32_t *)(src + p);
            bb |= (uint64_t)w << bc; bc += 32; p += 4;
        } else {
            while (bc <= 56 && p < end) { bb |= (uint64_t)src[p++] << bc; bc += 8; }
        }
    }
    WGBS_HD bool take(int n, uint32_t *v) {        // n <= 32
        if (bc < n) return false;
        *v = (uint32_t)(bb & ((1ull << n) - 1)); bb >>= n; bc -= n;
        return true;
    }
    WGBS_HD bool drop(int n) { if (bc < n) return false; bb >>= n; bc -= n; return true; }
};

template <class L>
struct Inflater {
    L lanes;
    Scratch *S;
    uint8_t *dst; uint32_t dst_len;

    // fixed code of RFC 1951 3.2.6
    WGBS_HD int fixed_tables() {
        uint8_t *l = S->lens;
        for (int i = 0; i < 144; i++) l[i] = 8;
        for (int i = 144; i < 256; i++) l[i] = 9;
        for (int i = 256; i < 280; i++) l[i] = 7;
        for (int i = 280; i < 288; i++) l[i] = 8;
        int rc = build_table(l, 288, S->lcnt, S->lsym, S->lt, LBITS, true);
        if (rc) return rc;
        for (int i = 0; i < 30; i++) l[i] = 5;
        // 30 five-bit codes leave two slots unused: incomplete by construction, accepted like zlib's fixed table
        for (int i = 0; i < 16; i++) S->dcnt[i] = 0;
        S->dcnt[5] = 30;
        for (int i = 0; i < (1 << DBITS); i++) S->dt[i] = 0;
        for (int s = 0; s < 30; s++) {
            S->dsym[s] = (uint16_t)s;
            for (uint32_t k = bitrev((uint32_t)s, 5); k < (1u << DBITS); k += 32) S->dt[k] = (uint16_t)(s | (5 << 9));
        }
        return OK;
    }

    // dynamic code of RFC 1951 3.2.7
    WGBS_HD int dynamic_tables(Bits &B) {
        uint32_t v;
        B.refill();
        if (!B.take(14, &v)) return E_INPUT;
        const int nlen = (int)(v & 31) + 257, ndist = (int)((v >> 5) & 31) + 1, ncode = (int)((v >> 10) & 15) + 4;
        if (nlen > 286 || ndist > 30) return E_CODES;
        const char *order = "\x10\x11\x12\x00\x08\x07\x09\x06\x0a\x05\x0b\x04\x0c\x03\x0d\x02\x0e\x01\x0f";
        uint8_t *l = S->lens;
        for (int i = 0; i < 19; i++) l[i] = 0;
        for (int i = 0; i < ncode; i++) { B.refill(); if (!B.take(3, &v)) return E_INPUT; l[(int)order[i]] = (uint8_t)v; }
        // the code-length code borrows the distance tables (7-bit codes fit the 8-bit fast table)
        int rc = build_table(l, 19, S->dcnt, S->dsym, S->dt, 7, false);
        if (rc) return rc;
        uint16_t clt[128], ccnt[16], csym[19];
        for (int i = 0; i < 128; i++) clt[i] = S->dt[i];
        for (int i = 0; i < 16; i++) ccnt[i] = S->dcnt[i];
        for (int i = 0; i < 19; i++) csym[i] = S->dsym[i];
        int idx = 0;
        while (idx < nlen + ndist) {
            B.refill();
            const uint16_t e = clt[B.bb & 127];
            int s, nb = e >> 9;
            if (nb) s = e & 511; else { s = slow_decode(B.bb, ccnt, csym, &nb); if (s < 0) return E_SYMBOL; }
            if (!B.drop(nb)) return E_INPUT;
            if (s < 16) { l[idx++] = (uint8_t)s; continue; }
            int rep; uint8_t val = 0;
            if (s == 16) { if (idx == 0) return E_CODES; val = l[idx - 1]; if (!B.take(2, &v)) return E_INPUT; rep = 3 + (int)v; }
            else if (s == 17) { if (!B.take(3, &v)) return E_INPUT; rep = 3 + (int)v; }
            else { if (!B.take(7, &v)) return E_INPUT; rep = 11 + (int)v; }
            if (idx + rep > nlen + ndist) return E_CODES;
            while (rep--) l[idx++] = val;
        }
        if (l[256] == 0) return E_CODES;                 // no end-of-block code
        rc = build_table(l, nlen, S->lcnt, S->lsym, S->lt, LBITS, true);
        if (rc) return rc;
        return build_table(l + nlen, ndist, S->dcnt, S->dsym, S->dt, DBITS, true);
    }

    // lane 0: decode up to `room` symbols of the current Huffman block into S->q.  *eob is set at the end-of-block code.
    WGBS_HD int decode_some(Bits &B, int room, int *nq, bool *eob) {
        int n = 0;
        while (n < room) {
            B.refill();
            uint16_t e = S->lt[B.bb & ((1u << LBITS) - 1)];
            int s, nb = e >> 9;
            if (nb) s = e & 511; else { s = slow_decode(B.bb, S->lcnt, S->lsym, &nb); if (s < 0) return E_SYMBOL; }
            if (!B.drop(nb)) return E_INPUT;
            if (s < 256) { S->q[n++] = (uint32_t)s; continue; }
            if (s == 256) { *eob = true; break; }
            if (s > 285) return E_SYMBOL;
            uint32_t len, v;
            if (s < 265) len = (uint32_t)(s - 254);
            else if (s == 285) len = 258;
            else { const int eb = ((s - 265) >> 2) + 1; if (!B.take(eb, &v)) return E_INPUT; len = 3 + ((4u + (uint32_t)((s - 265) & 3)) << eb) + v; }
            B.refill();
            e = S->dt[B.bb & ((1u << DBITS) - 1)];
            nb = e >> 9;
            if (nb) s = e & 511; else { s = slow_decode(B.bb, S->dcnt, S->dsym, &nb); if (s < 0) return E_SYMBOL; }
            if (!B.drop(nb)) return E_INPUT;
            if (s > 29) return E_SYMBOL;
            uint32_t dist;
            if (s < 4) dist = (uint32_t)s + 1;
            else { const int eb = (s >> 1) - 1; if (!B.take(eb, &v)) return E_INPUT; dist = 1 + ((2u + (uint32_t)(s & 1)) << eb) + v; }
            S->q[n++] = 0x80000000u | ((dist - 1) << 9) | len;
        }
        *nq = n;
        return OK;
    }

    // all lanes: write the queued symbols at *opos.  Literals first (independent), then the matches in stream order:
    // everything a match reads lies before its own output position, i.e. was produced by an earlier symbol.
    WGBS_HD int emit(int nq, uint32_t *opos) {
        const int l = lanes.id();
        const uint32_t e = l < nq ? S->q[l] : 0;
        const bool live = l < nq, is_match = live && (e >> 31);
        const uint32_t mylen = !live ? 0u : (is_match ? (e & 511u) : 1u);
        uint32_t total;
        const uint32_t at = *opos + lanes.exscan(mylen, &total, mylen);
        if (*opos + total > dst_len) return E_OUTPUT;
        if (live && !is_match) dst[at] = (uint8_t)e;
        uint32_t m = lanes.ballot(is_match);
        int bad = 0;
        lanes.sync();
        while (m) {
            int i = 0; while (!((m >> i) & 1)) i++;
            m &= m - 1;
            const uint32_t ee = lanes.shfl(e, i), p = lanes.shfl(at, i);
            const uint32_t len = ee & 511u, dist = ((ee >> 9) & 0xffffu) + 1;
            if (dist > p) { bad = 1; break; }
            const uint8_t *from = dst + p - dist;
            if (dist >= len) { for (uint32_t k = (uint32_t)l; k < len; k += L::N) dst[p + k] = from[k]; }
            else { for (uint32_t k = (uint32_t)l; k < len; k += L::N) dst[p + k] = from[k % dist]; }
            lanes.sync();
        }
        if (bad) return E_DIST;
        *opos += total;
        return OK;
    }

    // Inflate `src[0..src_len)` into `dst[0..dst_len)`; the stream must produce exactly dst_len bytes (the block's ISIZE).
    WGBS_HD int run(const uint8_t *src, uint32_t src_len) {
        Bits B; B.init(src, src_len);
        uint32_t opos = 0;
        const bool lead = lanes.id() == 0;
        int rc = OK;
        bool last = false;
        while (!last && rc == OK) {
            // ---- block header (lane 0), broadcast: type, final flag, stored length / source offset
            uint32_t hdr = 0, slen = 0, sfrom = 0;
            if (lead) {
                uint32_t v;
                B.refill();
                if (!B.take(3, &v)) rc = E_INPUT;
                else {
                    hdr = v;
                    const uint32_t type = v >> 1;
                    if (type == 0) {
                        B.drop(B.bc & 7);                                   // to the byte boundary
                        B.refill();
                        if (!B.take(32, &v)) rc = E_INPUT;
                        else if ((v & 0xffffu) != ((~v >> 16) & 0xffffu)) rc = E_STORED;
                        else {
                            slen = v & 0xffffu;
                            sfrom = B.p - (uint32_t)(B.bc >> 3);            // bytes still parked in the bit buffer belong to the payload
                            if (sfrom + slen > B.end) rc = E_INPUT;
                            else { B.p = sfrom + slen; B.bb = 0; B.bc = 0; }
                        }
                    } else if (type == 1) rc = fixed_tables();
                    else if (type == 2) rc = dynamic_tables(B);
                    else rc = E_BTYPE;
                }
            }
            rc = (int)lanes.shfl((uint32_t)rc, 0);
            if (rc != OK) break;
            hdr = lanes.shfl(hdr, 0);
            last = hdr & 1;
            if ((hdr >> 1) == 0) {
                slen = lanes.shfl(slen, 0); sfrom = lanes.shfl(sfrom, 0);
                if (opos + slen > dst_len) { rc = E_OUTPUT; break; }
                for (uint32_t k = (uint32_t)lanes.id(); k < slen; k += L::N) dst[opos + k] = src[sfrom + k];
                opos += slen;
                lanes.sync();
                continue;
            }
            lanes.sync();                                                   // tables visible (only lane 0 reads them; cheap)
            bool eob = false;
            while (!eob && rc == OK) {
                int nq = 0;
                if (lead) rc = decode_some(B, L::N, &nq, &eob);
                lanes.sync();
                rc = (int)lanes.shfl((uint32_t)rc, 0);
                if (rc != OK) break;
                nq = (int)lanes.shfl((uint32_t)nq, 0);
                eob = lanes.shfl(eob ? 1u : 0u, 0) != 0;
                if (nq) rc = emit(nq, &opos);
                lanes.sync();
            }
        }
        if (rc == OK && opos != dst_len) rc = E_SHORT;
        return rc;
    }
};

// ---- second decoder: every lane runs the SAME instruction stream ------------------------------------------------------------
// The first decoder spends one lane on the Huffman walk inside a divergent region (convergence barriers, a 64-bit bit
// buffer refilled from global memory, a shared-memory symbol queue): ~46 warp-instructions per symbol, and the kernel is
// issue-bound (measured: 10.3 ms for the 3 000 blocks of a 1M-read batch).  Here
//   * the compressed bytes are staged in a shared-memory ring by coalesced loads (a word index, no refill branches),
//   * a symbol costs two ring words + one funnel shift + one table probe, executed identically by all lanes (loads are
//     broadcasts, control flow is uniform: no convergence barriers),
//   * symbol i of a batch stays in lane i's register -- no queue in shared memory;
// the emit step (literals in parallel, matches in stream order) is the same.
constexpr uint32_t RING_WORDS = 512;            // 2 KiB of compressed stream per warp
constexpr uint32_t RING_KEEP = 480;             // words loaded ahead of the read position by a refill
constexpr uint32_t OWN_MAX = 16;                // matches up to this length are copied by the lane that decoded them
struct Ring { uint32_t w[RING_WORDS + 1]; };    // slot RING_WORDS mirrors slot 0: the word after any slot is at +1, no wrap

WGBS_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, sh);
#else
    return (uint32_t)(((((uint64_t)hi) << 32) | lo) >> (sh & 31));
#endif
}

template <class L>
struct Inflater2 {
    L lanes;
    Scratch *S;
    Ring *R;
    uint8_t *dst; uint32_t dst_len;
    // compressed stream: 32-bit words at gw[0 .. nwords), first payload bit = bit0, one past the last payload bit = end_bit
    const uint32_t *gw; uint32_t nwords, end_bit;
    uint32_t hi_w;                                  // ring holds words [.., hi_w)
    uint32_t bitpos;

    // make the ring hold at least `ahead` words from the current position on (uniform; ends with a warp barrier)
    WGBS_HD void refill(uint32_t ahead) {
        const uint32_t cur = bitpos >> 5;
        if (hi_w >= cur + ahead) return;
        if (hi_w < cur) hi_w = cur;                  // jumped over everything staged (stored block)
        const uint32_t target = cur + RING_KEEP;
        lanes.sync();                                // nobody still reads the slots about to be overwritten
        for (uint32_t w = hi_w + (uint32_t)lanes.id(); w < target; w += L::N) {
            const uint32_t x = w < nwords ? gw[w] : 0u, slot = w & (RING_WORDS - 1);
            R->w[slot] = x;
            if (slot == 0) R->w[RING_WORDS] = x;
        }
        hi_w = target;
        lanes.sync();
    }
    WGBS_HD uint32_t peek() const {                  // the next 32 bits of the stream
        const uint32_t *p = R->w + ((bitpos >> 5) & (RING_WORDS - 1));
        return funnel_r(p[0], p[1], bitpos & 31);
    }

    WGBS_HD int fixed_tables() {
        uint8_t *l = S->lens;
        for (int i = 0; i < 144; i++) l[i] = 8;
        for (int i = 144; i < 256; i++) l[i] = 9;
        for (int i = 256; i < 280; i++) l[i] = 7;
        for (int i = 280; i < 288; i++) l[i] = 8;
        int rc = build_table(l, 288, S->lcnt, S->lsym, S->lt, LBITS, true, S->lfirst, S->loffs);
        if (rc) return rc;
        for (int i = 0; i < 16; i++) S->dcnt[i] = 0;
        S->dcnt[5] = 30;
        for (int i = 0; i < 16; i++) { S->dfirst[i] = 0; S->doffs[i] = 0; }
        for (int i = 0; i < (1 << DBITS); i++) S->dt[i] = 0;
        for (int s = 0; s < 30; s++) {
            S->dsym[s] = (uint16_t)s;
            for (uint32_t k = bitrev((uint32_t)s, 5); k < (1u << DBITS); k += 32) S->dt[k] = (uint16_t)(s | (5 << 9));
        }
        return OK;
    }

    // one lane: code lengths of a dynamic block (the ring holds >= 300 words ahead: the longest header is < 600 bytes)
    WGBS_HD int dynamic_tables(uint32_t *bp) {
        uint32_t b = *bp;
        auto window = [&](uint32_t at) { const uint32_t *p = R->w + ((at >> 5) & (RING_WORDS - 1)); return funnel_r(p[0], p[1], at & 31); };
        uint32_t v = window(b); b += 14;
        const int nlen = (int)(v & 31) + 257, ndist = (int)((v >> 5) & 31) + 1, ncode = (int)((v >> 10) & 15) + 4;
        if (nlen > 286 || ndist > 30) return E_CODES;
        const char *order = "\x10\x11\x12\x00\x08\x07\x09\x06\x0a\x05\x0b\x04\x0c\x03\x0d\x02\x0e\x01\x0f";
        uint8_t *l = S->lens;
        for (int i = 0; i < 19; i++) l[i] = 0;
        for (int i = 0; i < ncode; i++) { l[(int)order[i]] = (uint8_t)(window(b) & 7); b += 3; }
        int rc = build_table(l, 19, S->dcnt, S->dsym, S->dt, 7, false);
        if (rc) return rc;
        uint16_t clt[128], ccnt[16], csym[19];
        for (int i = 0; i < 128; i++) clt[i] = S->dt[i];
        for (int i = 0; i < 16; i++) ccnt[i] = S->dcnt[i];
        for (int i = 0; i < 19; i++) csym[i] = S->dsym[i];
        int idx = 0;
        while (idx < nlen + ndist) {
            if (b > end_bit) return E_INPUT;
            v = window(b);
            const uint16_t e = clt[v & 127];
            int s, nb = e >> 9;
            if (nb) s = e & 511; else { s = slow_decode(v, ccnt, csym, &nb); if (s < 0) return E_SYMBOL; }
            b += (uint32_t)nb; v >>= nb;
            if (s < 16) { l[idx++] = (uint8_t)s; continue; }
            int rep; uint8_t val = 0;
            if (s == 16) { if (idx == 0) return E_CODES; val = l[idx - 1]; rep = 3 + (int)(v & 3); b += 2; }
            else if (s == 17) { rep = 3 + (int)(v & 7); b += 3; }
            else { rep = 11 + (int)(v & 127); b += 7; }
            if (idx + rep > nlen + ndist) return E_CODES;
            while (rep--) l[idx++] = val;
        }
        if (b > end_bit) return E_INPUT;
        if (l[256] == 0) return E_CODES;
        rc = build_table(l, nlen, S->lcnt, S->lsym, S->lt, LBITS, true, S->lfirst, S->loffs);
        if (rc) return rc;
        *bp = b;
        return build_table(l + nlen, ndist, S->dcnt, S->dsym, S->dt, DBITS, true, S->dfirst, S->doffs);
    }

    // all lanes, same path: decode up to L::N symbols; lane i keeps symbol i in *mine.  Needs >= 64 ring words ahead.
    WGBS_HD int decode_batch(uint32_t *mine, int *nq, bool *eob) {
        int n = 0; uint32_t my = 0;
        const int me = lanes.id();
        bool bad = false;                                                // invalid code / symbol: noted, reported after the batch
        while (n < L::N) {                                               // (no early exits in the loop: one uniform path)
            uint32_t v = peek();
            uint32_t e = S->lt[v & ((1u << LBITS) - 1)];
            int s, nb = (int)(e >> 9);
            if (nb) s = (int)(e & 511); else { s = long_decode(v, S->lcnt, S->lfirst, S->loffs, S->lsym, LBITS, &nb); if (s < 0) { bad = true; s = 256; nb = 0; } }
            bitpos += (uint32_t)nb;
            if (s < 256) { if (me == n) my = (uint32_t)s; n++; continue; }
            if (s == 256) { *eob = true; break; }
            if (s > 285) { bad = true; *eob = true; break; }
            v >>= nb;                                                   // code <= 15 bits + extra <= 5 bits: still inside the window
            uint32_t len;
            if (s < 265) len = (uint32_t)(s - 254);
            else if (s == 285) len = 258;
            else { const int eb = ((s - 265) >> 2) + 1; len = 3 + ((4u + (uint32_t)((s - 265) & 3)) << eb) + (v & ((1u << eb) - 1)); bitpos += (uint32_t)eb; }
            v = peek();
            e = S->dt[v & ((1u << DBITS) - 1)];
            nb = (int)(e >> 9);
            if (nb) s = (int)(e & 511); else { s = long_decode(v, S->dcnt, S->dfirst, S->doffs, S->dsym, DBITS, &nb); if (s < 0) { bad = true; s = 0; nb = 0; } }
            if (s > 29) { bad = true; s = 0; }
            bitpos += (uint32_t)nb; v >>= nb;
            uint32_t dist;
            if (s < 4) dist = (uint32_t)s + 1;
            else { const int eb = (s >> 1) - 1; dist = 1 + ((2u + (uint32_t)(s & 1)) << eb) + (v & ((1u << eb) - 1)); bitpos += (uint32_t)eb; }
            if (me == n) my = 0x80000000u | ((dist - 1) << 9) | len;
            n++;
        }
        if (bad) return E_SYMBOL;
        if (bitpos > end_bit) return E_INPUT;                          // some symbol of this batch read past the payload
        *mine = my; *nq = n;
        return OK;
    }

    // Write the batch at *opos.  Half of the symbols of a BAM stream are matches (measured: 4 000 literals + 4 000 matches of
    // ~15 bytes per 64 KiB block), so copying them one after the other -- a dependent L2 round trip each -- is what the
    // first decoder spends most of its time on.  Here a match whose source lies entirely BEFORE this batch's output
    // (93 % of them) is copied by its own lane, all lanes at once; only matches that read what this very batch produces
    // (short distances, runs) and long ones go through the in-order cooperative loop afterwards.
    WGBS_HD int emit(uint32_t e, int nq, uint32_t *opos) {
        const int l = lanes.id();
        const uint32_t base = *opos;
        const bool live = l < nq, is_match = live && (e >> 31);
        const uint32_t len = e & 511u, dist = ((e >> 9) & 0xffffu) + 1;
        const uint32_t mylen = !live ? 0u : (is_match ? len : 1u);
        uint32_t total;
        const uint32_t at = base + lanes.exscan(mylen, &total, mylen);
        if (base + total > dst_len) return E_OUTPUT;
        if (lanes.ballot(is_match && dist > at)) return E_DIST;
        // own: source entirely before this batch AND short (6 % of the matches are 80..258 bytes long and carry two thirds of
        // the bytes: one lane copying such a match byte by byte stalls the other 31, so long ones are copied by the whole warp)
        const bool own = is_match && (at - dist + len <= base) && len <= OWN_MAX;
        lanes.sync();                                                   // everything written by earlier batches is visible
        if (live && !is_match) dst[at] = (uint8_t)e;
        if (own) { const uint8_t *from = dst + at - dist; for (uint32_t k = 0; k < len; k++) dst[at + k] = from[k]; }
        uint32_t m = lanes.ballot(is_match && !own);
        if (m) {
            lanes.sync();
            while (m) {
                const int i = lowest_bit(m);
                m &= m - 1;
                const uint32_t ee = lanes.shfl(e, i), p = lanes.shfl(at, i);
                const uint32_t ln = ee & 511u, ds = ((ee >> 9) & 0xffffu) + 1;
                const uint8_t *from = dst + p - ds;
                if (ds >= ln) { for (uint32_t k = (uint32_t)l; k < ln; k += L::N) dst[p + k] = from[k]; }
                else { for (uint32_t k = (uint32_t)l; k < ln; k += L::N) dst[p + k] = from[k % ds]; }
                lanes.sync();
            }
        }
        *opos += total;
        return OK;
    }

    // src / src_len: the deflate payload; the words around it must be readable (the caller pads the buffer)
    WGBS_HD int run(const uint8_t *src, uint32_t src_len) {
        const uint32_t mis = (uint32_t)((uintptr_t)src & 3);
        gw = (const uint32_t *)(src - mis);
        bitpos = 8 * mis; end_bit = 8 * (mis + src_len); nwords = (mis + src_len + 3) >> 2;
        hi_w = 0;
        uint32_t opos = 0;
        int rc = OK;
        bool last = false;
        while (!last && rc == OK) {
            refill(300);
            if (bitpos + 3 > end_bit) { rc = E_INPUT; break; }
            const uint32_t hdr = peek() & 7; bitpos += 3;
            last = hdr & 1;
            const uint32_t type = hdr >> 1;
            if (type == 0) {
                bitpos = (bitpos + 7) & ~7u;
                if (bitpos + 32 > end_bit) { rc = E_INPUT; break; }
                const uint32_t v = peek(); bitpos += 32;
                if ((v & 0xffffu) != ((~v >> 16) & 0xffffu)) { rc = E_STORED; break; }
                const uint32_t slen = v & 0xffffu, from = (bitpos >> 3) - mis;
                if (from + slen > src_len) { rc = E_INPUT; break; }
                if (opos + slen > dst_len) { rc = E_OUTPUT; break; }
                for (uint32_t k = (uint32_t)lanes.id(); k < slen; k += L::N) dst[opos + k] = src[from + k];
                opos += slen; bitpos += 8 * slen;
                lanes.sync();
                continue;
            }
            if (type == 3) { rc = E_BTYPE; break; }
            // tables: one lane builds them (a few thousand instructions per deflate block), the position is broadcast
            uint32_t bp = bitpos;
            if (lanes.id() == 0) rc = type == 1 ? fixed_tables() : dynamic_tables(&bp);
            lanes.sync();
            rc = (int)lanes.shfl((uint32_t)rc, 0);
            if (rc != OK) break;
            bitpos = lanes.shfl(bp, 0);
            bool eob = false;
            while (!eob) {
                refill(64);
                uint32_t mine = 0; int nq = 0;
                rc = decode_batch(&mine, &nq, &eob);
                if (rc != OK) break;
                if (nq) { rc = emit(mine, nq, &opos); if (rc != OK) break; }
            }
        }
        if (rc == OK && opos != dst_len) rc = E_SHORT;
        return rc;
    }
};

// ---- CRC-32 of a block's output (gzip trailer; htslib checks it in bgzf_read_block -> check_header/inflate_block) -----------
// Reflected CRC-32 (polynomial 0xEDB88320).  Every lane runs the byte-wise table update over its own contiguous slice from a
// zero register; the slices are then chained: register(A||B) = register(A) * x^(8|B|) mod P  xor  register_0(B).
constexpr uint32_t CRC_POLY = 0xEDB88320u;
WGBS_HD uint32_t crc_table_entry(uint32_t i) { uint32_t c = i; for (int k = 0; k < 8; k++) c = (c & 1) ? (c >> 1) ^ CRC_POLY : c >> 1; return c; }
// a(x) * b(x) mod P in the reflected representation (bit 31 = x^0)
WGBS_HD uint32_t crc_mulmod(uint32_t a, uint32_t b) {
    uint32_t p = 0;
    for (uint32_t m = 1u << 31; m; m >>= 1) {
        if (a & m) p ^= b;
        b = (b & 1) ? (b >> 1) ^ CRC_POLY : b >> 1;
    }
    return p;
}
// x^(8n) mod P
WGBS_HD uint32_t crc_x8n(uint32_t n) {
    uint32_t r = 1u << 31, base = 1u << 30;       // x^0, x^1
    for (uint64_t e = 8ull * n; e; e >>= 1) { if (e & 1) r = crc_mulmod(r, base); base = crc_mulmod(base, base); }
    return r;
}
template <class L>
WGBS_HD uint32_t crc32_block(L lanes, const uint8_t *d, uint32_t n, const uint32_t *table /* 256 entries */) {
    const uint32_t per = (n + L::N - 1) / L::N;
    const uint32_t l = (uint32_t)lanes.id();
    const uint32_t a = l * per < n ? l * per : n, b = a + per < n ? a + per : n;
    uint32_t c = 0;
    for (uint32_t i = a; i < b; i++) c = table[(c ^ d[i]) & 0xff] ^ (c >> 8);
    // operators for a full slice and for the (shorter) last one: computed by lanes 0 / 1, broadcast
    const uint32_t nfull = per ? n / per : 0, rem = per ? n - nfull * per : 0;
    uint32_t op = 0;
    if (l == 0) op = crc_x8n(per);
    if (L::N > 1 && l == 1) op = crc_x8n(rem);
    const uint32_t x_full = lanes.shfl(op, 0), x_rem = L::N > 1 ? lanes.shfl(op, 1) : crc_x8n(rem);
    uint32_t t = 0xffffffffu;
    for (int i = 0; i < L::N; i++) {
        const uint32_t ci = lanes.shfl(c, i);
        const uint32_t len_i = (uint32_t)i < nfull ? per : ((uint32_t)i == nfull ? rem : 0);
        if (len_i == 0) continue;
        t = crc_mulmod(t, len_i == per ? x_full : x_rem) ^ ci;
    }
    return t ^ 0xffffffffu;
}

// ---- BGZF block framing (SAM spec 4.1) -----------------------------------------------------------------------------------
// header: 1f 8b 08 04 | mtime(4) xfl os | xlen(2) | subfields ... 'B' 'C' 02 00 bsize-1(2) ... | deflate | crc32(4) isize(4)
// returns the total block size (0: not a BGZF block header); *xlen_out = length of the extra field
WGBS_HD uint32_t bgzf_block_size(const uint8_t *h, uint64_t avail, uint32_t *xlen_out) {
    if (avail < 18 || h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) return 0;
    const uint32_t xlen = (uint32_t)h[10] | ((uint32_t)h[11] << 8);
    if (12ull + xlen > avail) return 0;
    for (uint32_t x = 0; x + 4 <= xlen;) {
        const uint8_t *sf = h + 12 + x;
        const uint32_t sl = (uint32_t)sf[2] | ((uint32_t)sf[3] << 8);
        if (sf[0] == 'B' && sf[1] == 'C' && sl == 2 && x + 6 <= xlen) { *xlen_out = xlen; return ((uint32_t)sf[4] | ((uint32_t)sf[5] << 8)) + 1u; }
        x += 4 + sl;
    }
    return 0;
}

}  // namespace dflate
