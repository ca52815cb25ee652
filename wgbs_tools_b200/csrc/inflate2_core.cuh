// inflate2_core.cuh -- raw DEFLATE (RFC 1951) for BGZF blocks in two phases: the pieces shared by the team decoder (inflate3_core.cuh).
//
//   phase 1  walks the Huffman stream: literals go to their final place at once, a match becomes an 8-byte token
//            (output position, length, distance) in a per-block list.  Here: the table entries (32 bits, everything precomputed:
//            code bits + extra bits, base value, kind; two-level tables in one arena), the tokens, and `Decoder` -- the bit reader
//            of ONE lane that parses deflate block headers (stored blocks, code lengths) and can build the tables serially.  The
//            walk itself, by the 32 lanes of a team, is inflate3_core.cuh.
//   phase 2  (resolve_bytes) one WARP per block replays the token list: per round the matches whose sources are final are laid end
//            to end and the lanes copy consecutive BYTES; then ISIZE and CRC32 (crc32_block4).
//
// History (numbers in DESIGN.md section 7 / profiles/README.md): round 1 decoded a block with one warp, every lane the same walk
// (inflate_core.cuh, 4.08 ms per 1M-read BAM; still the fallback for blocks whose tables exceed the arena); the first decoder of
// round 2 walked one block per THREAD with lane-interleaved tables (1.74 ms) and replayed one match per lane (1.28 ms); both
// were removed when the team decoder (0.94 ms) and the byte-per-lane replay replaced them.
//
// Host + device code: tests/bamdev_core_check.cpp builds it with g++ and pins it against zlib.  Stands in for the zlib inflate of
// the reference's `samtools view` stage (reference src/python/bam2pat.py:165).
#pragma once
#include "inflate_core.cuh"

#if defined(__GNUC__) || defined(__CUDACC__)
#define WGBS_UNLIKELY(x) __builtin_expect(!!(x), 0)
#else
#define WGBS_UNLIKELY(x) (x)
#endif

namespace dflate2 {

using dflate::OK; using dflate::E_INPUT; using dflate::E_BTYPE; using dflate::E_STORED; using dflate::E_CODES; using dflate::E_SYMBOL;
using dflate::E_DIST; using dflate::E_OUTPUT; using dflate::E_SHORT; using dflate::E_CRC;

constexpr int LB = 10;      // literal/length codes up to this many bits decode with one probe
constexpr int DB = 8;       // distance codes

// ---- table entries -----------------------------------------------------------------------------------------------------
// bits 0..4   code bits + extra bits (what the probe consumes; 0: not in the fast table)
// bits 5..8   extra bits
// bits 9..10  kind
// bits 16..31 literal byte / base of the length or distance
enum : uint32_t { K_LIT = 0, K_BASE = 1, K_EOB = 2, K_BAD = 3 };
WGBS_HD uint32_t mk_entry(uint32_t nb, uint32_t eb, uint32_t kind, uint32_t val) { return (nb + eb) | (eb << 5) | (kind << 9) | (val << 16); }
WGBS_HD uint32_t litlen_entry(uint32_t s, uint32_t nb) {
    if (s < 256) return mk_entry(nb, 0, K_LIT, s);
    if (s == 256) return mk_entry(nb, 0, K_EOB, 0);
    if (s > 285) return mk_entry(nb, 0, K_BAD, 0);
    if (s < 265) return mk_entry(nb, 0, K_BASE, s - 254);
    if (s == 285) return mk_entry(nb, 0, K_BASE, 258);
    const uint32_t eb = ((s - 265) >> 2) + 1;
    return mk_entry(nb, eb, K_BASE, 3 + ((4u + ((s - 265) & 3)) << eb));
}
WGBS_HD uint32_t dist_entry(uint32_t s, uint32_t nb) {
    if (s > 29) return mk_entry(nb, 0, K_BAD, 0);
    if (s < 4) return mk_entry(nb, 0, K_BASE, s + 1);
    const uint32_t eb = (s >> 1) - 1;
    return mk_entry(nb, eb, K_BASE, 1 + ((2u + (s & 1)) << eb));
}
WGBS_HD uint32_t clen_entry(uint32_t s, uint32_t nb) { return mk_entry(nb, 0, K_LIT, s); }

// ---- tokens ------------------------------------------------------------------------------------------------------------
// x = output position | length << 16;  y = distance (match), or TOK_STORED | offset of the bytes in the block's payload
struct alignas(8) Token { uint32_t x, y; };
constexpr uint32_t TOK_STORED = 0x80000000u;
constexpr uint32_t MAX_STORED_TOKENS = 16;   // further stored blocks of one BGZF block are copied by the decoding lane itself
WGBS_HD uint32_t token_cap(uint32_t usize) { return usize / 3 + 2 + MAX_STORED_TOKENS; }

// ---- working memory of one block ----------------------------------------------------------------------------------------------
// element i of an array lives at p[i << SHIFT] (SHIFT 0: plain arrays; the round-2 thread-per-block decoder interleaved the arrays
// of 32 lanes with SHIFT 5).
//
// Tables: ONE arena of ARENA entries.  [0, 1 << LB) is the literal/length root table; second-level tables of the literal/length
// codes longer than LB bits follow; then the distance root table (1 << DB entries, at dt_off) and its second-level tables.  A root
// entry with (e & 31) == 0 is indirect: bits 5..8 = index bits of the second-level table, bits 16.. = its offset.
// The worst case of RFC 1951 code sets needs ~1750 entries; real streams need 1300-1400.  A block whose tables do not fit is
// reported as E_FALLBACK and decoded by the warp-per-block decoder of inflate_core.cuh instead.
constexpr uint32_t ARENA = 1536;
constexpr uint32_t RING = 128;         // words of compressed stream staged for the header parser: 32 chunks of 16 bytes
constexpr int E_FALLBACK = -20;        // not an error of the stream: the tables of this block need more than ARENA entries
template <int SHIFT>
struct Mem {
    uint32_t *tab;                // ARENA entries
    uint32_t *rg;                 // RING words
    uint16_t *bk;                 // 32 words of bookkeeping for the table construction (count / cursor per code length)
    uint8_t *ln;                  // 320 code lengths
    static constexpr int shift = SHIFT;
};
struct HostLane {
    uint32_t tab[ARENA], rg[RING];
    uint16_t bk[32];
    uint8_t ln[320];
    WGBS_HD Mem<0> mem() { Mem<0> m; m.tab = tab; m.rg = rg; m.bk = bk; m.ln = ln; return m; }
};

enum : int { ST_HDR = 0 /* at a deflate block header */, ST_DEC = 1 /* inside a Huffman block */, ST_DONE = 2 };

template <int SHIFT>
struct Decoder {
    Mem<SHIFT> m;
    // compressed stream: 32-bit words gw[0 .. nwords); payload bits [.., end_bit) counted from gw
    const uint32_t *gw; uint32_t nwords, end_bit;
    const uint8_t *src; uint32_t mis;          // payload bytes (stored blocks); src - mis = (const uint8_t *)gw
    uint64_t bb; uint32_t bc;                  // bit buffer: bc valid bits (bits above bc may already hold stream bits: see probe loop)
    uint32_t wp, nw;                           // words consumed into bb; nw = word wp (already fetched from the ring)
    uint32_t hi_c;                             // chunks (4 words) of the stream staged in the ring so far: [.., hi_c)
    // output
    uint8_t *dst; uint32_t dst_len, opos;
    Token *tok; uint32_t ntok, nstored;
    int rc; bool last;
    // decode state
    uint32_t len;                              // pending match length (0: next probe is literal/length)
    uint32_t dt_off;                           // arena offset of the distance root table
    // team decoder (inflate3_core.cuh): header() stops behind the code lengths; the team builds the tables of ln[0 .. pend_nlen + pend_ndist) together
    uint16_t defer, pend_nlen, pend_ndist;

    static constexpr uint32_t S = (uint32_t)SHIFT;

    // ---- ring ----------------------------------------------------------------------------------------------------------
    WGBS_HD uint32_t ring_word(uint32_t k) const {
        const uint32_t s = k & (RING - 1);
        return SHIFT ? m.rg[((s >> 2) << 7) + (s & 3)] : m.rg[s];
    }
    // stage chunk c (words [4c, 4c + 4)) into its ring slot.  Chunks past the stream are skipped.
    WGBS_HD void ring_load_chunk(uint32_t c) {
        const uint32_t w = 4 * c, s = w & (RING - 1);
        if (w >= nwords) return;
        for (uint32_t k = 0; k < 4; k++) m.rg[SHIFT ? ((s >> 2) << 7) + k : s + k] = gw[w + k];      // (the buffer is padded: whole chunks are readable)
    }
    WGBS_HD void ring_commit_wait() {}         // (the thread-per-block decoder staged with cp.async and waited here)
    // make the ring complete up to its capacity, synchronously (after a header: its reads are not paced like a burst's)
    WGBS_HD void ring_fill_sync() {
        ring_commit_wait();
        const uint32_t limit = (wp >> 2) + RING / 4;
        while (hi_c < limit) { ring_load_chunk(hi_c); hi_c++; }
        ring_commit_wait();
    }
    // position the reader at bit `pos` of the stream (synchronous: block start and stored blocks only)
    WGBS_HD void seek(uint32_t pos) {
        wp = pos >> 5;
        ring_commit_wait();                    // nothing in flight may land in a slot after it is reloaded
        hi_c = wp >> 2;
        for (uint32_t k = 0; k < RING / 4; k++) { ring_load_chunk(hi_c); hi_c++; }
        ring_commit_wait();
        bb = 0; bc = 0;
        nw = ring_word(wp);
        refill_careful();
        bb >>= (pos & 31); bc -= (pos & 31);
    }
    WGBS_HD uint32_t bitpos() const { return 32 * wp - bc; }
    // append one word (bc <= 32 before).  Careful flavour (headers): stages more of the stream when the ring runs dry.
    WGBS_HD void refill_careful() {
        bb = (bb & ((bc ? (1ull << bc) : 1ull) - 1)) | ((uint64_t)nw << bc); bc += 32; wp++;
        if (wp >= 4 * hi_c) { ring_commit_wait(); for (int k = 0; k < 8; k++) { ring_load_chunk(hi_c); hi_c++; } ring_commit_wait(); }
        nw = ring_word(wp);
    }
    WGBS_HD uint32_t take(uint32_t n) {        // n <= 32; refills as needed (careful flavour)
        if (bc < n) refill_careful();
        const uint32_t v = (uint32_t)(bb & ((1ull << n) - 1)); bb >>= n; bc -= n; return v;
    }

    WGBS_HD void init(Mem<SHIFT> mem, const uint8_t *payload, uint32_t clen, uint8_t *out, uint32_t usize, Token *tokens) {
        m = mem;
        src = payload; mis = (uint32_t)((uintptr_t)payload & 15);
        gw = (const uint32_t *)(payload - mis);
        end_bit = 8 * (mis + clen); nwords = (mis + clen + 3) >> 2;
        dst = out; dst_len = usize; opos = 0; tok = tokens; ntok = 0; nstored = 0;
        rc = OK; last = false; len = 0; dt_off = 1u << LB; defer = 0; pend_nlen = pend_ndist = 0;
        wp = 0; nw = 0; bb = 0; bc = 0; hi_c = 0;
        seek(8 * mis);
    }

    // ---- canonical Huffman tables ----------------------------------------------------------------------------------------
    // lens ln[l0 .. l0 + n) -> root table of 1 << rbits entries at arena offset `root`, second-level tables from *cursor on.
    // which: 0 literal/length, 1 distance, 2 code-length code.  Same accept / reject rules as dflate::build_table (zlib's).
    // sw: arena word where the scratch for the symbols in canonical order starts (n entries of 16 bits); second-level tables stay
    // below `limit`, which keeps them clear of that scratch.
    WGBS_HD uint32_t entry_for(int which, uint32_t s, uint32_t l) const { return which == 0 ? litlen_entry(s, l) : which == 1 ? dist_entry(s, l) : clen_entry(s, l); }
    // symbol i of the canonical-order scratch that starts at arena word sw: two 16-bit entries per arena word
    WGBS_HD uint32_t sorted_get(uint32_t sw, uint32_t i) const { return (m.tab[(sw + (i >> 1)) << S] >> ((i & 1) * 16)) & 0xffffu; }
    WGBS_HD void sorted_set(uint32_t sw, uint32_t i, uint32_t v) {
        uint32_t &w = m.tab[(sw + (i >> 1)) << S];
        w = (i & 1) ? ((w & 0xffffu) | (v << 16)) : ((w & 0xffff0000u) | v);
    }
    WGBS_HD int build(uint32_t l0, int n, uint32_t root, int rbits, uint32_t *cursor, uint32_t limit, int which, uint32_t sw) {
        uint16_t *cnt = m.bk, *nxt = m.bk + (16u << S);
        for (uint32_t i = 0; i < 16; i++) cnt[i << S] = 0;
        for (int i = 0; i < n; i++) cnt[(uint32_t)m.ln[(l0 + (uint32_t)i) << S] << S]++;
        int left = 1, maxlen = 0;
        for (uint32_t l = 1; l <= 15; l++) { left <<= 1; left -= cnt[l << S]; if (left < 0) return E_CODES; if (cnt[l << S]) maxlen = (int)l; }
        if (left > 0 && (maxlen > 1 || which == 2)) return E_CODES;
        uint32_t o = 0;
        for (uint32_t l = 1; l <= 15; l++) { nxt[l << S] = (uint16_t)o; o += cnt[l << S]; }
        const uint32_t nsym = o;
        for (int s = 0; s < n; s++) {
            const uint32_t l = m.ln[(l0 + (uint32_t)s) << S];
            if (!l) continue;
            const uint32_t at = nxt[l << S]; nxt[l << S] = (uint16_t)(at + 1);
            sorted_set(sw, at, (uint32_t)s);
        }
        uint32_t *T = m.tab;
        const uint32_t rsize = 1u << rbits;
        // slots no code reaches decode as "invalid code": only incomplete sets have them (no code longer than 1 bit)
        if (left > 0) for (uint32_t i = 0; i < rsize; i++) T[(root + i) << S] = mk_entry(1, 0, K_BAD, 0);
        uint32_t code = 0, curlen = 0;
        for (uint32_t i = 0; i < nsym;) {
            const uint32_t s = sorted_get(sw, i), l = m.ln[(l0 + s) << S];
            code <<= (l - curlen); curlen = l;
            if (l <= (uint32_t)rbits) {
                const uint32_t e = entry_for(which, s, l), rev = dflate::brev32(code) >> (32 - l);
                for (uint32_t k = rev; k < rsize; k += (1u << l)) T[(root + k) << S] = e;
                code++; i++;
                continue;
            }
            // a group of long codes sharing their first rbits bits: one second-level table, indexed by the following bits
            const uint32_t prefix = code >> (l - (uint32_t)rbits);
            uint32_t j = i, c2 = code, l2 = l, lmax = l;
            while (j < nsym) {                                          // the group's last member has its longest code
                const uint32_t lj = m.ln[(l0 + sorted_get(sw, j)) << S];
                c2 <<= (lj - l2); l2 = lj;
                if ((c2 >> (lj - (uint32_t)rbits)) != prefix) break;
                lmax = lj; c2++; j++;
            }
            const uint32_t sb = lmax - (uint32_t)rbits, base = *cursor;
            if (base + (1u << sb) > limit) return E_FALLBACK;
            *cursor = base + (1u << sb);
            T[(root + (dflate::brev32(prefix) >> (32 - rbits))) << S] = (sb << 5) | (base << 16);      // indirect
            for (; i < j; i++) {
                const uint32_t si = sorted_get(sw, i), li = m.ln[(l0 + si) << S];
                code <<= (li - curlen); curlen = li;
                const uint32_t rest = li - (uint32_t)rbits;                                           // bits after the root bits
                const uint32_t rev = dflate::brev32(code & ((1u << rest) - 1)) >> (32 - rest);
                const uint32_t e = entry_for(which, si, li);
                for (uint32_t k = rev; k < (1u << sb); k += (1u << rest)) T[(base + k) << S] = e;
                code++;
            }
        }
        return OK;
    }
    // one probe of a root table + the second level when the entry is indirect (headers; the burst loop has its own copy)
    WGBS_HD uint32_t probe(uint32_t root, int rbits) const {
        uint32_t e = m.tab[(root + ((uint32_t)bb & ((1u << rbits) - 1))) << S];
        if (!(e & 31)) e = m.tab[((e >> 16) + (((uint32_t)(bb >> rbits)) & ((1u << ((e >> 5) & 15)) - 1))) << S];
        return e;
    }

    // (scratch of the table construction: the arena's last words -- 144 for the 288 literal/length symbols, 16 for the other sets)
    // literal/length tables from ln[0 .. nlen), distance tables from ln[nlen .. nlen + ndist)
    WGBS_HD int both_tables(int nlen, int ndist) {
        if (defer) { pend_nlen = (uint16_t)nlen; pend_ndist = (uint16_t)ndist; return OK; }
        uint32_t cur = 1u << LB;
        int r = build(0, nlen, 0, LB, &cur, ARENA - 144, 0, ARENA - 144);
        if (r) return r;
        dt_off = cur; cur += 1u << DB;
        if (cur > ARENA - 16) return E_FALLBACK;
        return build((uint32_t)nlen, ndist, dt_off, DB, &cur, ARENA - 16, 1, ARENA - 16);
    }
    WGBS_HD int fixed_tables() {
        for (uint32_t i = 0; i < 144; i++) m.ln[i << S] = 8;
        for (uint32_t i = 144; i < 256; i++) m.ln[i << S] = 9;
        for (uint32_t i = 256; i < 280; i++) m.ln[i << S] = 7;
        for (uint32_t i = 280; i < 288; i++) m.ln[i << S] = 8;
        // 30 five-bit distance codes + the two unused ones (they decode to "invalid" like zlib's fixed table): 32 symbols, complete
        for (uint32_t i = 288; i < 320; i++) m.ln[i << S] = 5;
        return both_tables(288, 32);
    }

    WGBS_HD int dynamic_tables() {
        uint32_t v = take(14);
        const int nlen = (int)(v & 31) + 257, ndist = (int)((v >> 5) & 31) + 1, ncode = (int)((v >> 10) & 15) + 4;
        if (nlen > 286 || ndist > 30) return E_CODES;
        const char *order = "\x10\x11\x12\x00\x08\x07\x09\x06\x0a\x05\x0b\x04\x0c\x03\x0d\x02\x0e\x01\x0f";
        for (uint32_t i = 0; i < 19; i++) m.ln[i << S] = 0;
        for (int i = 0; i < ncode; i++) m.ln[(uint32_t)order[i] << S] = (uint8_t)take(3);
        // the code-length code (<= 7 bits: a 128-entry root table, never a second level) is built behind the literal/length root
        uint32_t cur = ARENA;
        int r = build(0, 19, 1u << LB, 7, &cur, ARENA, 2, ARENA - 16);
        if (r) return r;
        int idx = 0; uint32_t prev = 0;
        // the lengths go to ln[0 ..): the code-length code's own lengths are no longer needed (its table is built)
        while (idx < nlen + ndist) {
            if (bc < 32) refill_careful();
            const uint32_t e = m.tab[((1u << LB) + ((uint32_t)bb & 127u)) << S];
            if (((e >> 9) & 3) != K_LIT) return E_SYMBOL;
            bb >>= (e & 31); bc -= (e & 31);
            const uint32_t s = e >> 16;
            if (s < 16) { m.ln[(uint32_t)idx++ << S] = (uint8_t)s; prev = s; continue; }
            int rep; uint32_t val = 0;
            if (s == 16) { if (idx == 0) return E_CODES; val = prev; rep = 3 + (int)take(2); }
            else if (s == 17) { rep = 3 + (int)take(3); }
            else { rep = 11 + (int)take(7); }
            if (idx + rep > nlen + ndist) return E_CODES;
            while (rep--) m.ln[(uint32_t)idx++ << S] = (uint8_t)val;
            prev = val;
        }
        if (bitpos() > end_bit) return E_INPUT;
        if (m.ln[256u << S] == 0) return E_CODES;
        return both_tables(nlen, ndist);
    }

    // bytes of a stored block copied by this lane itself (only past MAX_STORED_TOKENS stored blocks)
    WGBS_HD void stored_inline(uint32_t from, uint32_t n) { for (uint32_t k = 0; k < n; k++) dst[opos + k] = src[from + k]; }

    // ---- one deflate block header: returns the next state ------------------------------------------------------------------
    WGBS_HD int header() {
        ring_commit_wait();                    // chunks requested by the last top-up may still be in flight: a header reads on without pacing
        if (last) { if (opos != dst_len) rc = E_SHORT; else if (bitpos() > end_bit) rc = E_INPUT; return ST_DONE; }
        if (bitpos() + 3 > end_bit) { rc = E_INPUT; return ST_DONE; }
        const uint32_t hdr = take(3);
        last = hdr & 1;
        const uint32_t type = hdr >> 1;
        if (type == 0) {
            take(bc & 7);                                                  // to the byte boundary
            if (bitpos() + 32 > end_bit) { rc = E_INPUT; return ST_DONE; }
            const uint32_t v = take(32);
            if ((v & 0xffffu) != ((~v >> 16) & 0xffffu)) { rc = E_STORED; return ST_DONE; }
            const uint32_t slen = v & 0xffffu, pos = bitpos(), from = (pos >> 3) - mis;
            if (pos + 8 * slen > end_bit) { rc = E_INPUT; return ST_DONE; }
            if (opos + slen > dst_len) { rc = E_OUTPUT; return ST_DONE; }
            if (slen) {
                if (nstored < MAX_STORED_TOKENS) { tok[ntok].x = opos | (slen << 16); tok[ntok].y = TOK_STORED | from; ntok++; nstored++; }
                else stored_inline(from, slen);
                opos += slen;
                seek(pos + 8 * slen);
            }
            return ST_HDR;
        }
        if (type == 3) { rc = E_BTYPE; return ST_DONE; }
        rc = type == 1 ? fixed_tables() : dynamic_tables();
        if (rc != OK) return ST_DONE;
        len = 0;
        ring_fill_sync();
        return ST_DEC;
    }
};

// ---- phase 2: replay the token list of one block, lane = output BYTE ------------------------------------------------------------------
// lanes: dflate::WarpLanes / dflate::OneLane / the test's lock-step emulation.  dst: the block's output (literals already in place),
// payload: the block's deflate payload (stored blocks).  Every lane returns the same verdict.
// A replay with one MATCH per lane (each lane copying its own match byte by byte under predicates, long ones by the whole warp) spent
// ~65 warp instructions per match (ncu: 292 000 per block, 64 registers + spills) and was removed.  Here the ready
// matches of a round are laid end to end (exclusive sum of their lengths) and the lanes take consecutive BYTES of that space: a
// five-step search over the lanes' offsets (shuffles) finds the byte's token, one load and one store move it.  A round's sources all
// lie below its first unfinished match and its destinations at or above it, so the bytes of a round are independent of each other.
WGBS_HD void prefetch_l2(const void *p) {
#if defined(__CUDA_ARCH__)
    asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p));
#else
    (void)p;
#endif
}
template <class L>
WGBS_HD int resolve_bytes(L lanes, const Token *tok, uint32_t ntok, uint8_t *dst, uint32_t dst_len, const uint8_t *payload) {
    const uint32_t lane = (uint32_t)lanes.id();
    constexpr uint32_t N = (uint32_t)L::N;
    // tokens travel two batches ahead of the replay; the output bytes the NEXT batch will read (its sources) and write over (the
    // literals between its matches came from the decoder a millisecond ago: DRAM by now) are asked for one batch ahead -- a round
    // of the replay is one dependent memory round trip, and without this it is a DRAM round trip (ncu: L2 hit rate 44 %)
    Token n1, n2; n1.x = n1.y = n2.x = n2.y = 0;
    if (lane < ntok) n1 = tok[lane];
    if (N + lane < ntok) n2 = tok[N + lane];
    for (uint32_t t0 = 0; t0 < ntok; t0 += N) {
        const bool valid = t0 + lane < ntok;
        const Token tk = n1;
        n1 = n2;
        if (t0 + 2 * N + lane < ntok) n2 = tok[t0 + 2 * N + lane];
        if (t0 + N + lane < ntok) {
            const uint32_t a1 = n1.x & 0xffffu;
            prefetch_l2(dst + a1);
            if (!(n1.y & TOK_STORED)) prefetch_l2(dst + a1 - (n1.y & ~TOK_STORED));
        }
        const uint32_t at = tk.x & 0xffffu, len = valid ? tk.x >> 16 : 0u;
        const bool stored = (tk.y & TOK_STORED) != 0;
        const uint32_t dist = tk.y & ~TOK_STORED;
        const uint32_t src_end = stored ? 0u : at - dist + (len < dist ? len : dist);
        uint32_t pending = lanes.ballot(valid);
        while (pending) {
            const int first = dflate::lowest_bit(pending);
            const uint32_t hwm = lanes.shfl(at, first);              // everything below is final
            const bool ready = ((pending >> lane) & 1u) && src_end <= hwm;
            const uint32_t R = lanes.ballot(ready);
            const uint32_t l = ready ? len : 0u;
            uint32_t M = 0;
            const uint32_t off = lanes.exscan(l, &M, l);             // where this lane's match starts in the round's byte space
            // UN independent steps of N bytes per iteration: their searches (chains of dependent shuffles), loads and stores overlap
            constexpr uint32_t UN = 4;
            for (uint32_t j0 = 0; j0 < M; j0 += UN * N) {
                uint8_t b[UN]; uint32_t p[UN]; bool on[UN]; uint32_t lo[UN];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                for (uint32_t u = 0; u < UN; u++) lo[u] = 0;         // last lane whose match starts at or before byte j
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                for (uint32_t step = N / 2; step; step >>= 1) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                    for (uint32_t u = 0; u < UN; u++) { const uint32_t v = lanes.shfl(off, (int)(lo[u] + step)); if (v <= j0 + u * N + lane) lo[u] += step; }
                }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                for (uint32_t u = 0; u < UN; u++) {
                    const uint32_t j = j0 + u * N + lane;
                    const uint32_t x = lanes.shfl(tk.x, (int)lo[u]), y = lanes.shfl(tk.y, (int)lo[u]), o = lanes.shfl(off, (int)lo[u]);
                    const uint32_t r = j - o, a = x & 0xffffu, ln = x >> 16, d = y & ~TOK_STORED;
                    on[u] = j < M; p[u] = a + r; b[u] = 0;
                    if (on[u]) {
                        const bool st = (y & TOK_STORED) != 0;
                        uint32_t sr = r;                              // a run (dist < len) repeats its first dist bytes
                        if (!st && d < ln) { sr = 0; if (WGBS_UNLIKELY(d != 1)) sr = r % d; }
                        const uint8_t *from = st ? payload + d + r : dst + a - d + sr;
                        b[u] = *from;
                    }
                }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                for (uint32_t u = 0; u < UN; u++) if (on[u]) dst[p[u]] = b[u];
            }
            lanes.sync();                                            // this round's bytes are visible to the next round's loads
            pending &= ~R;
        }
    }
    (void)dst_len;
    return OK;
}

// ---- CRC-32 of a block's output, four bytes per step (slicing-by-4) -----------------------------------------------------------
// T = 4 tables of 256 entries: T[0] = the byte table, T[k][i] = T[0][T[k-1][i] & 0xff] ^ (T[k-1][i] >> 8).  Every lane folds its own
// contiguous slice from a zero register (bytes up to the first 4-byte boundary, aligned words, trailing bytes); the slices are
// chained like dflate::crc32_block: register(A||B) = register(A) * x^(8|B|) mod P  xor  register_0(B).
WGBS_HD uint32_t crc_slice_entry(uint32_t k, uint32_t i) {       // T[k][i]
    uint32_t c = dflate::crc_table_entry(i);
    for (uint32_t j = 0; j < k; j++) c = dflate::crc_table_entry(c & 0xff) ^ (c >> 8);
    return c;
}
template <class L>
WGBS_HD uint32_t crc32_block4(L lanes, const uint8_t *d, uint32_t n, const uint32_t *T) {
    const uint32_t per = (n + L::N - 1) / L::N;
    const uint32_t l = (uint32_t)lanes.id();
    const uint32_t a = l * per < n ? l * per : n, b = a + per < n ? a + per : n;
    uint32_t c = 0, i = a;
    auto word = [&](uint32_t w) { c ^= w; c = T[768 + (c & 0xff)] ^ T[512 + ((c >> 8) & 0xff)] ^ T[256 + ((c >> 16) & 0xff)] ^ T[c >> 24]; };
    for (; i < b && ((uintptr_t)(d + i) & 15); i++) c = T[(c ^ d[i]) & 0xff] ^ (c >> 8);
    // 16 bytes per step, loaded one step ahead of the table walk that consumes them
    if (i + 16 <= b) {
        const uint32_t *p = (const uint32_t *)(d + i);
        uint32_t w0 = p[0], w1 = p[1], w2 = p[2], w3 = p[3];
        i += 16;
        for (; i + 16 <= b; i += 16) {
            const uint32_t *q = (const uint32_t *)(d + i);
            const uint32_t v0 = q[0], v1 = q[1], v2 = q[2], v3 = q[3];
            word(w0); word(w1); word(w2); word(w3);
            w0 = v0; w1 = v1; w2 = v2; w3 = v3;
        }
        word(w0); word(w1); word(w2); word(w3);
    }
    for (; i + 4 <= b; i += 4) word(*(const uint32_t *)(d + i));
    for (; i < b; i++) c = T[(c ^ d[i]) & 0xff] ^ (c >> 8);
    const uint32_t nfull = per ? n / per : 0, rem = per ? n - nfull * per : 0;
    uint32_t op = 0;
    if (l == 0) op = dflate::crc_x8n(per);
    if (L::N > 1 && l == 1) op = dflate::crc_x8n(rem);
    const uint32_t x_full = lanes.shfl(op, 0), x_rem = L::N > 1 ? lanes.shfl(op, 1) : dflate::crc_x8n(rem);
    uint32_t t = 0xffffffffu;
    for (int k = 0; k < L::N; k++) {
        const uint32_t ck = lanes.shfl(c, k);
        const uint32_t len_k = (uint32_t)k < nfull ? per : ((uint32_t)k == nfull ? rem : 0);
        if (len_k == 0) continue;
        t = dflate::crc_mulmod(t, len_k == per ? x_full : x_rem) ^ ck;
    }
    return t ^ 0xffffffffu;
}

}  // namespace dflate2
