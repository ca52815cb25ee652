// inflate3_core.cuh -- team decoder: the lanes of a team walk ONE deflate block of a BGZF block together.
//
// The Huffman stream of a block is serial only in principle: a decoder started at a wrong bit offset falls into step with the
// true symbol boundaries after a few symbols (zlib level-6 BAM blocks: median 7 symbols / 92 bits, never more than 800 bits in
// 780 trials over 39 blocks, tools/deflate_sync_stats.py).  So the bits of a deflate block -- after its header, which one lane
// parses with the code of inflate2_core.cuh -- are cut into one span per lane, and
//
//   pass A   every lane but the first starts SYNC_BITS before its span and walks to the span's start WITHOUT writing anything.
//            The first symbol boundary it finds there is the lane's ANCHOR: from there on its walk is taken to be the true one.
//            (Lane 0 starts at the true start of the block.)
//   pass B   every lane walks from its anchor until it reaches or passes its successor's anchor, counting output bytes and
//            tokens.  Reaching the anchor exactly PROVES the successor's walk from there on (given this lane's own walk is true).
//   chain    lane 0 is true by construction; lane k + 1 is true if lane k is and reached its anchor.  A lane whose anchor was
//            missed is dropped and its predecessor walks on to the anchor after that (rare: costs one span of serial work).
//   pass C   exclusive sums of the counts give every surviving lane its output position and token slot; it repeats its walk
//            from its anchor, now writing literals to their final place and matches to the token list -- the same
//            (output position | length << 16, distance) tokens bgzf_resolve_k replays.
//
// A lane walks ~2 spans + SYNC_BITS instead of the whole block: with 32 lanes ~7 500 bits instead of ~100 000.
// Host + device code (lane policies of inflate_core.cuh; tests/bamdev_core_check.cpp runs it under the lock-step emulation
// against zlib).  Stands in for the zlib inflate inside `samtools view` (reference src/python/bam2pat.py:165).
#pragma once
#include "inflate2_core.cuh"

namespace dflate3 {

using dflate2::Token; using dflate2::LB; using dflate2::DB; using dflate2::K_LIT; using dflate2::K_BASE; using dflate2::K_EOB;
using dflate::OK; using dflate::E_INPUT; using dflate::E_SYMBOL; using dflate::E_DIST; using dflate::E_OUTPUT;

constexpr uint32_t NOPOS = 0xffffffffu;
#ifndef WGBS_SYNC_BITS
#define WGBS_SYNC_BITS 1024
#endif
constexpr uint32_t SYNC_BITS = WGBS_SYNC_BITS;      // a walk started anywhere is taken to be in step this many bits later (pass B verifies it)
constexpr uint32_t MIN_SPAN = 3072;       // shortest span worth a lane of its own (> SYNC_BITS: a lane's run-up stays inside the block)
enum : uint32_t { F_EOB = 1, F_BAD = 2, F_END = 4 };

// working memory of one team (shared memory on the device): the header decoder of inflate2_core.cuh with its plain-layout
// arrays, and the hand-over words between the leader and the team
struct TeamMem {
    dflate2::HostLane arrays;
    dflate2::Decoder<0> D;
    int st;
    uint32_t h;
};

// one lane's walk over the Huffman symbols of the current deflate block
struct Walker {
    const uint32_t *gw; uint32_t end_bit;             // the block's payload as 32-bit words (inflate2_core.cuh: Decoder::gw)
    const uint32_t *tab; uint32_t tab_sa, dt_off;      // the team's tables
    uint64_t bb; uint32_t bc, wp, nw;                  // bit buffer: bc >= 32 valid bits, word wp (= nw) is the next to join
    uint32_t out, ntk;                                 // bytes / tokens: counts (passes A, B), absolute positions (pass C)
    uint32_t flags;
    uint8_t *dst; uint32_t dst_len; Token *tok;
    int rc;
#if defined(WGBS_COUNT_ITERS)
    uint64_t iters;                                    // host statistics: table probes
#endif

    // (no bounds check: a walk stops at the first symbol boundary past end_bit, so it reads at most three words past the payload's last
    // one -- the gzip trailer and the padding every compressed buffer carries, bgzf.cuh)
    WGBS_HD uint32_t gword(uint32_t i) const {
#if defined(__CUDA_ARCH__)
        return __ldg(gw + i);
#else
        return gw[i];
#endif
    }
    WGBS_HD uint32_t tload(uint32_t idx) const {
#if defined(__CUDA_ARCH__)
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(v) : "r"(tab_sa + (idx << 2)));
        return v;
#else
        return tab[idx];
#endif
    }
    WGBS_HD uint32_t pos() const { return 32 * wp - bc; }
    // (word wp + 1 is loaded straight into nw at every refill.  A per-lane ring in shared memory filled by cp.async -- no register
    // scoreboard between the request and the next refill -- was measured too: long_scoreboard 4.85 -> 2.43 per issue, but 17 % more
    // instructions and the same 0.98 ms, profiles/README.md round 2)
    WGBS_HD void seek(uint32_t p) {
        const uint32_t w = p >> 5, s = p & 31;
        bb = ((uint64_t)gword(w) | ((uint64_t)gword(w + 1) << 32)) >> s;
        bc = 64 - s; wp = w + 2; nw = gword(wp);
        flags = 0;
    }
    // walk on until a symbol boundary at or past `stop` (NOPOS: until the end-of-block code).  Sets F_EOB (pos() = the bit after
    // the code), F_BAD (invalid code; rc says which) or F_END (the payload ended first).
    // One table probe per iteration -- literal/length, or the distance of the match whose length the previous probe gave -- and
    // no branch that depends on what the probe found (the lanes of a team are at different kinds of symbols all the time): only
    // the second-level probe of a long code and the exit are branches.
    template <bool WRITE>
    WGBS_HD void run_until(uint32_t stop) {
        if (flags) return;
        const uint32_t lim = stop < end_bit ? stop : end_bit;
        int32_t rem = (int32_t)(lim - pos());                     // bits to go (positions are < 2^20)
        uint32_t len = 0, wd = 0, fl = 0;
        for (;;) {
            if (!wd && rem <= 0) break;
#if defined(WGBS_COUNT_ITERS)
            iters++;
#endif
            uint32_t e = tload(wd ? dt_off + ((uint32_t)bb & ((1u << DB) - 1)) : (uint32_t)bb & ((1u << LB) - 1));
            if (WGBS_UNLIKELY(!(e & 31)))
                e = tload((e >> 16) + (((uint32_t)(bb >> (wd ? DB : LB))) & ((1u << ((e >> 5) & 15)) - 1)));
            const uint32_t tot = e & 31, nb = tot - ((e >> 5) & 15), kind = (e >> 9) & 3;
            const uint32_t val = (e >> 16) + (((uint32_t)bb & ((1u << tot) - 1)) >> nb);      // tot <= 28 of the >= 32 valid bits
            bb >>= tot; bc -= tot; rem -= (int32_t)tot;
            if (bc < 32) { bb |= (uint64_t)nw << bc; bc += 32; wp++; nw = gword(wp); }
            const bool base = kind == K_BASE, is_lit = !wd && kind == K_LIT, is_dst = wd && base;
            const uint32_t adv = is_lit ? 1u : (is_dst ? len : 0u);
            bool odd = !(is_lit || base) || tot == 0;
            if (WRITE) odd = odd || out + adv > dst_len || (is_dst && val > out);
            if (WGBS_UNLIKELY(odd)) {
                if (!wd && kind == K_EOB && tot) fl = F_EOB;
                else { fl = F_BAD; rc = (is_lit || base) ? ((is_dst && val > out) ? E_DIST : E_OUTPUT) : E_SYMBOL; }
                break;
            }
            if (WRITE) {
                if (is_lit) dst[out] = (uint8_t)val;
                if (is_dst) { Token t; t.x = out | (len << 16); t.y = val; tok[ntk] = t; }
            }
            out += adv; ntk += is_dst ? 1u : 0u;
            len = wd ? len : val;                                  // (a literal's value is never read as a length)
            wd = (!wd && base) ? 1u : 0u;
        }
        if (!fl && stop > end_bit) fl = F_END;
        flags |= fl;
    }
};

// ---- the Huffman tables of a deflate block, built by the team ------------------------------------------------------------------------
// Same tables, same accept / reject / fallback verdicts as Decoder::build (inflate2_core.cuh), which one lane runs symbol by symbol
// (~40 000 instructions for the two code sets of a dynamic block: a third of the team decoder's time when only the leader did it).
// Here lane l counts the codes of length l; the canonical code of a symbol = first code of its length + its rank among the symbols
// of that length (ballots over a chunk of N symbols + a running count per length); every lane then writes the root entries of its
// own symbols.  Codes longer than the root index: the first member of every group (codes sharing the root bits) allocates the
// group's second-level table (exclusive sum over the lanes), whose size comes from the longest member (one round of plain stores
// per code length), then every member writes its own entries.
// lens ln[l0 .. l0 + n) -> root table of 1 << rbits entries at arena offset `root`, second-level tables from *cursor on (below
// `limit`); which: 0 literal/length, 1 distance; sw: arena word where n 16-bit scratch entries (the codes) start.
WGBS_HD uint32_t popc32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__popc(v);
#else
    return (uint32_t)__builtin_popcount(v);
#endif
}
template <class L>
WGBS_HD int team_build(L lanes, dflate2::Mem<0> m, uint32_t l0, uint32_t n, uint32_t root, uint32_t rbits, uint32_t *cursor, uint32_t limit, int which, uint32_t sw) {
    const uint32_t lane = (uint32_t)lanes.id();
    constexpr uint32_t N = (uint32_t)L::N;
    uint16_t *bk = m.bk;                                        // [0, 16): symbols of each length seen so far; [16, 32): first code of each length
    uint16_t *code16 = reinterpret_cast<uint16_t *>(m.tab + sw);
    uint32_t *T = m.tab;
    for (uint32_t l = lane; l < 16; l += N) {
        uint32_t c = 0;
        for (uint32_t s = 0; s < n; s++) c += m.ln[l0 + s] == l ? 1u : 0u;
        bk[l] = (uint16_t)c;
    }
    lanes.sync();
    int left = 1; uint32_t maxlen = 0, code = 0;
    for (uint32_t l = 1; l <= 15; l++) {
        const uint32_t c = bk[l];
        left <<= 1; left -= (int)c;
        if (left < 0) return dflate::E_CODES;
        if (c) maxlen = l;
        code = (code + (l > 1 ? (uint32_t)bk[l - 1] : 0u)) << 1;
        if (lane == 0) bk[16 + l] = (uint16_t)code;             // first code of length l
    }
    if (left > 0 && (maxlen > 1 || which == 2)) return dflate::E_CODES;
    lanes.sync();                                               // every lane has read the counts
    if (lane == 0) for (uint32_t l = 1; l <= 15; l++) bk[l] = 0;
    const uint32_t rsize = 1u << rbits;
    if (left > 0) for (uint32_t i = lane; i < rsize; i += N) T[root + i] = dflate2::mk_entry(1, 0, dflate2::K_BAD, 0);
    lanes.sync();
    // codes of all symbols; root entries of the short ones
    for (uint32_t c0 = 0; c0 < n; c0 += N) {
        const uint32_t s = c0 + lane;
        const uint32_t l = s < n ? m.ln[l0 + s] : 0u;
        const uint32_t before = l ? bk[l] : 0u;
        lanes.sync();                                           // (reads of the running counts before this chunk's updates)
        uint32_t rank = 0;
        uint32_t todo = lanes.ballot(l != 0);
        while (todo) {
            const int f = dflate::lowest_bit(todo);
            const uint32_t lf = lanes.shfl(l, f);
            const uint32_t mm = lanes.ballot(l == lf);
            if (l == lf) rank = popc32(mm & ((1u << lane) - 1u));
            if ((int)lane == f) bk[lf] = (uint16_t)(before + popc32(mm));
            todo &= ~mm;
        }
        if (l) {
            const uint32_t cd = (uint32_t)bk[16 + l] + before + rank;
            code16[s] = (uint16_t)cd;
            if (l <= rbits) {
                const uint32_t e = which == 0 ? dflate2::litlen_entry(s, l) : dflate2::dist_entry(s, l), rev = dflate::brev32(cd) >> (32 - l);
                for (uint32_t k = rev; k < rsize; k += (1u << l)) T[root + k] = e;
            }
        }
        lanes.sync();
    }
    uint32_t cur = *cursor;
    if (maxlen > rbits) {
        // longest member of every group, left in the group's root slot
        for (uint32_t ll = rbits + 1; ll <= maxlen; ll++) {
            for (uint32_t c0 = 0; c0 < n; c0 += N) {
                const uint32_t s = c0 + lane;
                if (s < n && m.ln[l0 + s] == ll) T[root + (dflate::brev32((uint32_t)code16[s] >> (ll - rbits)) >> (32 - rbits))] = ll;
            }
            lanes.sync();
        }
        // the first member of a group (the code whose bits behind the root bits are all zero) allocates its table
        for (uint32_t c0 = 0; c0 < n; c0 += N) {
            const uint32_t s = c0 + lane;
            const uint32_t l = s < n ? m.ln[l0 + s] : 0u;
            bool owner = false; uint32_t r = 0, sb = 0;
            if (l > rbits) {
                const uint32_t cd = code16[s], rest = l - rbits;
                owner = (cd & ((1u << rest) - 1u)) == 0;
                r = dflate::brev32(cd >> rest) >> (32 - rbits);
                if (owner) sb = T[root + r] - rbits;
            }
            const uint32_t sz = owner ? (1u << sb) : 0u;
            uint32_t tot = 0;
            const uint32_t off = lanes.exscan(sz, &tot, sz);
            if (owner) T[root + r] = (sb << 5) | ((cur + off) << 16);      // indirect (offsets beyond the limit are never followed: verdict below)
            cur += tot;
        }
        if (cur > limit) return dflate2::E_FALLBACK;
        lanes.sync();
        for (uint32_t c0 = 0; c0 < n; c0 += N) {
            const uint32_t s = c0 + lane;
            const uint32_t l = s < n ? m.ln[l0 + s] : 0u;
            if (l > rbits) {
                const uint32_t cd = code16[s], rest = l - rbits;
                const uint32_t ent = T[root + (dflate::brev32(cd >> rest) >> (32 - rbits))], base = ent >> 16, sb = (ent >> 5) & 15;
                const uint32_t e = which == 0 ? dflate2::litlen_entry(s, l) : dflate2::dist_entry(s, l);
                const uint32_t rev = dflate::brev32(cd & ((1u << rest) - 1u)) >> (32 - rest);
                for (uint32_t k = rev; k < (1u << sb); k += (1u << rest)) T[base + k] = e;
            }
        }
    }
    lanes.sync();
    *cursor = cur;
    return OK;
}
// both tables of the block whose lengths the leader's header() left in ln[]: literal/length from ln[0 .. nlen), distance from
// ln[nlen .. nlen + ndist) -- Decoder::both_tables; *dt_off = arena offset of the distance root
template <class L>
WGBS_HD int team_tables(L lanes, dflate2::Mem<0> m, uint32_t nlen, uint32_t ndist, uint32_t *dt_off) {
    uint32_t cur = 1u << LB;
    int r = team_build(lanes, m, 0, nlen, 0, (uint32_t)LB, &cur, dflate2::ARENA - 144, 0, dflate2::ARENA - 144);
    if (r) return r;
    *dt_off = cur; cur += 1u << DB;
    if (cur > dflate2::ARENA - 16) return dflate2::E_FALLBACK;
    return team_build(lanes, m, nlen, ndist, *dt_off, (uint32_t)DB, &cur, dflate2::ARENA - 16, 1, dflate2::ARENA - 16);
}

// Statistics of the host tests (nullptr on the device): how the chain went
struct TeamStats { uint64_t blocks, lanes_started, lanes_dropped; };

// All lanes of the team call this together; every lane returns the same verdict.  *ntok_out: tokens written to `tok`.
template <class L>
WGBS_HD int team_inflate(L lanes, TeamMem *T, const uint8_t *payload, uint32_t clen, uint8_t *dst, uint32_t usize, Token *tok, uint32_t *ntok_out,
                         TeamStats *stats = nullptr) {
    const uint32_t lane = (uint32_t)lanes.id();
    constexpr uint32_t S = (uint32_t)L::N;
    if (lane == 0) { T->D.init(T->arrays.mem(), payload, clen, dst, usize, tok); T->D.defer = 1; T->st = dflate2::ST_HDR; }
    lanes.sync();
    for (;;) {
        if (lane == 0) {
            int st = T->st;
            while (st == dflate2::ST_HDR) st = T->D.header();          // stored blocks become tokens right here
            T->st = st; T->h = T->D.bitpos();
        }
        lanes.sync();
        if (T->st == dflate2::ST_DONE) break;
        // ---- a Huffman block starts at bit h: its tables first -------------------------------------------------------------
        {
            uint32_t dt = 0;
            const int r = team_tables(lanes, T->arrays.mem(), T->D.pend_nlen, T->D.pend_ndist, &dt);
            if (r != OK) {                                              // every lane has the same verdict
                lanes.sync();
                if (lane == 0) { T->D.rc = r; T->st = dflate2::ST_DONE; }
                lanes.sync();
                break;
            }
            if (lane == 0) T->D.dt_off = dt;
            lanes.sync();
        }
        Walker w;
        w.gw = T->D.gw; w.end_bit = T->D.end_bit;
        w.tab = T->D.m.tab; w.dt_off = T->D.dt_off; w.tab_sa = 0;
#if defined(__CUDA_ARCH__)
        w.tab_sa = (uint32_t)__cvta_generic_to_shared(T->D.m.tab);
#endif
        w.dst = dst; w.dst_len = usize; w.tok = tok; w.rc = OK; w.out = 0; w.ntk = 0; w.flags = F_END;
        const uint32_t h = T->h, opos0 = T->D.opos, ntok0 = T->D.ntok;
        const uint32_t R = w.end_bit > h ? w.end_bit - h : 0;
        uint32_t span = (R + S - 1) / S;
        if (span < MIN_SPAN) span = MIN_SPAN;
        uint32_t nl = (R + span - 1) / span;
        if (nl == 0) nl = 1;
        // lane k is to walk [h + k * span, h + (k + 1) * span) for real; it starts SYNC_BITS earlier to fall into step (lane 0: at h, in step)
        const uint32_t t0 = h + lane * span;
        const bool active = lane < nl;
        uint32_t anchor = NOPOS;
        // pass A: to the anchor
        if (active) {
            if (lane == 0) { w.seek(h); anchor = h; }
            else {
                w.seek(t0 - SYNC_BITS);
                w.run_until<false>(t0);
                // a guessed walk that runs into an invalid code or a bogus end-of-block code goes on from where it stands: a new guess
                for (int k = 0; k < 8 && w.flags && !(w.flags & F_END) && w.pos() < t0; k++) { w.flags = 0; w.rc = OK; w.run_until<false>(t0); }
                if (w.flags && !(w.flags & F_END) && w.pos() >= t0) { w.flags = 0; w.rc = OK; }     // (the failing guess ended past the span's start: guess on from there)
                if (!w.flags) { anchor = w.pos(); w.out = 0; w.ntk = 0; }
            }
        }
        // pass B: from the anchor to the successor's anchor (a successor without one: to the end of the own span; the chain sorts it out)
        {
            const uint32_t a_next = lanes.shfl(anchor, (int)((lane + 1) % S));
            if (active) w.run_until<false>(lane + 1 < nl ? (a_next != NOPOS ? a_next : t0 + span) : NOPOS);
        }
        // chain: which lanes' walks are true, and where each of them stops
        bool alive = lane == 0;
        uint32_t cur = 0;
        for (;;) {
            uint32_t fc = lanes.shfl(w.flags, (int)cur);
            if (fc) break;
            uint32_t t = cur + 1; bool found = false;
            while (t < nl) {
                const uint32_t a_t = lanes.shfl(anchor, (int)t);
                if (lane == cur && a_t != NOPOS) w.run_until<false>(a_t);
                const uint32_t p = lanes.shfl(w.pos(), (int)cur);
                fc = lanes.shfl(w.flags, (int)cur);
                if (fc) break;
                if (a_t != NOPOS && p == a_t) { found = true; break; }
                if (stats && lane == 0) stats->lanes_dropped++;
                t++;
            }
            if (fc) break;
            if (!found) { if (lane == cur) w.run_until<false>(NOPOS); break; }
            if (lane == t) alive = true;
            cur = t;
        }
        if (stats && lane == 0) { stats->blocks++; stats->lanes_started += nl; }
        const uint32_t term = cur;
        const uint32_t tf = lanes.shfl(w.flags, (int)term);
        const uint32_t eob_end = lanes.shfl(w.pos(), (int)term);
        uint32_t tot_out = 0, tot_tok = 0;
        const uint32_t c_out = alive ? w.out : 0u, c_tok = alive ? w.ntk : 0u;
        const uint32_t obase = lanes.exscan(c_out, &tot_out, c_out), tbase = lanes.exscan(c_tok, &tot_tok, c_tok);
        int rc = OK;
        if (!(tf & F_EOB)) rc = (tf & F_END) ? E_INPUT : E_SYMBOL;
        else if (eob_end > w.end_bit) rc = E_INPUT;
        else if (tot_out > usize - opos0) rc = E_OUTPUT;
        // pass C
        if (rc == OK) {
            const uint32_t stop = lane == term ? NOPOS : w.pos();
            bool bad = false;
            if (alive) {
                w.seek(anchor); w.rc = OK;
                w.out = opos0 + obase; w.ntk = ntok0 + tbase;
                w.run_until<true>(stop);
                bad = lane == term ? w.flags != F_EOB : (w.flags != 0 || w.pos() != stop);
            }
            const uint32_t bm = lanes.ballot(bad);
            if (bm) { const int f = dflate::lowest_bit(bm); rc = (int)lanes.shfl((uint32_t)(w.rc ? w.rc : E_SYMBOL), f); }
        }
        if (lane == 0) {
            if (rc != OK) { T->D.rc = rc; T->st = dflate2::ST_DONE; }
            else { T->D.opos = opos0 + tot_out; T->D.ntok = ntok0 + tot_tok; T->D.len = 0; T->D.seek(eob_end); T->st = dflate2::ST_HDR; }
        }
        lanes.sync();               // (the verdict is read at the top of the loop, behind the leader's next header: no lane may look at st here)
    }
    const int rc = T->D.rc;
    *ntok_out = T->D.ntok;
    lanes.sync();
    return rc;
}

}  // namespace dflate3
