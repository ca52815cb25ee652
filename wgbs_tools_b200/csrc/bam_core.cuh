// bam_core.cuh -- per-record logic of the device-side BAM front end: finding record boundaries in the
// inflated BAM stream, the `samtools view` filters, and one record -> one SAM text line.  Host + device: bamdev.cu runs
// it in kernels, tests/bamdev_core_check.cpp compiles it with g++ and pins it against the host decoder (bam.cu) and the
// original SAM without a GPU.  Format: SAM/BAM specification v1, sections 4.2 (records) and 1.5 / 4.2.4 (optional fields).
// Stands in for `samtools view BAM region -q -F -f ...` of reference src/python/bam2pat.py:126-165.
#pragma once
#include <stdint.h>

#include "inflate_core.cuh"   // WGBS_HD, lane policies

namespace bamcore {

WGBS_HD uint32_t ld32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
WGBS_HD int32_t ldi32(const uint8_t *p) { return (int32_t)ld32(p); }
WGBS_HD uint32_t ld16(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }

// reference dictionary of the BAM header, flattened: names[name_off[i] .. name_off[i+1]) without NUL
struct Refs {
    int32_t n;
    const uint32_t *name_off;
    const char *names;
    const int32_t *lens;
};

// fixed-length part of a record (offsets from the block_size field, SAM spec 4.2)
struct Rec {
    const uint8_t *r;      // -> block_size
    uint32_t bs;
    int32_t refid, pos, l_seq, nref, npos, tlen;
    uint32_t l_name, mapq, n_cig, flag;
    WGBS_HD void load(const uint8_t *p) {
        r = p; bs = ld32(p); refid = ldi32(p + 4); pos = ldi32(p + 8); l_name = p[12]; mapq = p[13]; n_cig = ld16(p + 16); flag = ld16(p + 18);
        l_seq = ldi32(p + 20); nref = ldi32(p + 24); npos = ldi32(p + 28); tlen = ldi32(p + 32);
    }
    WGBS_HD const uint8_t *name() const { return r + 36; }
    WGBS_HD const uint8_t *cigar() const { return r + 36 + l_name; }
    WGBS_HD const uint8_t *seq() const { return cigar() + 4 * (uint64_t)n_cig; }
    WGBS_HD const uint8_t *qual() const { return seq() + (uint64_t)((l_seq + 1) / 2); }
    WGBS_HD const uint8_t *tags() const { return qual() + (uint64_t)(l_seq > 0 ? l_seq : 0); }
    WGBS_HD const uint8_t *end() const { return r + 4 + (uint64_t)bs; }
    // the variable-length parts fit inside block_size
    WGBS_HD bool consistent() const {
        if (bs < 32 || l_seq < 0) return false;
        const uint64_t need = 32ull + l_name + 4ull * n_cig + (uint64_t)((l_seq + 1) / 2) + (uint64_t)l_seq;
        return need <= bs;
    }
};

// ---- record boundaries -------------------------------------------------------------------------------------------------
// Does a record plausibly start at offset o of data[0..n)?  Only a GUESS used to seed the segment walk: wrong guesses
// are found and repaired by the chain verification (bamdev.cu), so this test affects speed, never the result.
WGBS_HD bool plausible_at(const uint8_t *data, uint64_t n, uint64_t o, int32_t n_ref) {
    if (o + 36 > n) return false;
    const uint8_t *p = data + o;
    const uint32_t bs = ld32(p);
    if (bs < 32 || o + 4 + (uint64_t)bs > n) return false;
    const int32_t refid = ldi32(p + 4), pos = ldi32(p + 8), nref = ldi32(p + 24), npos = ldi32(p + 28), l_seq = ldi32(p + 20);
    if (refid < -1 || refid >= n_ref || nref < -1 || nref >= n_ref || pos < -1 || npos < -1 || l_seq < 0) return false;
    const uint32_t l_name = p[12], n_cig = ld16(p + 16);
    if (l_name < 2) return false;                   // at least one character + NUL (SAM spec 1.4: [!-?A-~]{1,254})
    if (32ull + l_name + 4ull * n_cig + (uint64_t)((l_seq + 1) / 2) + (uint64_t)l_seq > bs) return false;
    if (p[36 + l_name - 1] != 0) return false;
    for (uint32_t k = 0; k + 1 < l_name; k++) { const uint8_t c = p[36 + k]; if (c < 33 || c > 126) return false; }
    return true;
}
// plausible chain: `depth` consecutive plausible records (or the exact end of the stream)
WGBS_HD bool plausible_chain(const uint8_t *data, uint64_t n, uint64_t o, int32_t n_ref, int depth) {
    for (int d = 0; d < depth; d++) {
        if (o == n) return d > 0;
        if (!plausible_at(data, n, o, n_ref)) return false;
        o += 4 + (uint64_t)ld32(data + o);
    }
    return true;
}

// First offset >= base at which a plausible chain starts (n: none).  Lanes test neighbouring candidates in parallel.
template <class L>
WGBS_HD uint64_t guess_entry(L lanes, const uint8_t *data, uint64_t n, uint64_t base, int32_t n_ref, int depth) {
    for (uint64_t o = base; o < n; o += L::N) {
        const uint64_t c = o + (uint64_t)lanes.id();
        const bool ok = c < n && plausible_chain(data, n, c, n_ref, depth);
        const uint32_t m = lanes.ballot(ok);
        if (m) { int i = 0; while (!((m >> i) & 1)) i++; return o + (uint64_t)i; }
    }
    return n;
}

// Walk the length-prefixed chain from `from` while the record STARTS before `limit`.  Returns the offset of the first record
// start >= limit (n at the end of the stream); *count = records visited; out (optional) receives their offsets.
// *bad is set to the offset of a record that cannot be one (block_size < 32 or running past the end of the stream).
WGBS_HD uint64_t walk_chain(const uint8_t *data, uint64_t n, uint64_t from, uint64_t limit, uint32_t *count, uint64_t *out, uint64_t *bad) {
    uint64_t o = from; uint32_t c = 0;
    while (o < limit && o < n) {
        if (o + 4 > n) { *bad = o; break; }
        const uint32_t bs = ld32(data + o);
        if (bs < 32 || o + 4 + (uint64_t)bs > n) { *bad = o; break; }
        if (out) out[c] = o;
        c++; o += 4 + (uint64_t)bs;
    }
    *count = c;
    return o;
}

// Repair step of the boundary search.  Invariant wanted: entry[s+1] == exit[s] for every s (entry[0] is exact).  Segment s
// may overrule the entry of its successor only if its own entry agrees with ITS predecessor's exit ("supported"): a wrong
// guess is inconsistent with the segment before it, so whatever it walked into cannot spread; isolated wrong guesses are
// repaired in one round, a run of k wrong guesses in k rounds.  Returns the new entry of segment s+1.
WGBS_HD uint64_t repaired_entry(const uint64_t *entry, const uint64_t *exit_, const uint64_t *bad, uint64_t s, bool *changed) {
    const uint64_t cur = entry[s + 1];
    if (bad[s] == ~0ull && cur != exit_[s] && (s == 0 || entry[s] == exit_[s - 1])) { *changed = true; return exit_[s]; }
    return cur;
}

// ---- filters (`samtools view` options of bam2pat.py:126-159; same semantics as wgbs_bam_view_ex in bam.cu) --------------
struct ViewParams {
    int32_t refid;                 // -1: any
    int32_t min_mapq, exclude_flags, include_flags;
    int64_t beg, end;              // 1-based closed interval, end <= 0: none
    int32_t n_flag_eq; int32_t flag_eq[4];
    const int64_t *iv_beg, *iv_end; uint64_t n_iv; int32_t iv_exclude;   // sorted disjoint 0-based half-open intervals
    const char *rg; uint32_t rg_len; int32_t have_rg;
    int64_t key_beg, key_end;      // template window on max(POS, PNEXT) (0-based half-open), key_end <= 0: none (wgbs_b200.h)
};

// the position a record's TEMPLATE is filed under: both mates of a pair on one reference share it (0-based)
WGBS_HD int64_t template_key(int32_t flag, int32_t refid, int32_t pos, int32_t nref, int32_t npos) {
    if ((flag & 1) && !(flag & 8) && nref == refid && npos > pos) return npos;
    return pos;
}

// value of the first Z-typed tag `ab`, or nullptr (malformed tag area: nullptr)
WGBS_HD const uint8_t *find_z_tag(const uint8_t *t, const uint8_t *end, char a, char b, uint32_t *len) {
    while (t + 3 <= end) {
        const char ty = (char)t[2]; const bool hit = (char)t[0] == a && (char)t[1] == b; t += 3;
        if (ty == 'A' || ty == 'c' || ty == 'C') t += 1;
        else if (ty == 's' || ty == 'S') t += 2;
        else if (ty == 'i' || ty == 'I' || ty == 'f') t += 4;
        else if (ty == 'Z' || ty == 'H') {
            const uint8_t *z = t; while (t < end && *t) t++;
            if (hit && ty == 'Z') { *len = (uint32_t)(t - z); return z; }
            t++;
        } else if (ty == 'B') {
            if (t + 5 > end) return nullptr;
            const char sub = (char)t[0]; const uint32_t cnt = ld32(t + 1); t += 5;
            const uint32_t w = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : (sub == 'i' || sub == 'I' || sub == 'f') ? 4 : 0;
            if (!w) return nullptr;
            if ((uint64_t)cnt * w > (uint64_t)(end - t)) return nullptr;
            t += (uint64_t)cnt * w;
        } else return nullptr;
    }
    return nullptr;
}

// the tags `patter --nanopore` reads (ont.cpp:418-438 get_np_tags, on SAM text: the LAST "MM:Z:" / "Mm:Z:" field and the LAST
// "ML:B:C" / "Ml:B:C" field): *mm = the Z string (mm_len bytes, no NUL), *ml = the uint8 values (ml_cnt of them); nullptr when absent
// or when the tag area is malformed
WGBS_HD void find_np_tags(const uint8_t *t, const uint8_t *end, const uint8_t **mm, uint32_t *mm_len, const uint8_t **ml, uint32_t *ml_cnt) {
    *mm = nullptr; *ml = nullptr; *mm_len = 0; *ml_cnt = 0;
    while (t + 3 <= end) {
        const char a = (char)t[0], b = (char)t[1], ty = (char)t[2]; t += 3;
        if (ty == 'A' || ty == 'c' || ty == 'C') t += 1;
        else if (ty == 's' || ty == 'S') t += 2;
        else if (ty == 'i' || ty == 'I' || ty == 'f') t += 4;
        else if (ty == 'Z' || ty == 'H') {
            const uint8_t *z = t; while (t < end && *t) t++;
            if (ty == 'Z' && a == 'M' && (b == 'M' || b == 'm')) { *mm = z; *mm_len = (uint32_t)(t - z); }
            t++;
        } else if (ty == 'B') {
            if (t + 5 > end) return;
            const char sub = (char)t[0]; const uint32_t cnt = ld32(t + 1); t += 5;
            const uint32_t w = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : (sub == 'i' || sub == 'I' || sub == 'f') ? 4 : 0;
            if (!w || (uint64_t)cnt * w > (uint64_t)(end - t)) return;
            if (sub == 'C' && a == 'M' && (b == 'L' || b == 'l')) { *ml = t; *ml_cnt = cnt; }
            t += (uint64_t)cnt * w;
        } else return;
    }
}

WGBS_HD bool passes(const Rec &R, const ViewParams &V) {
    if (V.refid >= 0 && R.refid != V.refid) return false;
    if ((int32_t)R.mapq < V.min_mapq || ((int32_t)R.flag & V.exclude_flags)) return false;
    if (V.include_flags && ((int32_t)R.flag & V.include_flags) != V.include_flags) return false;
    if (V.n_flag_eq) { bool ok = false; for (int k = 0; k < V.n_flag_eq; k++) ok |= (int32_t)R.flag == V.flag_eq[k]; if (!ok) return false; }
    if (V.key_end > 0) { const int64_t key = template_key((int32_t)R.flag, R.refid, R.pos, R.nref, R.npos); if (key < V.key_beg || key >= V.key_end) return false; }
    if (V.end > 0 || V.n_iv) {
        const int64_t pos0 = R.pos;
        int64_t span = 0; const uint8_t *cig = R.cigar();
        for (uint32_t k = 0; k < R.n_cig; k++) { const uint32_t c = ld32(cig + 4 * k), op = c & 15; if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) span += c >> 4; }
        if (span < 1) span = 1;
        if (V.end > 0 && (pos0 + 1 > V.end || pos0 + span < V.beg)) return false;
        if (V.n_iv) {
            uint64_t lo = 0, hi = V.n_iv;                       // first interval ending after pos0
            while (lo < hi) { const uint64_t m = (lo + hi) >> 1; if (V.iv_end[m] <= pos0) lo = m + 1; else hi = m; }
            const bool hit = lo != V.n_iv && V.iv_beg[lo] < pos0 + span;
            if (hit == (V.iv_exclude != 0)) return false;
        }
    }
    if (V.have_rg) {
        uint32_t zl = 0; const uint8_t *z = find_z_tag(R.tags(), R.end(), 'R', 'G', &zl);
        if (!z || zl != V.rg_len) return false;
        for (uint32_t k = 0; k < zl; k++) if ((char)z[k] != V.rg[k]) return false;
    }
    return true;
}

// ---- printf("%g") for a float, exactly (glibc: correctly rounded, ties to even) -----------------------------------------------
// A float is m * 2^e with m < 2^24: the integer part and every fraction digit are produced exactly with a 160-bit integer.
struct Big5 {
    uint32_t w[5];
    WGBS_HD void zero() { for (int i = 0; i < 5; i++) w[i] = 0; }
    WGBS_HD bool is_zero() const { return !(w[0] | w[1] | w[2] | w[3] | w[4]); }
    WGBS_HD void shl(int k) {      // k < 160
        const int ws = k >> 5, bs = k & 31;
        for (int i = 4; i >= 0; i--) {
            uint32_t v = 0;
            if (i - ws >= 0) { v = w[i - ws] << bs; if (bs && i - ws - 1 >= 0) v |= w[i - ws - 1] >> (32 - bs); }
            w[i] = v;
        }
    }
    WGBS_HD uint32_t divmod10() { uint64_t rem = 0; for (int i = 4; i >= 0; i--) { const uint64_t cur = (rem << 32) | w[i]; w[i] = (uint32_t)(cur / 10); rem = cur % 10; } return (uint32_t)rem; }
    WGBS_HD void mul10() { uint64_t c = 0; for (int i = 0; i < 5; i++) { const uint64_t cur = (uint64_t)w[i] * 10 + c; w[i] = (uint32_t)cur; c = cur >> 32; } }
    // take the bits at and above position k (value >> k, < 16 here) and clear them
    WGBS_HD uint32_t take_above(int k) {
        const int ws = k >> 5, bs = k & 31;
        uint32_t v = w[ws] >> bs;
        if (bs && ws + 1 < 5) v |= w[ws + 1] << (32 - bs);
        v &= 15u;
        // clear bits >= k
        w[ws] &= bs ? ((1u << bs) - 1) : 0u;
        for (int i = ws + 1; i < 5; i++) w[i] = 0;
        return v;
    }
    // compare with 2^(k-1): -1 below, 0 equal, 1 above    (k >= 1)
    WGBS_HD int cmp_half(int k) const {
        const int hb = k - 1, ws = hb >> 5, bs = hb & 31;
        for (int i = 4; i > ws; i--) if (w[i]) return 1;
        if (w[ws] >> bs > 1) return 1;
        if (!((w[ws] >> bs) & 1)) return -1;
        if (w[ws] & ((1u << bs) - 1)) return 1;
        for (int i = ws - 1; i >= 0; i--) if (w[i]) return 1;
        return 0;
    }
};

// writes at most 16 chars into out, returns the count
WGBS_HD int fmt_g(float f, char *out) {
    union { float f; uint32_t u; } cv; cv.f = f;
    const uint32_t u = cv.u;
    int n = 0;
    if (u >> 31) out[n++] = '-';
    const uint32_t ex = (u >> 23) & 0xff, fr = u & 0x7fffffu;
    if (ex == 0xff) { const char *s = fr ? "nan" : "inf"; for (int i = 0; i < 3; i++) out[n++] = s[i]; return n; }
    if (ex == 0 && fr == 0) { out[n++] = '0'; return n; }
    const uint32_t m = ex ? (fr | 0x800000u) : fr;
    const int e = (ex ? (int)ex : 1) - 150;                  // value = m * 2^e
    // significant digits d[0..nd) with decimal exponent X of d[0]; `more`: -1/0/1 = the dropped tail vs half a unit of d[nd-1]
    char d[8]; int nd = 0, X = 0, tail = -1;
    if (e >= 0) {
        Big5 b; b.zero(); b.w[0] = m; b.shl(e);
        char all[48]; int na = 0;
        while (!b.is_zero()) all[na++] = (char)b.divmod10();        // least significant first
        X = na - 1;
        for (int i = 0; i < 7 && i < na; i++) d[nd++] = all[na - 1 - i];
        if (na <= 6) tail = -1;
        else {
            // digit 7 onwards vs 0.5
            const int r = all[na - 7];
            bool rest = false; for (int i = 0; i < na - 7; i++) rest |= all[i] != 0;
            tail = r > 5 ? 1 : r < 5 ? -1 : (rest ? 1 : 0);
            nd = 6;
        }
        if (nd > 6) nd = 6;
        // a tail of exact zeros below digit 6 must not round: r == 0 && !rest -> -1 already (r < 5)
    } else {
        const int k = -e;                                            // value = m / 2^k
        uint32_t ip = k < 32 ? (m >> k) : 0;
        Big5 fq; fq.zero(); fq.w[0] = k < 32 ? (m & ((k ? (1u << k) : 1u) - 1u)) : m;
        if (k == 0) fq.w[0] = 0;
        char ipd[10]; int ni = 0;
        while (ip) { ipd[ni++] = (char)(ip % 10); ip /= 10; }
        bool started = ni > 0;
        if (started) { X = ni - 1; for (int i = ni - 1; i >= 0 && nd < 6; i--) d[nd++] = ipd[i]; }
        // ni <= 8 and a float < 2^24 with a fraction has at most 8 integer digits: if nd == 6 and digits of ip remain, fold them into the tail
        if (started && ni > 6) {
            const int r = ipd[ni - 7];
            bool rest = !fq.is_zero(); for (int i = 0; i < ni - 7; i++) rest |= ipd[i] != 0;
            tail = r > 5 ? 1 : r < 5 ? -1 : (rest ? 1 : 0);
        } else {
            int lead = 0;
            while (nd < 6 && !fq.is_zero()) {
                fq.mul10();
                const uint32_t dg = fq.take_above(k);
                if (!started) { lead++; if (dg == 0) continue; started = true; X = -lead; }
                d[nd++] = (char)dg;
            }
            tail = fq.is_zero() ? -1 : fq.cmp_half(k);
            if (fq.is_zero()) tail = -1;
        }
    }
    // round half to even at 6 significant digits
    if (nd == 6 && (tail > 0 || (tail == 0 && (d[5] & 1)))) {
        int i = 5;
        while (i >= 0 && d[i] == 9) d[i--] = 0;
        if (i < 0) { d[0] = 1; X++; } else d[i]++;
    }
    while (nd > 1 && d[nd - 1] == 0) nd--;                            // %g strips trailing zeros
    if (X < -4 || X >= 6) {
        out[n++] = (char)('0' + d[0]);
        if (nd > 1) { out[n++] = '.'; for (int i = 1; i < nd; i++) out[n++] = (char)('0' + d[i]); }
        out[n++] = 'e'; out[n++] = X < 0 ? '-' : '+';
        const int ax = X < 0 ? -X : X;
        out[n++] = (char)('0' + ax / 10); out[n++] = (char)('0' + ax % 10);
    } else if (X >= 0) {
        for (int i = 0; i <= X; i++) out[n++] = (char)('0' + (i < nd ? d[i] : 0));
        if (nd > X + 1) { out[n++] = '.'; for (int i = X + 1; i < nd; i++) out[n++] = (char)('0' + d[i]); }
    } else {
        out[n++] = '0'; out[n++] = '.';
        for (int i = 0; i < -X - 1; i++) out[n++] = '0';
        for (int i = 0; i < nd; i++) out[n++] = (char)('0' + d[i]);
    }
    return n;
}

// ---- one record -> one SAM line ----------------------------------------------------------------------------------------------
// The walker below is the only description of the line; it runs with a counting sink (line length) and with a writing
// sink.  With a multi-lane policy every lane runs the same walk; scalars are stored by lane 0, SEQ and QUAL by all lanes.
struct CountSink {
    uint64_t n = 0;
    WGBS_HD void ch(char) { n++; }
    WGBS_HD void bytes(const uint8_t *, uint64_t k) { n += k; }
    WGBS_HD void seq(const uint8_t *, int32_t l) { n += (uint64_t)l; }
    WGBS_HD void qual(const uint8_t *, int32_t l) { n += (uint64_t)l; }
};
template <class L>
struct WriteSink {
    L lanes; char *o; uint64_t n = 0;
    WGBS_HD void ch(char c) { if (lanes.id() == 0) o[n] = c; n++; }
    WGBS_HD void bytes(const uint8_t *s, uint64_t k) { for (uint64_t i = (uint64_t)lanes.id(); i < k; i += L::N) o[n + i] = (char)s[i]; n += k; }
    WGBS_HD void seq(const uint8_t *s, int32_t l) {
        const char *a = "=ACMGRSVTWYHKDBN";
        for (int32_t i = lanes.id(); i < l; i += L::N) o[n + (uint64_t)i] = a[(s[i >> 1] >> ((~i & 1) << 2)) & 15];
        n += (uint64_t)l;
    }
    WGBS_HD void qual(const uint8_t *s, int32_t l) { for (int32_t i = lanes.id(); i < l; i += L::N) o[n + (uint64_t)i] = (char)(s[i] + 33); n += (uint64_t)l; }
};

template <class S> WGBS_HD void put_u64(S &s, uint64_t v) {
    char t[20]; int k = 0;
    do { t[k++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (k) s.ch(t[--k]);
}
template <class S> WGBS_HD void put_i64(S &s, int64_t v) { if (v < 0) { s.ch('-'); put_u64(s, 0ull - (uint64_t)v); } else put_u64(s, (uint64_t)v); }
template <class S> WGBS_HD void put_ref(S &s, const Refs &F, int32_t id) { s.bytes((const uint8_t *)F.names + F.name_off[id], F.name_off[id + 1] - F.name_off[id]); }
template <class S> WGBS_HD void put_g(S &s, const uint8_t *p) {
    union { float f; uint32_t u; } cv; cv.u = ld32(p);
    char t[16]; const int k = fmt_g(cv.f, t);
    for (int i = 0; i < k; i++) s.ch(t[i]);
}
// one item of a numeric tag / B-array of type ty; returns the item size (0: unknown type)
template <class S> WGBS_HD uint32_t put_num(S &s, char ty, const uint8_t *t) {
    switch (ty) {
        case 'c': put_i64(s, (int8_t)t[0]); return 1;
        case 'C': put_u64(s, t[0]); return 1;
        case 's': put_i64(s, (int16_t)ld16(t)); return 2;
        case 'S': put_u64(s, ld16(t)); return 2;
        case 'i': put_i64(s, ldi32(t)); return 4;
        case 'I': put_u64(s, ld32(t)); return 4;
        case 'f': put_g(s, t); return 4;
        default: return 0;
    }
}

template <class S> WGBS_HD void format_record(const Rec &R, const Refs &F, S &s) {
    const uint8_t *end = R.end();
    { const uint8_t *nm = R.name(); uint32_t k = 0; while (k < R.l_name && nm[k]) k++; s.bytes(nm, k); }      // l_read_name counts the NUL
    s.ch('\t'); put_u64(s, R.flag); s.ch('\t');
    if (R.refid >= 0 && R.refid < F.n) put_ref(s, F, R.refid); else s.ch('*');
    s.ch('\t'); put_i64(s, (int64_t)R.pos + 1); s.ch('\t'); put_u64(s, R.mapq); s.ch('\t');
    if (R.n_cig == 0) s.ch('*');
    else {
        const uint8_t *cg = R.cigar();
        for (uint32_t k = 0; k < R.n_cig; k++) { const uint32_t c = ld32(cg + 4 * k); put_u64(s, c >> 4); s.ch("MIDNSHP=XB??????"[c & 15]); }
    }
    s.ch('\t');
    if (R.nref < 0) s.ch('*'); else if (R.nref == R.refid) s.ch('='); else if (R.nref < F.n) put_ref(s, F, R.nref); else s.ch('*');
    s.ch('\t'); put_i64(s, (int64_t)R.npos + 1); s.ch('\t'); put_i64(s, R.tlen); s.ch('\t');
    if (R.l_seq <= 0) s.ch('*'); else s.seq(R.seq(), R.l_seq);
    s.ch('\t');
    if (R.l_seq <= 0 || R.qual()[0] == 0xff) s.ch('*'); else s.qual(R.qual(), R.l_seq);
    const uint8_t *t = R.tags();
    while (t + 3 <= end) {
        s.ch('\t'); s.ch((char)t[0]); s.ch((char)t[1]); s.ch(':');
        const char ty = (char)t[2]; t += 3;
        if (ty == 'A') { s.ch('A'); s.ch(':'); s.ch((char)t[0]); t += 1; }
        else if (ty == 'c' || ty == 'C' || ty == 's' || ty == 'S' || ty == 'i' || ty == 'I') { s.ch('i'); s.ch(':'); t += put_num(s, ty, t); }
        else if (ty == 'f') { s.ch('f'); s.ch(':'); t += put_num(s, ty, t); }
        else if (ty == 'Z' || ty == 'H') {
            s.ch(ty); s.ch(':');
            const uint8_t *z = t; while (t < end && *t) t++;
            s.bytes(z, (uint64_t)(t - z)); t++;
        } else if (ty == 'B') {
            const char sub = (char)t[0]; const uint32_t cnt = ld32(t + 1); t += 5;
            s.ch('B'); s.ch(':'); s.ch(sub);
            for (uint32_t k = 0; k < cnt && t < end; k++) {
                s.ch(',');
                const uint32_t w = put_num(s, sub, t);
                if (!w) { t = end; break; }
                t += w;
            }
        } else t = end;                                           // unknown type: stop (malformed)
    }
    s.ch('\n');
}

}  // namespace bamcore
