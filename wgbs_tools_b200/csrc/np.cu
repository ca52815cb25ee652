// np.cu -- modification-aware pileup: MM/ML tags (5mC 'm', 5hmC 'h', Biomodal 'C+C') -> C / T / H / . calls.
//
// Reference behaviour restated (reference src/pipeline_wgbs/ont.cpp):
//   :418-438 get_np_tags            (tokenizer, sam.cu: last MM:Z:/Mm:Z: and ML:B:C/Ml:B:C field)
//   :310-333 find_Cm_substring, :361-416 subset_to_Cm_section, :269-308 parse_np_fields_by_mod   -> Section / parse_section
//   :223-267 parse_np_fields (C+C? merge, np_dot bookkeeping)                                    -> NpTags::load
//   :22-87   make_meth_mask                                                                     -> resolve_c
//   :90-221  np_samLineToPatVec                                                                 -> np_measure_k / np_call_k
//
// The reference builds a per-base mask string over the read in its ORIGINAL orientation and pushes it through
// clean_CIGAR.  Here nothing per-base is materialised: for each CpG candidate the call kernel finds the read base through
// the CIGAR, computes the ordinal of that C among the C's of the original-orientation read (a running count of 'C'
// for top-strand reads, of 'G' from the far end for bottom-strand reads), and looks the ordinal up in the MM delta
// lists, which are walked in place in the tag text.
#include "pileup_dev.cuh"

namespace {

// istream >> int then "if (peek == ',') ignore" (split_by_comma, patter_utils.cpp:83-94): parse one int at p; false = stop
__device__ __forceinline__ bool next_int(const char *__restrict__ t, uint32_t &p, uint32_t e, int32_t *v) {
    uint32_t q = p;
    while (q < e && (t[q] == ' ' || (t[q] >= 9 && t[q] <= 13))) q++;
    bool neg = false;
    if (q < e && (t[q] == '+' || t[q] == '-')) { neg = t[q] == '-'; q++; }
    if (q >= e || t[q] < '0' || t[q] > '9') return false;
    int64_t x = 0;
    while (q < e && t[q] >= '0' && t[q] <= '9') { x = x * 10 + (t[q] - '0'); if (x > 0x80000000LL) return false; q++; }
    if (neg) x = -x;
    if (x > 0x7fffffffLL || x < -0x80000000LL) return false;
    if (q < e && t[q] == ',') q++;
    *v = (int32_t)x; p = q;
    return true;
}
__device__ __forceinline__ uint32_t count_ints(const char *__restrict__ t, uint32_t p, uint32_t e) {
    uint32_t n = 0; int32_t v;
    while (next_int(t, p, e, &v)) n++;
    return n;
}
// text after the first comma of [p,e) (trim_from_first_comma, ont.cpp:335-344); empty range when there is none
__device__ __forceinline__ uint32_t after_comma(const char *__restrict__ t, uint32_t p, uint32_t e) {
    while (p < e && t[p] != ',') p++;
    return p < e ? p + 1 : e;
}

// one "C+x" section of the MM string and the slice of ML values that belongs to it
struct Section {
    bool found, dot, bad;       // bad: the reference throws -> read invalid
    uint32_t dp, de;            // delta list text
    uint32_t nr;                // number of deltas
    bool has_ml; uint32_t mlp, mle; uint32_t ml_base;   // ML values text; this section's values start at index ml_base
    bool ml_bin;                // BAM flavour: the values are the bytes of the B:C array [mlp, mle), not decimal text
};

// BAM flavour (direct route, bamdev.cu bam_np_tags_k): mm = the Z string as it stands in the record (the same bytes SAM text shows
// behind "MM:Z:"), ml = the uint8 values of the B:C array, ml_len = their number + 1 (0: no ML tag)
__device__ Section parse_section(const char *__restrict__ t, uint32_t mm, uint32_t mm_len, uint32_t ml, uint32_t ml_len, char mod, bool bam) {
    Section s; s.ml_bin = bam; s.found = false; s.dot = false; s.bad = false; s.dp = s.de = 0; s.nr = 0; s.has_ml = false; s.mlp = s.mle = 0; s.ml_base = 0;
    if (mm_len == 0) return s;                                   // no MM tag -> get_np_tags false -> empty lists
    const uint32_t e = mm + mm_len;
    uint32_t st = mm, pos = 0;
    for (uint32_t i = mm; i <= e; i++) {
        if (i == e || t[i] == ';') {
            if (i == e && st == e) break;                         // getline: nothing after the last ';'
            if (i - st >= 3 && t[st] == 'C' && t[st + 1] == '+' && t[st + 2] == mod) { s.found = true; s.dp = st; s.de = i; break; }
            pos++; st = i + 1;
        }
    }
    if (!s.found) return s;
    s.dot = !((s.de - s.dp > 3) && t[s.dp + 3] == '?');           // "C+m." / "C+m" vs "C+m?"  (ont.cpp:386)
    s.dp = after_comma(t, s.dp, s.de);
    s.nr = count_ints(t, s.dp, s.de);
    if (ml_len == 0) return s;                                    // no ML: every listed base has ML 255 (ont.cpp:292-294)
    const uint32_t vp = bam ? ml : after_comma(t, ml, ml + ml_len), ve = bam ? ml + ml_len - 1 : ml + ml_len;
    const uint32_t total = bam ? ml_len - 1 : count_ints(t, vp, ve);
    if (s.nr == 0) return s;                                      // ML_str = "" (ont.cpp:398-401)
    if ((total % s.nr != 0) && total > 0) { s.bad = true; return s; }          // :403-407
    s.has_ml = true; s.mlp = vp; s.mle = ve;
    if (total >= (pos + 1) * s.nr) s.ml_base = pos * s.nr;        // :411-415 slice
    else if (total != s.nr) s.bad = true;                         // un-sliced ML must match the delta count (:296-299)
    return s;
}

// forward walker over one section: absolute C ordinals (pos += delta; ordinal = pos++) with their ML values
struct ModWalk {
    const char *t; uint32_t dp, de, mlp, mle; bool has_ml, ml_bin; uint32_t left; int64_t pos; int64_t next; int32_t ml; bool valid;
    __device__ void init(const char *text, const Section &s) {
        t = text; dp = s.dp; de = s.de; has_ml = s.has_ml; ml_bin = s.ml_bin; mlp = s.mlp; mle = s.mle; left = s.found && !s.bad ? s.nr : 0; pos = 0; valid = false; ml = 255; next = -1;
        if (has_ml && ml_bin) mlp += s.ml_base;
        else if (has_ml) { int32_t v; for (uint32_t k = 0; k < s.ml_base; k++) next_int(t, mlp, mle, &v); }
        step();
    }
    __device__ void step() {
        if (!left) { valid = false; return; }
        int32_t d = 0; next_int(t, dp, de, &d);
        pos += d; next = pos++; left--;
        ml = 255;
        if (has_ml && ml_bin) { ml = mlp < mle ? (int32_t)(uint8_t)t[mlp] : 0; mlp++; }
        else if (has_ml) { int32_t v = 0; next_int(t, mlp, mle, &v); ml = v; }
        valid = true;
    }
    // is ordinal x listed?  x must not decrease between calls
    __device__ bool at(int64_t x, int32_t *mlv) {
        while (valid && next < x) step();
        if (valid && next == x) { *mlv = ml; return true; }
        return false;
    }
};

struct NpTags {
    Section h, m, c; bool np_dot, bad;
    __device__ void load(const ReadBatchView &rb, uint32_t r) {
        const char *t = rb.text;
        const bool bam = rb.bam != 0;
        np_dot = false;
        h = parse_section(t, rb.mm_off[r], rb.mm_len[r], rb.ml_off[r], rb.ml_len[r], 'h', bam); if (h.found) np_dot = h.dot;
        m = parse_section(t, rb.mm_off[r], rb.mm_len[r], rb.ml_off[r], rb.ml_len[r], 'm', bam); if (m.found) np_dot = m.dot;
        c = parse_section(t, rb.mm_off[r], rb.mm_len[r], rb.ml_off[r], rb.ml_len[r], 'C', bam);   // np_dot is restored to the C+m value (ont.cpp:263)
        bad = h.bad || m.bad || c.bad;
    }
};

// call for the C with ordinal x (make_meth_mask, ont.cpp:22-87) -> pat symbol code.
__device__ uint32_t resolve_c(int64_t x, ModWalk &wh, ModWalk &wm, ModWalk &wc, const PileupOpts &o, bool np_dot) {
    const float hi_t = 255 * o.np_thresh, lo_t = 255 * (1 - o.np_thresh);
    int32_t vh = 0, vm = 0, vc = 0;
    bool in_h = wh.at(x, &vh), in_m = wm.at(x, &vm), in_c = wc.at(x, &vc);
    // C+C? positions are merged into the 5mC (cpc_call 'C') or 5hmC ('H') list with ML 255 unless already present (ont.cpp:246-260)
    if (in_c && o.cpc_call == 'C' && !in_m) { in_m = true; vm = 255; }
    if (in_c && o.cpc_call == 'H' && !in_h) { in_h = true; vh = 255; }
    char mask = 'E';
    if (o.combine_mods) {
        if (in_h || in_m) {
            int comb = (in_h ? vh : 0) + (in_m ? vm : 0); if (comb > 255) comb = 255;
            mask = comb > hi_t ? 'M' : (comb < lo_t ? 'U' : 'N');
        }
    } else {
        char cur = 'N';
        if (in_h) { cur = vh > hi_t ? 'H' : (vh < lo_t ? 'U' : 'N'); mask = cur; }
        if (in_m) {
            if (vm > hi_t) cur = 'M';
            else if (vm < lo_t) { if (cur != 'H') cur = 'U'; }
            else if (cur != 'H') cur = 'N';
            mask = cur;
        }
    }
    if (mask == 'M') return SYM_C;
    if (mask == 'U') return SYM_T;
    if (mask == 'H') return SYM_H;
    if (mask == 'E') return np_dot ? SYM_T : SYM_DOT;             // unlisted C under the implicit-unmodified convention (ont.cpp:168-173)
    return SYM_DOT;                                                // 'N'
}

__global__ void __launch_bounds__(128) np_measure_k(ReadBatchView rb, const uint32_t *__restrict__ loci, uint32_t nloci, PileupOpts o,
                                                     uint32_t *__restrict__ r_lo, uint32_t *__restrict__ r_ncand, uint32_t *__restrict__ words,
                                                     unsigned long long *__restrict__ stats) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t inval = 0, empty = 0;
    if (r < rb.n) {
        uint32_t lo = 0, nc = 0;
        const uint8_t st = rb.status[r];
        if (st == REC_INVALID) inval = 1;
        else if (st == REC_OK || st == REC_BADINT) {
            NpTags tg; tg.load(rb, r);
            if (tg.bad) inval = 1;
            else {
                const bool m_empty = !(tg.m.found && tg.m.nr) && !(o.cpc_call == 'C' && tg.c.found && tg.c.nr);
                const bool h_empty = !(tg.h.found && tg.h.nr) && !(o.cpc_call == 'H' && tg.c.found && tg.c.nr);
                const bool seq_star = rb.bam ? rb.seq_off[r] == SEQ_STAR : (rb.seq_len[r] == 1 && rb.text[rb.seq_off[r]] == '*');
                if ((m_empty && h_empty && !tg.np_dot) || seq_star) empty = 1;              // ont.cpp:97-100
                else if (st == REC_BADINT) inval = 1;                                        // stoul / stoi throw (ont.cpp:102-103)
                else {
                    int64_t span = 0;
                    bool ok = rec_cig_validate(rb, r, &span);
                    const bool bottom = (rb.flag[r] & 0x10) == 16;
                    if (ok && bottom) {                                                      // reverse_comp throws on anything but ACGTN
                        const uint32_t so = rb.seq_off[r];
                        for (uint32_t q = 0; q < rb.seq_len[r]; q++) { char ch = rec_base(rb, so, q); if (ch != 'A' && ch != 'C' && ch != 'G' && ch != 'T' && ch != 'N') { ok = false; break; } }
                    }
                    if (!ok) inval = 1;
                    else {
                        const int64_t pos = (int64_t)(((uint64_t)(uint32_t)rb.pos_hi[r] << 32) | (uint32_t)rb.pos[r]);   // 64-bit POS (ont.cpp:102)
                        // bottom-strand reads also see the CpG whose G is their first aligned base (ont.cpp:142-145)
                        lo = lower_bound_u32(loci, nloci, bottom ? pos - 1 : pos);
                        uint32_t hi = lower_bound_u32(loci, nloci, pos + span);
                        nc = hi > lo ? hi - lo : 0;
                    }
                }
            }
        }
        r_lo[r] = lo; r_ncand[r] = (inval || empty) ? NONE : nc;
        words[r] = (inval || empty) ? 0 : (nc + 15) >> 4;
    }
    for (int d = 16; d >= 1; d >>= 1) { inval += __shfl_xor_sync(0xffffffffu, inval, d); empty += __shfl_xor_sync(0xffffffffu, empty, d); }
    if ((threadIdx.x & 31) == 0) { if (inval) atomicAdd(&stats[ST_INVALID], (unsigned long long)inval); if (empty) atomicAdd(&stats[ST_EMPTY], (unsigned long long)empty); }
}

// scratch word per candidate: 0 = '.', else 1 + ordinal of the C (original orientation)
__global__ void __launch_bounds__(128) np_call_k(ReadBatchView rb, const uint32_t *__restrict__ loci, uint32_t first_idx, PileupOpts o,
                                                  const uint32_t *__restrict__ r_lo, const uint32_t *__restrict__ r_ncand, const uint32_t *__restrict__ off,
                                                  uint32_t *__restrict__ pool, uint32_t *__restrict__ scratch, int32_t *__restrict__ r_idx,
                                                  uint32_t *__restrict__ r_len, unsigned long long *__restrict__ stats) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t empty = 0;
    if (r < rb.n) { r_idx[r] = 0; r_len[r] = 0; }
    if (r < rb.n && rb.status[r] == REC_OK && r_ncand[r] != NONE) {
        const uint32_t nc = r_ncand[r], lo = r_lo[r];
        const int64_t pos = (int64_t)(((uint64_t)(uint32_t)rb.pos_hi[r] << 32) | (uint32_t)rb.pos[r]);
        const bool bottom = (rb.flag[r] & 0x10) == 16;
        int64_t span; rec_cig_validate(rb, r, &span);
        const uint32_t so = rb.seq_off[r];
        const uint32_t slen = rb.seq_len[r];
        uint32_t *sc = scratch + (size_t)off[r] * 16;
        // phase A (ascending): candidate -> ordinal of its C in the original-orientation read
        int64_t totalG = 0;
        if (bottom) for (uint32_t q = 0; q < slen; q++) totalG += rec_base(rb, so, q) == 'G';
        CigCursor cc; cc.init(rb, r);
        int64_t qscan = 0, run = 0;                  // run = # of 'C' (top) / 'G' (bottom) in seq[0, qscan)
        const char want = bottom ? 'G' : 'C';
        for (uint32_t j = 0; j < nc; j++) {
            const int64_t i = (int64_t)loci[lo + j] - pos;       // -1 possible for bottom reads
            const int64_t di = bottom ? i + 1 : i;
            uint32_t v = 0;
            if (di < span && cc.seek(di) && cc.op == 'M') {      // di >= mask.size(): skipped; deleted base: 'N' -> '.'
                const int64_t q = cc.q0 + (di - cc.r0);
                while (qscan < q) { run += rec_base(rb, so, qscan) == want; qscan++; }
                if (rec_base(rb, so, q) == want) {
                    const int64_t ord = bottom ? totalG - run - 1 : run;   // G's after q / C's before q
                    const int64_t clip_pos = di;                             // ont.cpp:195-199 (di for bottom, i == di for top)
                    if ((clip_pos >= o.clip) && (clip_pos < span - o.clip)) v = (uint32_t)ord + 1;
                }
            }
            sc[j] = v;
        }
        // phase B: resolve ordinals against the MM lists in ascending ordinal order (descending j for bottom reads)
        NpTags tg; tg.load(rb, r);
        ModWalk wh, wm, wc; wh.init(rb.text, tg.h); wm.init(rb.text, tg.m); wc.init(rb.text, tg.c);
        if (o.cpc_call == '.') wc.left = 0, wc.valid = false;
        for (uint32_t k = 0; k < nc; k++) {
            const uint32_t j = bottom ? nc - 1 - k : k;
            const uint32_t v = sc[j];
            sc[j] = v ? resolve_c((int64_t)v - 1, wh, wm, wc, o, tg.np_dot) : SYM_DOT;
        }
        // pack, dropping leading / trailing '.'
        int32_t first = -1, last = -1;
        for (uint32_t j = 0; j < nc; j++) if (sc[j]) { if (first < 0) first = (int32_t)j; last = (int32_t)j; }
        if (first < 0) empty = 1;
        else {
            uint32_t *wp = pool + off[r]; uint32_t w = 0; int ns = 0;
            for (int32_t j = first; j <= last; j++) { w |= sc[j] << (30 - 2 * (ns & 15)); if ((++ns & 15) == 0) { *wp++ = w; w = 0; } }
            if (ns & 15) *wp = w;
            r_idx[r] = (int32_t)(first_idx + lo + (uint32_t)first); r_len[r] = (uint32_t)(last - first + 1);
        }
    }
    for (int d = 16; d >= 1; d >>= 1) empty += __shfl_xor_sync(0xffffffffu, empty, d);
    if ((threadIdx.x & 31) == 0 && empty) atomicAdd(&stats[ST_EMPTY], (unsigned long long)empty);
}

}  // namespace

int np_measure(wgbs_ctx *ctx, const ReadBatch &rb, const uint32_t *loci, uint32_t nloci, PileupOpts o, uint32_t *r_lo, uint32_t *r_ncand,
               uint32_t *words, unsigned long long *d_stats) {
    LAUNCH(ctx, np_measure_k, grid_for(rb.n, 128), 128, 0, view_of(rb), loci, nloci, o, r_lo, r_ncand, words, d_stats);
    LAUNCH_CHECK();
    return 0;
}

int np_call(wgbs_ctx *ctx, const ReadBatch &rb, const uint32_t *loci, uint32_t first_idx, PileupOpts o, const uint32_t *r_lo,
            const uint32_t *r_ncand, const uint32_t *off, uint32_t *pool, uint32_t pool_words, int32_t *r_idx, uint32_t *r_len,
            unsigned long long *d_stats) {
    Temps T(ctx);
    uint32_t *scratch;
    RC_TRY(T.alloc(&scratch, (size_t)pool_words * 16));
    LAUNCH(ctx, np_call_k, grid_for(rb.n, 128), 128, 0, view_of(rb), loci, first_idx, o, r_lo, r_ncand, off, pool, scratch, r_idx, r_len, d_stats);
    LAUNCH_CHECK();
    return 0;
}
