// sam.cu -- SAM text (the stdin of the reference's `patter`, pipeline_wgbs/patter.cpp:381-416) -> ReadBatch.
//
// One warp per line: 16 bytes per lane per iteration (512 B per warp-iteration, coalesced), tab positions found with
// byte-compare masks + a warp prefix sum.  Replaces line2tokens (pipeline_wgbs/patter_utils.cpp:9-18) and the
// per-field std::stoi calls; also produces the 64-bit QNAME hash that template pairing sorts on.
#include "lines.cuh"
#include "reads.cuh"

namespace {

constexpr int TK_T = 256, TK_WARPS = TK_T / 32;
constexpr int NTAB = 11;   // tab ordinals 0..10 delimit the 11 mandatory fields

__device__ __forceinline__ uint64_t fmix64(uint64_t k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
    return k;
}

// std::stoi semantics on text[s,e): optional blanks, sign, >=1 digit, int range; trailing junk ignored
__device__ __forceinline__ bool parse_i32(const char *__restrict__ t, uint32_t s, uint32_t e, int32_t *out) {
    while (s < e && (t[s] == ' ' || (t[s] >= 9 && t[s] <= 13))) s++;
    bool neg = false;
    if (s < e && (t[s] == '+' || t[s] == '-')) { neg = t[s] == '-'; s++; }
    if (s >= e || t[s] < '0' || t[s] > '9') return false;
    int64_t v = 0;
    while (s < e && t[s] >= '0' && t[s] <= '9') { v = v * 10 + (t[s] - '0'); if (v > 0x80000000LL) return false; s++; }
    if (neg) v = -v;
    if (v > 0x7fffffffLL || v < -0x80000000LL) return false;
    *out = (int32_t)v;
    return true;
}

__device__ __forceinline__ bool is_tag(const char *__restrict__ t, uint32_t p, uint32_t e, char c0, char c1a, char c1b, char c2, char c3, char c4, char c5, int n) {
    if (p + n > e) return false;
    if (t[p] != c0 || (t[p + 1] != c1a && t[p + 1] != c1b) || t[p + 2] != c2 || t[p + 3] != c3 || t[p + 4] != c4) return false;
    if (n == 6 && t[p + 5] != c5) return false;
    return true;
}

// per-(byte, position) mixing summed over the name: independent of how the name is aligned to the 16-byte chunk grid
__device__ __forceinline__ uint64_t name_byte_mix(uint32_t byte, uint32_t pos) {
    uint64_t x = ((uint64_t)(byte | (pos << 8)) + 1) * 0x9e3779b97f4a7c15ULL;
    x ^= x >> 29; x *= 0xbf58476d1ce4e5b9ULL; x ^= x >> 32;
    return x;
}

__global__ void __launch_bounds__(TK_T) sam_fields_k(const char *__restrict__ text, uint32_t n, const uint32_t *__restrict__ nlpos,
                                                      uint32_t n_nl, uint32_t n_lines, int want_tags, ReadBatch rb) {
    __shared__ uint32_t tabs[TK_WARPS][NTAB];
    const unsigned w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t line = blockIdx.x * TK_WARPS + w;
    if (line >= n_lines) return;
    const uint32_t s = line == 0 ? 0 : nlpos[line - 1] + 1;
    const uint32_t e = line < n_nl ? nlpos[line] : n;
    uint32_t ntab = 0;
    uint64_t h = 0;
    uint32_t qend = e;                     // end of QNAME = first tab (or the line end)
    uint32_t mm_off = 0, mm_len = 0, ml_off = 0, ml_len = 0;
    // chunks are aligned to the 16-byte grid of the (256 B aligned) text buffer: every load is one aligned LDG.128
    const uint32_t a0 = s & ~15u;
    for (uint32_t base = a0; base < e; base += 512) {
        const uint32_t p = base + lane * 16;
        uint4 v = make_uint4(0, 0, 0, 0);
        uint32_t tm = 0, valid = 0;
        if (p < e && p + 16 > s) {
            v = (p + 16 <= n) ? *reinterpret_cast<const uint4 *>(text + p) : load16_guard(text, p, n);
            valid = 0xffffu;
            if (p < s) valid &= 0xffffu << (s - p);
            if (e - p < 16) valid &= (1u << (e - p)) - 1;
            tm = eq_mask16(v, '\t') & valid;
        }
        const uint32_t c = __popc(tm);
        uint32_t inc = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= (unsigned)d) inc += t; }
        uint32_t ord = ntab + inc - c;
        uint32_t m = tm;
        while (m) {
            int b = __ffs(m) - 1; m &= m - 1;
            if (ord < NTAB) tabs[w][ord] = p + b;
            if (want_tags && ord >= 10) {                           // a field with index >= 11 starts after this tab
                uint32_t f = p + b + 1;
                bool mm = is_tag(text, f, e, 'M', 'M', 'm', ':', 'Z', ':', 0, 5);
                bool ml = !mm && is_tag(text, f, e, 'M', 'L', 'l', ':', 'B', ':', 'C', 6);
                if (mm || ml) {
                    uint32_t q = f + (mm ? 5 : 6), z = q;
                    while (z < e && text[z] != '\t') z++;
                    if (mm) { mm_off = q; mm_len = z - q; } else { ml_off = q; ml_len = z - q; }
                }
            }
            ord++;
        }
        // QNAME: bytes before the first tab of the line
        if (ntab == 0) {
            const uint32_t first = __ballot_sync(0xffffffffu, c != 0);
            uint32_t stop = e;                                       // exclusive end of name bytes seen so far
            if (first) { const int fl = __ffs(first) - 1; const uint32_t ftm = __shfl_sync(0xffffffffu, tm, fl); stop = base + fl * 16 + (__ffs(ftm) - 1); }
            if (first) qend = stop;
            uint32_t w4[4] = {v.x, v.y, v.z, v.w};
            uint32_t nm = valid;
            if (p + 16 > stop) nm &= p >= stop ? 0u : ((1u << (stop - p)) - 1);
            while (nm) { int b = __ffs(nm) - 1; nm &= nm - 1; h += name_byte_mix((w4[b >> 2] >> ((b & 3) * 8)) & 255u, p + b - s); }
        }
        ntab += __shfl_sync(0xffffffffu, inc, 31);
    }
    __syncwarp();
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) h += __shfl_xor_sync(0xffffffffu, h, d);
    h = fmix64(h ^ (uint64_t)(qend - s));
    if (e == s) h = fmix64(0x5851f42d4c957f2dULL ^ (uint64_t)line);   // blank lines: unique keys, never paired
    // tags: the last occurrence wins (get_np_tags, ont.cpp:418-438): highest lane that saw one
    if (want_tags) {
        uint32_t bm = __ballot_sync(0xffffffffu, mm_len || mm_off);
        if (bm) { int src = 31 - __clz(bm); mm_off = __shfl_sync(0xffffffffu, mm_off, src); mm_len = __shfl_sync(0xffffffffu, mm_len, src); }
        uint32_t bl = __ballot_sync(0xffffffffu, ml_len || ml_off);
        if (bl) { int src = 31 - __clz(bl); ml_off = __shfl_sync(0xffffffffu, ml_off, src); ml_len = __shfl_sync(0xffffffffu, ml_len, src); }
    }
    if (lane == 0) {
        uint8_t st = REC_OK;
        int32_t flag = 0, pos = 0;
        uint32_t cig_off = s, cig_len = 0, seq_off = s, seq_len = 0;
        if (e == s) st = REC_BLANK;
        else {
            // number of tokens line2tokens would produce: tabs+1, minus one when the line ends with a tab
            uint32_t nfields = ntab + 1 - ((text[e - 1] == '\t') ? 1 : 0);
            if (nfields < 11) st = REC_INVALID;
            else {
                const uint32_t *tb = tabs[w];
                if (!parse_i32(text, tb[0] + 1, tb[1], &flag) || !parse_i32(text, tb[2] + 1, tb[3], &pos)) st = REC_BADINT;
                cig_off = tb[4] + 1; cig_len = tb[5] - tb[4] - 1;
                seq_off = tb[8] + 1; seq_len = tb[9] - tb[8] - 1;
            }
        }
        rb.line_off[line] = s; rb.line_len[line] = e - s; rb.qn_len[line] = qend - s;
        rb.flag[line] = flag; rb.pos[line] = pos;
        rb.cig_off[line] = cig_off; rb.cig_len[line] = cig_len; rb.seq_off[line] = seq_off; rb.seq_len[line] = seq_len;
        rb.hash_lo[line] = (uint32_t)h; rb.hash_hi[line] = (uint32_t)(h >> 32);
        rb.status[line] = st;
        if (want_tags) { rb.mm_off[line] = mm_off; rb.mm_len[line] = mm_len; rb.ml_off[line] = ml_off; rb.ml_len[line] = ml_len; }
    }
}

}  // namespace

int sam_tokenize(wgbs_ctx *ctx, const char *dtext, size_t nbytes, bool want_tags, Temps &T, ReadBatch *out) {
    if (nbytes >= 0xfffffff0ull) return wgbs_set_err("SAM text must be < 4 GiB per call (got %zu); split on line boundaries", nbytes);
    uint32_t *nlpos = nullptr, n_nl = 0, n_lines = 0;
    RC_TRY(find_lines(ctx, dtext, nbytes, T, &nlpos, &n_nl, &n_lines));
    ReadBatch rb;
    rb.text = dtext; rb.nbytes = (uint32_t)nbytes; rb.n = n_lines;
    RC_TRY(T.alloc(&rb.line_off, n_lines)); RC_TRY(T.alloc(&rb.line_len, n_lines)); RC_TRY(T.alloc(&rb.qn_len, n_lines));
    RC_TRY(T.alloc(&rb.flag, n_lines)); RC_TRY(T.alloc(&rb.pos, n_lines));
    RC_TRY(T.alloc(&rb.cig_off, n_lines)); RC_TRY(T.alloc(&rb.cig_len, n_lines));
    RC_TRY(T.alloc(&rb.seq_off, n_lines)); RC_TRY(T.alloc(&rb.seq_len, n_lines));
    RC_TRY(T.alloc(&rb.hash_lo, n_lines)); RC_TRY(T.alloc(&rb.hash_hi, n_lines));
    RC_TRY(T.alloc(&rb.status, n_lines));
    if (want_tags) {
        RC_TRY(T.alloc(&rb.mm_off, n_lines)); RC_TRY(T.alloc(&rb.mm_len, n_lines));
        RC_TRY(T.alloc(&rb.ml_off, n_lines)); RC_TRY(T.alloc(&rb.ml_len, n_lines));
    }
    if (n_lines) {
        LAUNCH(ctx, sam_fields_k, (n_lines + TK_WARPS - 1) / TK_WARPS, TK_T, 0, dtext, (uint32_t)nbytes, nlpos, n_nl, n_lines, want_tags ? 1 : 0, rb);
        LAUNCH_CHECK();
    }
    *out = rb;
    return 0;
}
