// pileup_dev.cuh -- device helpers shared by the WGBS (pileup.cu) and MM/ML (np.cu) call kernels.
#pragma once
#include "reads.cuh"

constexpr uint32_t NONE = 0xffffffffu;
constexpr int MAX_PE_PAT_LEN = 300;     // patter_utils.h:21

struct PileupOpts {
    int min_cpg, clip, paired, nanopore, combine_mods;
    float np_thresh;
    char cpc_call;
};

__device__ __forceinline__ bool is_bottom(int flag, int paired) {
    if (paired) return ((flag & 0x53) == 83) || ((flag & 0xA3) == 163);
    return (flag & 0x10) == 16;
}

// ---- CIGAR ---------------------------------------------------------------------------------------------------------
// clean_CIGAR throws (-> read counted invalid) when: an op char has no number before it, the number overflows int,
// the op is not one of M = X D N I S H, or an M/=/X/I/S op consumes more read bases than remain.
static __device__ bool cig_validate(const char *__restrict__ t, uint32_t p, uint32_t e, uint32_t seq_len, int64_t *span_out) {
    int64_t span = 0, q = 0;
    bool ok = true;
    uint64_t num = 0; bool have = false;
    for (; p < e; p++) {
        char c = t[p];
        if (c >= '0' && c <= '9') { num = num * 10 + (uint64_t)(c - '0'); if (num > 0x7fffffffull) num = 0x80000000ull; have = true; continue; }
        if (!have || num > 0x7fffffffull) { ok = false; break; }
        if (c == 'M' || c == '=' || c == 'X' || c == 'I' || c == 'S') {
            if ((int64_t)num > (int64_t)seq_len - q) ok = false;       // checked after the tokenising phase in the reference; same verdict
            q += (int64_t)num;
            if (c != 'I' && c != 'S') span += (int64_t)num;
        } else if (c == 'D' || c == 'N') span += (int64_t)num;
        else if (c == 'H') {}
        else ok = false;
        num = 0; have = false;
    }
    // an M op that overruns still appended the available bases before throwing; irrelevant: the read is invalid
    *span_out = span;
    return ok;
}

// ---- BAM flavour (reads.cuh): binary CIGAR = n little-endian words (length << 4 | op), ops "MIDNSHP=X" = 0..8 -----------------
__device__ __forceinline__ uint32_t bam_ld32(const char *__restrict__ t, uint32_t off) {
    const uint8_t *p = (const uint8_t *)t + off;
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
// the verdict clean_CIGAR reaches on the TEXT of this CIGAR: no ops ("*") and ops other than M I D N S H = X (P, B, reserved) throw
static __device__ bool cig_validate_bam(const char *__restrict__ t, uint32_t off, uint32_t n_ops, uint32_t seq_len, int64_t *span_out) {
    int64_t span = 0, q = 0;
    bool ok = n_ops != 0;
    for (uint32_t k = 0; k < n_ops; k++) {
        const uint32_t v = bam_ld32(t, off + 4 * k), op = v & 15; const int64_t num = (int64_t)(v >> 4);
        if (op == 0 || op == 7 || op == 8 || op == 1 || op == 4) {
            if (num > (int64_t)seq_len - q) ok = false;
            q += num;
            if (op != 1 && op != 4) span += num;
        } else if (op == 2 || op == 3) span += num;
        else if (op == 5) {}
        else ok = false;
    }
    *span_out = span;
    return ok;
}
// one base of SEQ as the character SAM text shows ("=ACMGRSVTWYHKDBN"; SEQ "*" is the one character '*')
__device__ __forceinline__ char bam_base(const char *__restrict__ t, uint32_t seq_off, int64_t q) {
    if (seq_off == SEQ_STAR) return '*';
    const uint32_t b = ((const uint8_t *)t)[seq_off + (uint32_t)(q >> 1)];
    return "=ACMGRSVTWYHKDBN"[(q & 1) ? (b & 15) : (b >> 4)];
}
// per-record front ends used by the kernels: pick the flavour (uniform over the whole grid)
static __device__ __forceinline__ bool rec_cig_validate(const ReadBatchView &rb, uint32_t r, int64_t *span) {
    return rb.bam ? cig_validate_bam(rb.text, rb.cig_off[r], rb.cig_len[r], rb.seq_len[r], span)
                  : cig_validate(rb.text, rb.cig_off[r], rb.cig_off[r] + rb.cig_len[r], rb.seq_len[r], span);
}
__device__ __forceinline__ char rec_base(const ReadBatchView &rb, uint32_t seq_off, int64_t q) {
    return rb.bam ? bam_base(rb.text, seq_off, q) : rb.text[seq_off + q];
}

struct CigCursor {
    const char *t; uint32_t p, e;
    int64_t r0, q0, len;   // current ref-consuming op covers ref offsets [r0, r0+len); its first read base is q0 (op 'M')
    char op;               // 'M' (M,=,X) or 'D' (D,N); 0 before the first op
    bool bin;              // BAM flavour: p / e count bytes of 4-byte op words
    __device__ void init(const char *text, uint32_t off, uint32_t n) { t = text; p = off; e = off + n; r0 = 0; q0 = 0; len = 0; op = 0; bin = false; }
    __device__ void init(const ReadBatchView &rb, uint32_t r) {
        t = rb.text; p = rb.cig_off[r]; bin = rb.bam != 0; e = p + (bin ? 4 * rb.cig_len[r] : rb.cig_len[r]); r0 = 0; q0 = 0; len = 0; op = 0;
    }
    // position on the op covering ref offset x (x must be non-decreasing across calls); false past the end
    __device__ bool seek(int64_t x) {
        while (true) {
            if (x < r0 + len) return true;
            // leave the current op
            r0 += len; if (op == 'M') q0 += len;
            len = 0; op = 0;
            // next op
            bool found = false;
            while (bin && p < e) {
                const uint32_t v = bam_ld32(t, p), o4 = v & 15; const int64_t num = (int64_t)(v >> 4);
                p += 4;
                if (o4 == 0 || o4 == 7 || o4 == 8) { op = 'M'; len = num; found = true; break; }
                if (o4 == 2 || o4 == 3) { op = 'D'; len = num; found = true; break; }
                if (o4 == 1 || o4 == 4) q0 += num;
            }
            while (!bin && p < e) {
                int64_t num = 0;
                while (p < e && t[p] >= '0' && t[p] <= '9') { num = num * 10 + (t[p] - '0'); p++; }
                if (p >= e) break;
                char c = t[p++];
                if (c == 'M' || c == '=' || c == 'X') { op = 'M'; len = num; found = true; break; }
                if (c == 'D' || c == 'N') { op = 'D'; len = num; found = true; break; }
                if (c == 'I' || c == 'S') q0 += num;
            }
            if (!found) return false;
        }
    }
};

__device__ __forceinline__ uint32_t lower_bound_u32(const uint32_t *__restrict__ a, uint32_t n, int64_t key) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) { uint32_t m = (lo + hi) >> 1; if ((int64_t)a[m] < key) lo = m + 1; else hi = m; }
    return lo;
}

