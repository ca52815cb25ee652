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

struct CigCursor {
    const char *t; uint32_t p, e;
    int64_t r0, q0, len;   // current ref-consuming op covers ref offsets [r0, r0+len); its first read base is q0 (op 'M')
    char op;               // 'M' (M,=,X) or 'D' (D,N); 0 before the first op
    __device__ void init(const char *text, uint32_t off, uint32_t n) { t = text; p = off; e = off + n; r0 = 0; q0 = 0; len = 0; op = 0; }
    // position on the op covering ref offset x (x must be non-decreasing across calls); false past the end
    __device__ bool seek(int64_t x) {
        while (true) {
            if (x < r0 + len) return true;
            // leave the current op
            r0 += len; if (op == 'M') q0 += len;
            len = 0; op = 0;
            // next op
            bool found = false;
            while (p < e) {
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

