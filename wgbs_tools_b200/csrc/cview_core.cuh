// cview_core.cuh -- what `cview` does to ONE pat record (reference src/cview/cview.cpp:87-167 proc_line, :8-17 pass_read,
// pipeline_wgbs/patter_utils.cpp:260-280 strip_read/strip_pat), written once for the device kernels (view.cu) and for the
// host-compiled logic check in tests/ (tests/cview_core_check.cpp builds this header with g++ and compares it with the
// reference executable; nothing in the shipped library runs it on the CPU).
//
// The reference streams sorted records past a cursor `cur_block_ind` into the blocks (sorted by startCpG).  For records
// in non-decreasing start order that cursor is a pure function of the record's start s:
//     cur(s) = first block i whose running maximum of ends max(end_0..end_i) exceeds s
// (every block before it ended at or before some earlier-or-equal start, every block from it on is still to be visited),
// and the whole run stops at the first record with s >= end of the LAST block (cview.cpp:104-110).  Both rules are
// evaluated per record here, so records can be processed independently.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define CV_HD __device__ __forceinline__
#else
#define CV_HD inline
#endif

struct CviewParams {
    const int32_t *bs, *be;   // blocks [startCpG, endCpG), sorted by start
    const int32_t *pmax;      // running maximum of be
    int32_t nb;
    const int32_t *pre_lo, *pre_hi;   // `tabix` pre-selection: closed ranges of start indices, sorted, disjoint; npre == 0: everything
    int32_t npre;
    int strict, strip, no_gaps, min_cpgs;
};

// 2-bit symbol k of a packed record (16 per word, first symbol in bits 31:30); 0 == '.'
CV_HD uint32_t cv_sym(const uint32_t *wp, uint32_t k) { return (wp[k >> 4] >> (30 - 2 * (k & 15))) & 3u; }

// pass_read (cview.cpp:8-17) for the piece [a, a+len) of the record starting at CpG s: strip, min_cpgs, no_gaps; then emit
template <class Emit>
CV_HD void cv_pass(const CviewParams &p, const uint32_t *wp, int32_t s, uint32_t a, uint32_t len, Emit &emit) {
    if (p.strip) {                                                   // strip_pat: trailing dots, then leading dots
        while (len && cv_sym(wp, a + len - 1) == 0) len--;
        if (!len) return;                                            // all dots: dropped
        while (cv_sym(wp, a) == 0) { a++; len--; }
    }
    if (p.min_cpgs < 0 || (uint32_t)p.min_cpgs > len) return;       // `min_cpgs > length()` is a size_t comparison: a negative value drops everything
    if (p.no_gaps) {
        for (uint32_t k = 0; k < len; k++) if (cv_sym(wp, a + k) == 0) return;
    }
    emit(s + (int32_t)a, a, len);
}

// One record (start s, L symbols).  emit(new_start, first_symbol, n_symbols) is called once per output line, in the
// reference's output order.
template <class Emit>
CV_HD void cview_record(const CviewParams &p, int32_t s, uint32_t L, const uint32_t *wp, Emit &emit) {
    if (p.npre) {                                                    // tabix region(s): start index inside one of the closed ranges
        int lo = 0, hi = p.npre;
        while (lo < hi) { int m = (lo + hi) >> 1; if (p.pre_hi[m] < s) lo = m + 1; else hi = m; }
        if (lo >= p.npre || p.pre_lo[lo] > s) return;
    }
    if (p.nb <= 0) return;
    const int32_t e = s + (int32_t)L - 1;                            // read_end (inclusive)
    if (s >= p.be[p.nb - 1]) return;                                 // past the last block: the run is over (cview.cpp:104-110)
    int lo = 0, hi = p.nb;                                           // cur = first block with running max end > s
    while (lo < hi) { int m = (lo + hi) >> 1; if (p.pmax[m] <= s) lo = m + 1; else hi = m; }
    const int cur = lo;
    if (cur >= p.nb) return;
    if (e < p.bs[cur]) return;                                       // ends before the current block starts
    if (!p.strict) { cv_pass(p, wp, s, 0, L, emit); return; }
    // --strict: one output per block the read overlaps, clipped to the block (cview.cpp:125-165; blocks are disjoint here)
    int32_t pos = s < p.bs[cur] ? p.bs[cur] : s;
    for (int t = cur; pos <= e && t < p.nb;) {
        if (pos >= p.bs[t] && pos < p.be[t]) {
            const int32_t room = p.be[t] - pos, left = e + 1 - pos;
            const int32_t head = room < left ? room : left;
            cv_pass(p, wp, s, (uint32_t)(pos - s), (uint32_t)head, emit);
            pos += head; t++;
        } else if (e < p.bs[t]) {
            break;
        } else {
            pos = p.bs[t];
        }
    }
}
