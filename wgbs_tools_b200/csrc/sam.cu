// sam.cu -- SAM text (the stdin of the reference's `patter`, pipeline_wgbs/patter.cpp:381-416) -> ReadBatch.
//
// Replaces line2tokens (pipeline_wgbs/patter_utils.cpp:9-18) and the per-field std::stoi calls; also produces the QNAME
// hash that template pairing keys on.  Two kernels (see below): newline offsets in one coalesced pass, then one thread per line.
#include <stdlib.h>
#include <algorithm>

#include "lines.cuh"
#include "reads.cuh"

namespace {

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

// std::stoul semantics on text[s,e): optional blanks, sign (negation wraps modulo 2^64), >=1 digit, throws past ULONG_MAX
__device__ __forceinline__ bool parse_u64(const char *__restrict__ t, uint32_t s, uint32_t e, uint64_t *out) {
    while (s < e && (t[s] == ' ' || (t[s] >= 9 && t[s] <= 13))) s++;
    bool neg = false;
    if (s < e && (t[s] == '+' || t[s] == '-')) { neg = t[s] == '-'; s++; }
    if (s >= e || t[s] < '0' || t[s] > '9') return false;
    uint64_t v = 0;
    while (s < e && t[s] >= '0' && t[s] <= '9') {
        const uint64_t d = (uint64_t)(t[s] - '0');
        if (v > (0xffffffffffffffffull - d) / 10) return false;     // out_of_range
        v = v * 10 + d; s++;
    }
    *out = neg ? (0ull - v) : v;
    return true;
}

// does the field starting at p carry an MM / ML tag?  1 = "MM:Z:" | "Mm:Z:", 2 = "ML:B:C" | "Ml:B:C"
__device__ __forceinline__ int tag_kind(const char *__restrict__ t, uint32_t p, uint32_t e) {
    if (p + 5 > e || t[p] != 'M' || t[p + 2] != ':') return 0;
    const char c1 = t[p + 1];
    if ((c1 == 'M' || c1 == 'm') && t[p + 3] == 'Z' && t[p + 4] == ':') return 1;
    if ((c1 == 'L' || c1 == 'l') && p + 6 <= e && t[p + 3] == 'B' && t[p + 4] == ':' && t[p + 5] == 'C') return 2;
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// Tokenizer = two kernels.
//   nl_scan_k      newline positions in ONE pass over the text.  64 KiB tiles handed out by an atomic ticket; a warp owns
//                  8 contiguous KiB (16 coalesced LDG.128 rounds, masks re-distributed through shared memory so that a
//                  lane owns 16 consecutive chunks); tile offsets come from a warp-wide decoupled look-back.
//   sam_lines_k    one thread per line: walks the line in aligned 16-byte chunks, finds the first ten tabs, parses FLAG and
//                  POS, hashes the QNAME, and (MM/ML mode only) keeps walking to find the MM:Z: / ML:B:C tag fields.  In
//                  bisulfite mode the walk stops at the end of SEQ: QUAL and the tags are never read.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int NLS_T = 256, NLS_ROUNDS = 16, NLS_TILE = NLS_T * NLS_ROUNDS * 16;   // 64 KiB per CTA
constexpr unsigned long long NS_AGG = 1ull << 62, NS_INC = 2ull << 62, NS_VAL = (1ull << 62) - 1;

// A tile that lies wholly inside an aligned text takes a branch-free path in which NLS_BATCH loads are issued back to back before
// the first mask is computed -- NLS_BATCH x 16 bytes in flight per thread.  Measured on the 1M-read batch (357 MB): 0.104 ms = 0.53 of the
// HBM roofline with 8 loads in flight, 0.106 ms with 16, 0.129 ms with the guarded one-load-at-a-time loop (the SASS of that one shows
// the 16 LDG.128 of a thread ~170 instructions apart, each behind the branches of its guard), 0.147 ms with the tile fetched by one
// TMA bulk copy into shared memory; a warp-per-tile variant without block barriers lost as well.  Only this form is kept.
constexpr int NLS_BATCH = 8;
__global__ void __launch_bounds__(NLS_T) nl_scan_k(const char *__restrict__ text, size_t n, uint32_t cap,
                                                    unsigned long long *__restrict__ status, unsigned int *__restrict__ ticket,
                                                    uint32_t *__restrict__ nlpos, uint32_t *__restrict__ total) {
    __shared__ uint16_t sm_mask[NLS_T * NLS_ROUNDS];   // one 16-bit newline mask per chunk; warp w owns [w*512, w*512+512)
    __shared__ uint32_t ws[NLS_T / 32];
    __shared__ uint64_t s_base;
    __shared__ unsigned s_tile;
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const unsigned tile = s_tile;
    const size_t warp0 = (size_t)tile * NLS_TILE + (size_t)w * (32 * NLS_ROUNDS * 16);
    const bool whole = (size_t)(tile + 1) * NLS_TILE <= n && (((uintptr_t)text) & 15) == 0;     // uniform over the CTA
    if (whole) {
        const uint4 *src = reinterpret_cast<const uint4 *>(text + warp0) + lane;
#pragma unroll
        for (int h = 0; h < NLS_ROUNDS; h += NLS_BATCH) {
            uint4 v[NLS_BATCH];
#pragma unroll
            for (int c = 0; c < NLS_BATCH; c++)      // volatile asm: ptxas keeps these loads together, ahead of the first use
                asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v[c].x), "=r"(v[c].y), "=r"(v[c].z), "=r"(v[c].w) : "l"(src + (h + c) * 32));
#pragma unroll
            for (int c = 0; c < NLS_BATCH; c++) sm_mask[w * 512 + (h + c) * 32 + lane] = (uint16_t)eq_mask16(v[c], '\n');
        }
    } else
#pragma unroll
    for (int c = 0; c < NLS_ROUNDS; c++) {
        const size_t p = warp0 + (size_t)(c * 32 + lane) * 16;
        uint32_t nm = 0;
        if (p < n) {
            const uint4 v = (p + 16 <= n) ? *reinterpret_cast<const uint4 *>(text + p) : load16_guard(text, p, n);
            nm = eq_mask16(v, '\n');
            if (n - p < 16) nm &= (1u << (n - p)) - 1;
        }
        sm_mask[w * 512 + c * 32 + lane] = (uint16_t)nm;
    }
    __syncwarp();
    // this lane's 16 consecutive chunks (256 bytes): 32 contiguous bytes of shared memory
    const uint4 q0 = *reinterpret_cast<const uint4 *>(&sm_mask[w * 512 + lane * 16]);
    const uint4 q1 = *reinterpret_cast<const uint4 *>(&sm_mask[w * 512 + lane * 16 + 8]);
    const uint32_t mk[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};   // two 16-bit masks per word
    uint32_t cnt = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) cnt += __popc(mk[i]);
    uint32_t inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= (unsigned)d) inc += t; }
    if (lane == 31) ws[w] = inc;
    __syncthreads();
    uint32_t wpre = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < NLS_T / 32; i++) { const uint32_t x = ws[i]; if (i < (int)w) wpre += x; tot += x; }
    if (w == 0) {
        volatile unsigned long long *st = status;
        uint64_t base = 0;
        if (tile == 0) { if (lane == 0) st[0] = NS_INC | tot; }
        else {
            if (lane == 0) st[tile] = NS_AGG | tot;
            long long top = (long long)tile - 1;
            while (true) {
                const long long j = top - lane;
                unsigned long long x = NS_INC;
                if (j >= 0) x = st[j];
                while (__any_sync(0xffffffffu, (x >> 62) == 0)) { if ((x >> 62) == 0) x = st[j]; }
                const unsigned incm = __ballot_sync(0xffffffffu, (x >> 62) == 2);
                uint64_t val = x & NS_VAL;
                if (incm) { const int L = __ffs(incm) - 1; if ((int)lane > L) val = 0; }
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) val += __shfl_xor_sync(0xffffffffu, val, d);
                base += val;
                if (incm) break;
                top -= 32;
            }
            if (lane == 0) st[tile] = NS_INC | ((base + tot) & NS_VAL);
        }
        if (lane == 0) { s_base = base; if ((size_t)(tile + 1) * NLS_TILE >= n) *total = (uint32_t)(base + tot); }
    }
    __syncthreads();
    uint32_t o = (uint32_t)s_base + wpre + inc - cnt;
    const size_t p0 = warp0 + (size_t)lane * 256;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint32_t m = mk[i];
        while (m) { const int b = __ffs(m) - 1; m &= m - 1; if (o < cap) nlpos[o] = (uint32_t)(p0 + (size_t)i * 32 + b); o++; }
    }
}


// PF: 16-byte chunks fetched together (independent loads in flight per thread) before they are examined one by one
template <int PF>
__global__ void __launch_bounds__(128) sam_lines_k(const char *__restrict__ text, uint32_t n, const uint32_t *__restrict__ nlpos, uint32_t n_nl,
                                                    uint32_t n_lines, int want_tags, ReadBatch rb) {
    const uint32_t line = blockIdx.x * blockDim.x + threadIdx.x;
    if (line >= n_lines) return;
    const uint32_t s = line == 0 ? 0 : nlpos[line - 1] + 1;
    const uint32_t e = line < n_nl ? nlpos[line] : n;
    uint32_t tb[10];
    uint32_t ntab = 0, mm_off = 0, ml_off = 0;
    uint4 buf[PF];
    for (uint32_t base = s & ~15u, k = PF; base < e; base += 16, k++) {
        if (PF > 1) {
            if (k >= PF) {                                           // refill: PF loads issued back to back
#pragma unroll
                for (int j = 0; j < PF; j++) {
                    const uint32_t b2 = base + 16 * j;
                    buf[j] = make_uint4(0, 0, 0, 0);
                    if (b2 < e) buf[j] = (b2 + 16 <= n) ? *reinterpret_cast<const uint4 *>(text + b2) : load16_guard(text, b2, n);
                }
                k = 0;
            }
        }
        uint4 v;
        if (PF > 1) {
            v = buf[0];
#pragma unroll
            for (int j = 1; j < PF; j++) if (k == j) v = buf[j];
        } else {
            v = (base + 16 <= n) ? *reinterpret_cast<const uint4 *>(text + base) : load16_guard(text, base, n);
        }
        uint32_t m = eq_mask16(v, '\t');
        if (base < s) m &= 0xffffu << (s - base);
        if (e - base < 16) m &= (1u << (e - base)) - 1;
        while (m) {
            const int b = __ffs(m) - 1; m &= m - 1;
            const uint32_t x = base + b;
            if (ntab < 10) tb[ntab] = x;
            else if (want_tags) {                                   // ntab >= 10: field index ntab+1 >= 11 starts at x+1
                const int k = tag_kind(text, x + 1, e);
                if (k == 1) mm_off = x + 6; else if (k == 2) ml_off = x + 7;                 // last occurrence wins
            }
            ntab++;
        }
        if (ntab >= 10 && !want_tags) break;                        // bisulfite mode: nothing beyond SEQ is needed
    }
    const uint32_t qend = ntab > 0 ? tb[0] : e;
    uint64_t h = 0;
    for (uint32_t p = s; p < qend; p++) h += name_byte_mix((uint8_t)text[p], p - s);
    h = fmix64(h ^ (uint64_t)(qend - s));
    uint8_t st = REC_OK;
    int32_t flag = 0; uint64_t pos64 = 0;
    uint32_t cig_off = s, cig_len = 0, seq_off = s, seq_len = 0;
    if (e == s) { st = REC_BLANK; h = fmix64(0x5851f42d4c957f2dULL ^ (uint64_t)line); }   // blank lines: unique keys, never paired
    else if (ntab < 10 || tb[9] == e - 1) {
        // line2tokens yields < 11 tokens: fewer than 10 tabs, or exactly 10 with nothing after the last one.  (With > 10 tabs
        // tb[9] == e-1 is impossible.)
        st = REC_INVALID;
    } else {
        // FLAG: std::stoi.  POS: std::stoul (patter.cpp:208), i.e. 64-bit with wrap-around; the bisulfite path then narrows it to
        // `int` (compareSeqToRef's parameter), the MM/ML path keeps all 64 bits (ont.cpp:102).
        if (!parse_i32(text, tb[0] + 1, tb[1], &flag) || !parse_u64(text, tb[2] + 1, tb[3], &pos64)) st = REC_BADINT;
        cig_off = tb[4] + 1; cig_len = tb[5] - tb[4] - 1;
        seq_off = tb[8] + 1; seq_len = tb[9] - tb[8] - 1;
    }
    rb.line_off[line] = s; rb.line_len[line] = e - s; rb.qn_len[line] = qend - s;
    rb.flag[line] = flag; rb.pos[line] = (int32_t)(uint32_t)pos64; rb.pos_hi[line] = (int32_t)(uint32_t)(pos64 >> 32);
    rb.cig_off[line] = cig_off; rb.cig_len[line] = cig_len; rb.seq_off[line] = seq_off; rb.seq_len[line] = seq_len;
    rb.hash_lo[line] = (uint32_t)h; rb.hash_hi[line] = (uint32_t)(h >> 32);
    rb.status[line] = st;
    if (want_tags) {
        uint32_t z = mm_off; if (mm_off) while (z < e && text[z] != '\t') z++;
        rb.mm_off[line] = mm_off; rb.mm_len[line] = mm_off ? z - mm_off : 0;
        z = ml_off; if (ml_off) while (z < e && text[z] != '\t') z++;
        rb.ml_off[line] = ml_off; rb.ml_len[line] = ml_off ? z - ml_off : 0;
    }
}


}  // namespace

int sam_tokenize(wgbs_ctx *ctx, const char *dtext, size_t nbytes, int want_tags /* 1 yes, 0 no, -1 decide from the first line */, Temps &T,
                 ReadBatch *out) {
    if (nbytes >= 0xfffffff0ull) return wgbs_set_err("SAM text must be < 4 GiB per call (got %zu); split on line boundaries", nbytes);
    const uint32_t ntiles = (uint32_t)((nbytes + NLS_TILE - 1) / NLS_TILE);
    ReadBatch rb;
    rb.text = dtext; rb.nbytes = (uint32_t)nbytes;
    uint32_t *nlpos = nullptr, *totals = ctx->d_flags + 8;
    unsigned long long *status = nullptr;
    if (ntiles) RC_TRY(T.alloc(&status, (size_t)ntiles + 1));
    uint32_t cap = (uint32_t)(nbytes / 64 + 1024);            // optimistic: average line >= 64 bytes (a 50 bp SAM record is ~130)
    uint32_t n_nl = 0, first_mm = 0; char last = '\n';
    // first_line (patter.cpp:337-338): does the first non-empty line carry an MM tag?  Decided on the host from the head of the
    // text (one small D2H that rides on the synchronisation the line count needs anyway).
    std::vector<char> head(want_tags < 0 ? std::min<size_t>(nbytes, 1u << 16) : 0);
    for (int attempt = 0; attempt < 2 && ntiles; attempt++) {
        if (nlpos) { T.keep(nlpos); dfree(ctx, nlpos); }
        RC_TRY(T.alloc(&nlpos, cap));
        CUDA_TRY(cudaMemsetAsync(status, 0, ((size_t)ntiles + 1) * 8, ctx->stream));
        LAUNCH(ctx, nl_scan_k, ntiles, NLS_T, 0, dtext, nbytes, cap, status, (unsigned int *)(status + ntiles), nlpos, totals);
        CUDA_TRY(cudaMemcpyAsync(&n_nl, totals, 4, cudaMemcpyDeviceToHost, ctx->stream));
        if (attempt == 0 && !head.empty()) CUDA_TRY(cudaMemcpyAsync(head.data(), dtext, head.size(), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(&last, dtext + nbytes - 1, 1, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        if (n_nl <= cap) break;
        cap = n_nl;                                            // many very short lines: repeat with the exact size
    }
    if (!head.empty()) {
        size_t p = 0, hn = head.size();
        while (p < hn && head[p] == '\n') p++;
        size_t le = p; while (le < hn && head[le] != '\n') le++;
        if (le == hn && hn < nbytes) {                              // first line longer than the peek (long ONT read): fetch all of it
            std::vector<uint32_t> nl0(1);
            if (n_nl > p) { CUDA_TRY(cudaMemcpyAsync(nl0.data(), nlpos + p, 4, cudaMemcpyDeviceToHost, ctx->stream)); CUDA_TRY(cudaStreamSynchronize(ctx->stream)); }
            size_t want = n_nl > p ? (size_t)nl0[0] + 1 : nbytes;
            head.resize(want);
            CUDA_TRY(cudaMemcpyAsync(head.data(), dtext, want, cudaMemcpyDeviceToHost, ctx->stream)); CUDA_TRY(cudaStreamSynchronize(ctx->stream));
            hn = head.size(); le = p; while (le < hn && head[le] != '\n') le++;
        }
        int field = 0; size_t fs = p;
        for (size_t i = p; i <= le; i++) {
            if (i == le || head[i] == '\t') {
                if (field >= 11 && i - fs > 5 && head[fs] == 'M' && (head[fs + 1] == 'M' || head[fs + 1] == 'm') && head[fs + 2] == ':' && head[fs + 3] == 'Z' && head[fs + 4] == ':') first_mm = 1;
                field++; fs = i + 1;
            }
        }
    }
    const bool tags = want_tags > 0 || (want_tags < 0 && first_mm);
    const uint32_t n_lines = n_nl + ((nbytes && last != '\n') ? 1 : 0);
    rb.n = n_lines;
    RC_TRY(T.alloc(&rb.line_off, n_lines)); RC_TRY(T.alloc(&rb.line_len, n_lines)); RC_TRY(T.alloc(&rb.qn_len, n_lines));
    RC_TRY(T.alloc(&rb.flag, n_lines)); RC_TRY(T.alloc(&rb.pos, n_lines)); RC_TRY(T.alloc(&rb.pos_hi, n_lines));
    RC_TRY(T.alloc(&rb.cig_off, n_lines)); RC_TRY(T.alloc(&rb.cig_len, n_lines));
    RC_TRY(T.alloc(&rb.seq_off, n_lines)); RC_TRY(T.alloc(&rb.seq_len, n_lines));
    RC_TRY(T.alloc(&rb.hash_lo, n_lines)); RC_TRY(T.alloc(&rb.hash_hi, n_lines));
    RC_TRY(T.alloc(&rb.status, n_lines));
    if (tags) {
        RC_TRY(T.alloc(&rb.mm_off, n_lines)); RC_TRY(T.alloc(&rb.mm_len, n_lines));
        RC_TRY(T.alloc(&rb.ml_off, n_lines)); RC_TRY(T.alloc(&rb.ml_len, n_lines));
    }
    if (n_lines) {
        LAUNCH(ctx, sam_lines_k<2>, grid_for(n_lines, 128), 128, 0, dtext, (uint32_t)nbytes, nlpos, n_nl, n_lines, tags ? 1 : 0, rb);
        LAUNCH_CHECK();
    }
    *out = rb;
    return 0;
}
