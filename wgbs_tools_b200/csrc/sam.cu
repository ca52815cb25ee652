// sam.cu -- SAM text (the stdin of the reference's `patter`, pipeline_wgbs/patter.cpp:381-416) -> ReadBatch.
//
// Replaces line2tokens (pipeline_wgbs/patter_utils.cpp:9-18) and the per-field std::stoi calls; also produces the QNAME
// hash that template pairing sorts on.  Flat data-parallel design (see below): the text is read twice with aligned
// 16-byte loads, every tab learns its ordinal inside its line from a segmented scan.
#include <stdlib.h>
#include <algorithm>

#include "lines.cuh"
#include "reads.cuh"

#ifndef WGBS_NLSCAN_DEFAULT_WARP
#define WGBS_NLSCAN_DEFAULT_WARP 0
#endif
#ifndef WGBS_LINES_PF_DEFAULT
#define WGBS_LINES_PF_DEFAULT 2
#endif
#ifndef WGBS_TOKENIZER_DEFAULT_FUSED
#define WGBS_TOKENIZER_DEFAULT_FUSED 0
#endif

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

// BATCH = 0: every 16-byte chunk is loaded, guarded and examined in turn (the SASS shows the 16 LDG.128 of a thread ~170
// instructions apart, each behind the branches of its guard: ONE load in flight per thread).  BATCH = 4 / 8 / 16 (staged,
// WGBS_NLSCAN=batch4|batch8|batch16): a tile that lies wholly inside an aligned text takes a branch-free path in which BATCH
// loads are issued back to back before the first mask is computed -- BATCH x 16 bytes in flight per thread.
template <int BATCH>
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
    const bool whole = BATCH != 0 && (size_t)(tile + 1) * NLS_TILE <= n && (((uintptr_t)text) & 15) == 0;     // uniform over the CTA
    if (BATCH < 0 && whole) {
        // BATCH = -1 (WGBS_NLSCAN=tma): the 64 KiB tile is fetched by ONE bulk asynchronous copy (TMA engine, global -> shared,
        // completion counted in bytes on an mbarrier); the threads then take their chunks from shared memory (conflict-free:
        // consecutive lanes, consecutive 16 bytes).  64 KiB in flight per CTA whatever the compiler makes of the loop.
        extern __shared__ __align__(128) unsigned char nl_tile[];
        __shared__ __align__(8) unsigned long long nl_bar;
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&nl_bar), dst = (uint32_t)__cvta_generic_to_shared(nl_tile);
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)NLS_TILE) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst), "l"(text + (size_t)tile * NLS_TILE), "r"((uint32_t)NLS_TILE), "r"(bar) : "memory");
        }
        uint32_t landed = 0;
        while (!landed)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(landed) : "r"(bar), "r"(0u) : "memory");
        const uint4 *src = reinterpret_cast<const uint4 *>(nl_tile + (size_t)w * (32 * NLS_ROUNDS * 16)) + lane;
#pragma unroll
        for (int c = 0; c < NLS_ROUNDS; c++) sm_mask[w * 512 + c * 32 + lane] = (uint16_t)eq_mask16(src[c * 32], '\n');
    } else if (whole) {
        const uint4 *src = reinterpret_cast<const uint4 *>(text + warp0) + lane;
#pragma unroll
        for (int h = 0; h < NLS_ROUNDS; h += (BATCH > 0 ? BATCH : 1)) {
            uint4 v[BATCH > 0 ? BATCH : 1];
#pragma unroll
            for (int c = 0; c < (BATCH > 0 ? BATCH : 1); c++)      // volatile asm: ptxas keeps these loads together, ahead of the first use
                asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v[c].x), "=r"(v[c].y), "=r"(v[c].z), "=r"(v[c].w) : "l"(src + (h + c) * 32));
#pragma unroll
            for (int c = 0; c < (BATCH > 0 ? BATCH : 1); c++) sm_mask[w * 512 + (h + c) * 32 + lane] = (uint16_t)eq_mask16(v[c], '\n');
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

// Variant of nl_scan_k without block-wide barriers: every WARP is its own 8 KiB tile with its own ticket and its own
// look-back, so no warp ever waits for the look-back of another one (the barrier stall was the top stall reason of the
// CTA-wide version).  Selected with WGBS_NLSCAN=warp|cta.
constexpr int NLW_T = 128, NLW_TILE = 32 * NLS_ROUNDS * 16;            // 8 KiB per warp
__global__ void __launch_bounds__(NLW_T) nl_scan_warp_k(const char *__restrict__ text, size_t n, uint32_t cap,
                                                         unsigned long long *__restrict__ status, unsigned int *__restrict__ ticket,
                                                         uint32_t *__restrict__ nlpos, uint32_t *__restrict__ total) {
    __shared__ __align__(16) uint16_t sm_mask[(NLW_T / 32) * 32 * NLS_ROUNDS];
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned tile = 0;
    if (lane == 0) tile = atomicAdd(ticket, 1u);
    tile = __shfl_sync(0xffffffffu, tile, 0);
    const size_t warp0 = (size_t)tile * NLW_TILE;
    if (warp0 >= n) return;                                          // whole warp: grid is rounded up to full CTAs
    uint16_t *mym = sm_mask + w * (32 * NLS_ROUNDS);
#pragma unroll
    for (int h = 0; h < NLS_ROUNDS; h += 8) {                        // 8 independent 16-byte loads in flight per lane
        uint4 v[8];
#pragma unroll
        for (int c = 0; c < 8; c++) {
            const size_t p = warp0 + (size_t)((h + c) * 32 + lane) * 16;
            v[c] = make_uint4(0, 0, 0, 0);
            if (p < n) v[c] = (p + 16 <= n) ? *reinterpret_cast<const uint4 *>(text + p) : load16_guard(text, p, n);
        }
#pragma unroll
        for (int c = 0; c < 8; c++) {
            const size_t p = warp0 + (size_t)((h + c) * 32 + lane) * 16;
            uint32_t nm = 0;
            if (p < n) { nm = eq_mask16(v[c], '\n'); if (n - p < 16) nm &= (1u << (n - p)) - 1; }
            mym[(h + c) * 32 + lane] = (uint16_t)nm;
        }
    }
    __syncwarp();
    const uint4 q0 = *reinterpret_cast<const uint4 *>(&mym[lane * 16]);
    const uint4 q1 = *reinterpret_cast<const uint4 *>(&mym[lane * 16 + 8]);
    const uint32_t mk[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
    uint32_t cnt = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) cnt += __popc(mk[i]);
    uint32_t inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= (unsigned)d) inc += t; }
    const uint32_t tot = __shfl_sync(0xffffffffu, inc, 31);
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
    if (lane == 0 && warp0 + NLW_TILE >= n) *total = (uint32_t)(base + tot);
    uint32_t o = (uint32_t)base + inc - cnt;
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


// ---------------------------------------------------------------------------------------------------------------------
// Fused tokenizer (ONE pass over the text): sam_scan_k = nl_scan_k + sam_lines_k on a 32 KiB tile that stays in shared memory.
//   phase A  as nl_scan_k (coalesced LDG.128, all of a thread's loads in flight before the first use), and the bytes are
//            parked in shared memory; newline positions go to nlpos[] at their global rank (decoupled look-back).
//   phase B  the tile that holds a newline owns the line that STARTS behind it (tile 0 also owns line 0).  One thread per
//            owned line walks it in 16-byte chunks -- from shared memory while inside the tile, from global memory beyond it
//            (the tail of the tile's last line; long ONT reads) -- finds the tabs AND the line's end, and writes the record.
// The text is read from HBM once; what sam_lines_k fetched again (and only partly used, sector by sector) now comes out of
// shared memory.  Selected with WGBS_TOKENIZER=fused|split (see sam_tokenize).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int FT_T = 256, FT_ROUNDS = 8, FT_TILE = FT_T * FT_ROUNDS * 16;   // 32 KiB per CTA

struct TileAcc {
    const char *g; const unsigned char *sm; uint32_t t0, t1, n;      // bytes [t0, t1) of the text are resident in sm
    __device__ __forceinline__ uint32_t byte(uint32_t p) const { return (p >= t0 && p < t1) ? sm[p - t0] : (uint32_t)(uint8_t)g[p]; }
    __device__ __forceinline__ uint4 chunk(uint32_t base) const {    // base % 16 == 0, base < n
        if (base >= t0 && base < t1) return *reinterpret_cast<const uint4 *>(sm + (base - t0));
        return (base + 16 <= n) ? *reinterpret_cast<const uint4 *>(g + base) : load16_guard(g, base, n);
    }
};
__device__ __forceinline__ bool is_blank(uint32_t c) { return c == ' ' || (c >= 9 && c <= 13); }
// parse_i32 / parse_u64 / tag_kind on an accessor (same semantics as the pointer versions above)
__device__ __forceinline__ bool parse_i32_acc(const TileAcc &t, uint32_t s, uint32_t e, int32_t *out) {
    while (s < e && is_blank(t.byte(s))) s++;
    bool neg = false;
    if (s < e) { const uint32_t c = t.byte(s); if (c == '+' || c == '-') { neg = c == '-'; s++; } }
    if (s >= e) return false;
    uint32_t c = t.byte(s);
    if (c < '0' || c > '9') return false;
    int64_t v = 0;
    while (true) {
        v = v * 10 + (int64_t)(c - '0'); if (v > 0x80000000LL) return false;
        if (++s >= e) break;
        c = t.byte(s); if (c < '0' || c > '9') break;
    }
    if (neg) v = -v;
    if (v > 0x7fffffffLL || v < -0x80000000LL) return false;
    *out = (int32_t)v;
    return true;
}
__device__ __forceinline__ bool parse_u64_acc(const TileAcc &t, uint32_t s, uint32_t e, uint64_t *out) {
    while (s < e && is_blank(t.byte(s))) s++;
    bool neg = false;
    if (s < e) { const uint32_t c = t.byte(s); if (c == '+' || c == '-') { neg = c == '-'; s++; } }
    if (s >= e) return false;
    uint32_t c = t.byte(s);
    if (c < '0' || c > '9') return false;
    uint64_t v = 0;
    while (true) {
        const uint64_t d = (uint64_t)(c - '0');
        if (v > (0xffffffffffffffffull - d) / 10) return false;       // out_of_range
        v = v * 10 + d;
        if (++s >= e) break;
        c = t.byte(s); if (c < '0' || c > '9') break;
    }
    *out = neg ? (0ull - v) : v;
    return true;
}
__device__ __forceinline__ int tag_kind_acc(const TileAcc &t, uint32_t p, uint32_t e) {
    if (p + 5 > e || t.byte(p) != 'M' || t.byte(p + 2) != ':') return 0;
    const uint32_t c1 = t.byte(p + 1);
    if ((c1 == 'M' || c1 == 'm') && t.byte(p + 3) == 'Z' && t.byte(p + 4) == ':') return 1;
    if ((c1 == 'L' || c1 == 'l') && p + 6 <= e && t.byte(p + 3) == 'B' && t.byte(p + 4) == ':' && t.byte(p + 5) == 'C') return 2;
    return 0;
}

__global__ void __launch_bounds__(FT_T) sam_scan_k(const char *__restrict__ text, uint32_t n, uint32_t cap_nl, uint32_t cap_lines,
                                                    unsigned long long *__restrict__ status, unsigned int *__restrict__ ticket,
                                                    uint32_t *nlpos, uint32_t *__restrict__ total, int want_tags, ReadBatch rb) {
    __shared__ __align__(16) unsigned char sm_text[FT_TILE];
    __shared__ __align__(16) uint16_t sm_mask[FT_T * FT_ROUNDS];     // one 16-bit newline mask per chunk; warp w owns [w*256, w*256+256)
    __shared__ uint32_t ws[FT_T / 32];
    __shared__ uint64_t s_base;
    __shared__ unsigned s_tile;
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const unsigned tile = s_tile;
    const size_t t0 = (size_t)tile * FT_TILE;
    const size_t warp0 = t0 + (size_t)w * (32 * FT_ROUNDS * 16);
    // ---- phase A: load, park, newline masks ---------------------------------------------------------------------------
    uint4 v[FT_ROUNDS];
#pragma unroll
    for (int c = 0; c < FT_ROUNDS; c++) {
        const size_t p = warp0 + (size_t)(c * 32 + lane) * 16;
        v[c] = make_uint4(0, 0, 0, 0);
        if (p < n) v[c] = (p + 16 <= n) ? *reinterpret_cast<const uint4 *>(text + p) : load16_guard(text, p, n);
    }
#pragma unroll
    for (int c = 0; c < FT_ROUNDS; c++) {
        const size_t p = warp0 + (size_t)(c * 32 + lane) * 16;
        *reinterpret_cast<uint4 *>(sm_text + (p - t0)) = v[c];
        uint32_t nm = 0;
        if (p < n) { nm = eq_mask16(v[c], '\n'); if (n - p < 16) nm &= (1u << (n - p)) - 1; }
        sm_mask[w * (32 * FT_ROUNDS) + c * 32 + lane] = (uint16_t)nm;
    }
    __syncwarp();
    // this lane's 8 consecutive chunks (128 bytes): 16 contiguous bytes of shared memory, two 16-bit masks per word
    const uint4 q0 = *reinterpret_cast<const uint4 *>(&sm_mask[w * (32 * FT_ROUNDS) + lane * FT_ROUNDS]);
    const uint32_t mk[4] = {q0.x, q0.y, q0.z, q0.w};
    uint32_t cnt = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) cnt += __popc(mk[i]);
    uint32_t inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= (unsigned)d) inc += t; }
    if (lane == 31) ws[w] = inc;
    __syncthreads();
    uint32_t wpre = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < FT_T / 32; i++) { const uint32_t x = ws[i]; if (i < (int)w) wpre += x; tot += x; }
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
        if (lane == 0) { s_base = base; if (t0 + FT_TILE >= (size_t)n) *total = (uint32_t)(base + tot); }
    }
    __syncthreads();
    const uint32_t gbase = (uint32_t)s_base;
    {
        uint32_t o = gbase + wpre + inc - cnt;
        const size_t p0 = warp0 + (size_t)lane * (16 * FT_ROUNDS);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            uint32_t m = mk[i];
            while (m) { const int b = __ffs(m) - 1; m &= m - 1; if (o < cap_nl) nlpos[o] = (uint32_t)(p0 + (size_t)i * 32 + b); o++; }
        }
    }
    __syncthreads();                                                 // this CTA's nlpos[] entries and sm_text are complete
    // ---- phase B: one thread per line that starts behind a newline of this tile -----------------------------------------
    const uint32_t t1 = (t0 + FT_TILE < (size_t)n) ? (uint32_t)(t0 + FT_TILE) : n;
    const TileAcc acc{text, sm_text, (uint32_t)t0, t1, n};
    const uint32_t first = tile == 0 ? 1u : 0u;
    const uint32_t L = tot + first;
    for (uint32_t q = threadIdx.x; q < L; q += FT_T) {
        uint32_t line = 0, s = 0;
        if (q >= first) {
            const uint32_t g = gbase + (q - first);
            if (g >= cap_nl) continue;                               // the host repeats the call with larger arrays
            s = __ldcg(&nlpos[g]) + 1; line = g + 1;
        }
        if (s >= n || line >= cap_lines) continue;                   // nothing behind the final newline
        uint32_t tb[10];
        uint32_t ntab = 0, mm_off = 0, ml_off = 0, e = n;
        for (uint32_t base = s & ~15u; base < n; base += 16) {
            const uint4 c = acc.chunk(base);
            uint32_t mt = eq_mask16(c, '\t'), mn = eq_mask16(c, '\n');
            if (base < s) { const uint32_t k = 0xffffu << (s - base); mt &= k; mn &= k; }
            if (n - base < 16) { const uint32_t k = (1u << (n - base)) - 1; mt &= k; mn &= k; }
            if (mn) { const int b = __ffs(mn) - 1; e = base + b; mt &= (1u << b) - 1; }
            while (mt) {
                const int b = __ffs(mt) - 1; mt &= mt - 1;
                const uint32_t x = base + b;
                if (ntab < 10) tb[ntab] = x;
                else if (want_tags) {                               // field index ntab+1 >= 11 starts at x+1 (bytes past the line's end cannot match)
                    const int k = tag_kind_acc(acc, x + 1, e);
                    if (k == 1) mm_off = x + 6; else if (k == 2) ml_off = x + 7;                 // last occurrence wins
                }
                ntab++;
            }
            if (mn) break;
            if (ntab >= 10 && !want_tags) {                          // bisulfite mode: only the line's end is still needed
                for (base += 16; base < n; base += 16) {
                    uint32_t m2 = eq_mask16(acc.chunk(base), '\n');
                    if (n - base < 16) m2 &= (1u << (n - base)) - 1;
                    if (m2) { e = base + __ffs(m2) - 1; break; }
                }
                break;
            }
        }
        const uint32_t qend = ntab > 0 ? tb[0] : e;
        uint64_t h = 0;
        for (uint32_t p = s; p < qend; p++) h += name_byte_mix(acc.byte(p), p - s);
        h = fmix64(h ^ (uint64_t)(qend - s));
        uint8_t st = REC_OK;
        int32_t flag = 0; uint64_t pos64 = 0;
        uint32_t cig_off = s, cig_len = 0, seq_off = s, seq_len = 0;
        if (e == s) { st = REC_BLANK; h = fmix64(0x5851f42d4c957f2dULL ^ (uint64_t)line); }
        else if (ntab < 10 || tb[9] == e - 1) st = REC_INVALID;
        else {
            if (!parse_i32_acc(acc, tb[0] + 1, tb[1], &flag) || !parse_u64_acc(acc, tb[2] + 1, tb[3], &pos64)) st = REC_BADINT;
            cig_off = tb[4] + 1; cig_len = tb[5] - tb[4] - 1;
            seq_off = tb[8] + 1; seq_len = tb[9] - tb[8] - 1;
        }
        rb.line_off[line] = s; rb.line_len[line] = e - s; rb.qn_len[line] = qend - s;
        rb.flag[line] = flag; rb.pos[line] = (int32_t)(uint32_t)pos64; rb.pos_hi[line] = (int32_t)(uint32_t)(pos64 >> 32);
        rb.cig_off[line] = cig_off; rb.cig_len[line] = cig_len; rb.seq_off[line] = seq_off; rb.seq_len[line] = seq_len;
        rb.hash_lo[line] = (uint32_t)h; rb.hash_hi[line] = (uint32_t)(h >> 32);
        rb.status[line] = st;
        if (want_tags) {
            uint32_t z = mm_off; if (mm_off) while (z < e && acc.byte(z) != '\t') z++;
            rb.mm_off[line] = mm_off; rb.mm_len[line] = mm_off ? z - mm_off : 0;
            z = ml_off; if (ml_off) while (z < e && acc.byte(z) != '\t') z++;
            rb.ml_off[line] = ml_off; rb.ml_len[line] = ml_off ? z - ml_off : 0;
        }
    }
}

}  // namespace

// first_line (patter.cpp:337-338): does the first non-empty line of `head` (the first bytes of the text) carry an MM tag?
// returns -1 when the line does not end inside `head` (and head is not the whole text)
static int first_line_has_mm(const std::vector<char> &head, size_t nbytes) {
    size_t p = 0; const size_t hn = head.size();
    while (p < hn && head[p] == '\n') p++;
    size_t le = p; while (le < hn && head[le] != '\n') le++;
    if (le == hn && hn < nbytes) return -1;
    int field = 0, mm = 0; size_t fs = p;
    for (size_t i = p; i <= le; i++) {
        if (i == le || head[i] == '\t') {
            if (field >= 11 && i - fs > 5 && head[fs] == 'M' && (head[fs + 1] == 'M' || head[fs + 1] == 'm') && head[fs + 2] == ':' && head[fs + 3] == 'Z' && head[fs + 4] == ':') mm = 1;
            field++; fs = i + 1;
        }
    }
    return mm;
}

static int alloc_records(Temps &T, ReadBatch &rb, size_t cap, bool tags) {
    RC_TRY(T.alloc(&rb.line_off, cap)); RC_TRY(T.alloc(&rb.line_len, cap)); RC_TRY(T.alloc(&rb.qn_len, cap));
    RC_TRY(T.alloc(&rb.flag, cap)); RC_TRY(T.alloc(&rb.pos, cap)); RC_TRY(T.alloc(&rb.pos_hi, cap));
    RC_TRY(T.alloc(&rb.cig_off, cap)); RC_TRY(T.alloc(&rb.cig_len, cap));
    RC_TRY(T.alloc(&rb.seq_off, cap)); RC_TRY(T.alloc(&rb.seq_len, cap));
    RC_TRY(T.alloc(&rb.hash_lo, cap)); RC_TRY(T.alloc(&rb.hash_hi, cap));
    RC_TRY(T.alloc(&rb.status, cap));
    if (tags) {
        RC_TRY(T.alloc(&rb.mm_off, cap)); RC_TRY(T.alloc(&rb.mm_len, cap));
        RC_TRY(T.alloc(&rb.ml_off, cap)); RC_TRY(T.alloc(&rb.ml_len, cap));
    }
    return 0;
}
static void free_records(wgbs_ctx *ctx, Temps &T, ReadBatch &rb) {
    void *ps[] = {rb.line_off, rb.line_len, rb.qn_len, rb.flag, rb.pos, rb.pos_hi, rb.cig_off, rb.cig_len, rb.seq_off, rb.seq_len,
                  rb.hash_lo, rb.hash_hi, rb.status, rb.mm_off, rb.mm_len, rb.ml_off, rb.ml_len};
    for (void *q : ps) if (q) { T.keep(q); dfree(ctx, q); }
    const char *t = rb.text; const uint32_t nb = rb.nbytes;
    rb = ReadBatch(); rb.text = t; rb.nbytes = nb;
}

// one pass: sam_scan_k.  The record arrays are sized from an estimate of the line count before the kernel runs (average
// line >= 64 bytes); a text with more lines than that is tokenized again with exact sizes.
static int sam_tokenize_fused(wgbs_ctx *ctx, const char *dtext, size_t nbytes, int want_tags, Temps &T, ReadBatch *out) {
    const uint32_t ntiles = (uint32_t)((nbytes + FT_TILE - 1) / FT_TILE);
    ReadBatch rb;
    rb.text = dtext; rb.nbytes = (uint32_t)nbytes;
    if (!ntiles) { rb.n = 0; RC_TRY(alloc_records(T, rb, 1, want_tags > 0)); *out = rb; return 0; }
    bool tags = want_tags > 0;
    if (want_tags < 0) {                                            // decided on the host from the head of the text, before the launch
        size_t want = std::min<size_t>(nbytes, 1u << 16);
        while (true) {
            std::vector<char> head(want);
            CUDA_TRY(cudaMemcpyAsync(head.data(), dtext, want, cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_TRY(cudaStreamSynchronize(ctx->stream));
            const int r = first_line_has_mm(head, nbytes);
            if (r >= 0) { tags = r == 1; break; }
            want = std::min<size_t>(nbytes, want * 8);              // first line longer than the peek (long ONT read)
        }
    }
    uint32_t *nlpos = nullptr, *totals = ctx->d_flags + 8;
    unsigned long long *status = nullptr;
    RC_TRY(T.alloc(&status, (size_t)ntiles + 1));
    uint32_t cap = (uint32_t)(nbytes / 64 + 1024), n_nl = 0; char last = '\n';
    for (int attempt = 0; attempt < 2; attempt++) {
        if (nlpos) { T.keep(nlpos); dfree(ctx, nlpos); free_records(ctx, T, rb); }
        RC_TRY(T.alloc(&nlpos, cap));
        RC_TRY(alloc_records(T, rb, (size_t)cap + 1, tags));
        CUDA_TRY(cudaMemsetAsync(status, 0, ((size_t)ntiles + 1) * 8, ctx->stream));
        LAUNCH(ctx, sam_scan_k, ntiles, FT_T, 0, dtext, (uint32_t)nbytes, cap, cap + 1, status, (unsigned int *)(status + ntiles), nlpos, totals, tags ? 1 : 0, rb);
        CUDA_TRY(cudaMemcpyAsync(&n_nl, totals, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(&last, dtext + nbytes - 1, 1, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        if (n_nl <= cap) break;
        cap = n_nl;                                                // many very short lines: repeat with the exact size
    }
    rb.n = n_nl + ((last != '\n') ? 1 : 0);
    LAUNCH_CHECK();
    *out = rb;
    return 0;
}

static bool use_fused_tokenizer() {
    static const int v = [] { const char *e = getenv("WGBS_TOKENIZER"); return (e && !strcmp(e, "fused")) ? 1 : (e && !strcmp(e, "split")) ? 0 : WGBS_TOKENIZER_DEFAULT_FUSED; }();
    return v != 0;
}

int sam_tokenize(wgbs_ctx *ctx, const char *dtext, size_t nbytes, int want_tags /* 1 yes, 0 no, -1 decide from the first line */, Temps &T,
                 ReadBatch *out) {
    if (nbytes >= 0xfffffff0ull) return wgbs_set_err("SAM text must be < 4 GiB per call (got %zu); split on line boundaries", nbytes);
    if (use_fused_tokenizer()) return sam_tokenize_fused(ctx, dtext, nbytes, want_tags, T, out);
    const uint32_t ntiles = (uint32_t)((nbytes + NLS_TILE - 1) / NLS_TILE);
    ReadBatch rb;
    rb.text = dtext; rb.nbytes = (uint32_t)nbytes;
    uint32_t *nlpos = nullptr, *totals = ctx->d_flags + 8;
    unsigned long long *status = nullptr;
    // 1: warp tiles; 0: CTA tiles (default); -4 / -8 / -16: CTA tiles with batched loads; -1: CTA tiles fetched by TMA (staged)
    static const int warp_scan = [] {
        const char *e = getenv("WGBS_NLSCAN");
        if (e && !strncmp(e, "batch", 5)) { const int b = atoi(e + 5); return (b == 4 || b == 8 || b == 16) ? -b : 0; }
        if (e && !strcmp(e, "tma")) return -1;
        return (e && !strcmp(e, "warp")) ? 1 : (e && !strcmp(e, "cta")) ? 0 : WGBS_NLSCAN_DEFAULT_WARP;
    }();
    static const int lines_pf = [] { const char *e = getenv("WGBS_LINES_PF"); return e ? atoi(e) : WGBS_LINES_PF_DEFAULT; }();
    const uint32_t wtiles = (uint32_t)((nbytes + NLW_TILE - 1) / NLW_TILE);
    if (ntiles) RC_TRY(T.alloc(&status, (size_t)(warp_scan == 1 ? wtiles : ntiles) + 1));
    uint32_t cap = (uint32_t)(nbytes / 64 + 1024);            // optimistic: average line >= 64 bytes (a 50 bp SAM record is ~130)
    uint32_t n_nl = 0, first_mm = 0; char last = '\n';
    // first_line (patter.cpp:337-338): does the first non-empty line carry an MM tag?  Decided on the host from the head of the
    // text (one small D2H that rides on the synchronisation the line count needs anyway).
    std::vector<char> head(want_tags < 0 ? std::min<size_t>(nbytes, 1u << 16) : 0);
    for (int attempt = 0; attempt < 2 && ntiles; attempt++) {
        if (nlpos) { T.keep(nlpos); dfree(ctx, nlpos); }
        RC_TRY(T.alloc(&nlpos, cap));
        CUDA_TRY(cudaMemsetAsync(status, 0, ((size_t)(warp_scan == 1 ? wtiles : ntiles) + 1) * 8, ctx->stream));
        if (warp_scan == 1) LAUNCH(ctx, nl_scan_warp_k, (wtiles + NLW_T / 32 - 1) / (NLW_T / 32), NLW_T, 0, dtext, nbytes, cap, status, (unsigned int *)(status + wtiles), nlpos, totals);
        else if (warp_scan == -4) LAUNCH(ctx, nl_scan_k<4>, ntiles, NLS_T, 0, dtext, nbytes, cap, status, (unsigned int *)(status + ntiles), nlpos, totals);
        else if (warp_scan == -8) LAUNCH(ctx, nl_scan_k<8>, ntiles, NLS_T, 0, dtext, nbytes, cap, status, (unsigned int *)(status + ntiles), nlpos, totals);
        else if (warp_scan == -16) LAUNCH(ctx, nl_scan_k<16>, ntiles, NLS_T, 0, dtext, nbytes, cap, status, (unsigned int *)(status + ntiles), nlpos, totals);
        else if (warp_scan == -1) {
            CUDA_TRY(cudaFuncSetAttribute(nl_scan_k<-1>, cudaFuncAttributeMaxDynamicSharedMemorySize, NLS_TILE));
            LAUNCH(ctx, nl_scan_k<-1>, ntiles, NLS_T, NLS_TILE, dtext, nbytes, cap, status, (unsigned int *)(status + ntiles), nlpos, totals);
        }
        else LAUNCH(ctx, nl_scan_k<0>, ntiles, NLS_T, 0, dtext, nbytes, cap, status, (unsigned int *)(status + ntiles), nlpos, totals);
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
        if (lines_pf >= 4) LAUNCH(ctx, sam_lines_k<4>, grid_for(n_lines, 128), 128, 0, dtext, (uint32_t)nbytes, nlpos, n_nl, n_lines, tags ? 1 : 0, rb);
        else if (lines_pf >= 2) LAUNCH(ctx, sam_lines_k<2>, grid_for(n_lines, 128), 128, 0, dtext, (uint32_t)nbytes, nlpos, n_nl, n_lines, tags ? 1 : 0, rb);
        else LAUNCH(ctx, sam_lines_k<1>, grid_for(n_lines, 128), 128, 0, dtext, (uint32_t)nbytes, nlpos, n_nl, n_lines, tags ? 1 : 0, rb);
        LAUNCH_CHECK();
    }
    *out = rb;
    return 0;
}
