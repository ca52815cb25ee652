// pair.cu -- template pairing: which two records are mates.
//
// Replaces the reference's `match_maker` (pipeline_wgbs/match_maker.cpp:48-183) and patter's adjacent-QNAME pairing
// (pipeline_wgbs/patter.cpp:395-412).  match_maker sorts buffered SAM lines as whole strings, pairs adjacent lines
// with equal QNAME greedily and lets everything else through as singles; for a coordinate-sorted single-chromosome
// stream with consistent PNEXT that is exactly "group records by QNAME; within a group, in whole-line order, pair
// greedily".  Here: every record drops its 64-bit QNAME hash into an open-addressing table in global memory (L2-resident:
// 24 B per slot, load <= 0.5) with atomicCAS and adds itself to the slot's (count, min id, max id).  A slot with exactly two
// records whose names are byte-equal is a pair -- the common case, settled without any sort.  Records in slots holding 3+
// records (supplementary alignments sharing a QNAME, or the astronomically rare 64-bit collision) are rare: they are gathered, radix-sorted by
// hash, and each equal-hash run is ordered by whole-line bytes and paired greedily by one thread (pair_runs_k).  Names are
// always compared byte for byte, so hash collisions cost time, never correctness.
//
// Output: mate[r] = record id of r's mate or NONE.  The template's slot is the smaller record id of the two, so
// templates stay in coordinate order.
#include <stdlib.h>

#include "reads.cuh"
#include "sort.cuh"

namespace {

constexpr uint32_t NONE = 0xffffffffu;

__global__ void __launch_bounds__(256) fill_u32_k(uint32_t *p, size_t n, uint32_t v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
struct Slot { unsigned long long key; uint32_t cnt, mn, mx, pad; };
constexpr uint32_t EMPTY = 0xffffffffu;
constexpr unsigned long long EMPTY_KEY = ~0ull;
__device__ __forceinline__ unsigned long long rec_key(const ReadBatchView &rb, uint32_t r, unsigned long long hash_mask) {
    unsigned long long h = (((unsigned long long)rb.hash_hi[r] << 32) | rb.hash_lo[r]) & hash_mask;
    return h == EMPTY_KEY ? EMPTY_KEY - 1 : h;
}

__global__ void __launch_bounds__(256) slots_init_k(Slot *__restrict__ tab, size_t nslots) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nslots) { Slot s; s.key = EMPTY_KEY; s.cnt = 0; s.mn = EMPTY; s.mx = 0; s.pad = 0; tab[i] = s; }
}
__global__ void __launch_bounds__(256) pair_insert_k(ReadBatchView rb, Slot *__restrict__ tab, uint32_t mask, unsigned long long hash_mask, uint32_t *__restrict__ slot_of) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rb.n) return;
    if (rb.status[r] == REC_BLANK) { slot_of[r] = EMPTY; return; }
    const unsigned long long h = rec_key(rb, r, hash_mask);
    uint32_t s = (uint32_t)((h * 0x9e3779b97f4a7c15ULL) >> 32) & mask;
    while (true) {
        const unsigned long long old = atomicCAS(&tab[s].key, EMPTY_KEY, h);
        if (old == EMPTY_KEY || old == h) { atomicAdd(&tab[s].cnt, 1u); atomicMin(&tab[s].mn, r); atomicMax(&tab[s].mx, r); slot_of[r] = s; return; }
        s = (s + 1) & mask;
    }
}

// whole-line comparison, std::string operator< semantics (unsigned bytes, shorter prefix first)
__device__ int line_cmp(const ReadBatchView &rb, uint32_t a, uint32_t b) {
    const unsigned char *x = (const unsigned char *)rb.text + rb.line_off[a], *y = (const unsigned char *)rb.text + rb.line_off[b];
    uint32_t la = rb.line_len[a], lb = rb.line_len[b], m = la < lb ? la : lb;
    for (uint32_t i = 0; i < m; i++) if (x[i] != y[i]) return x[i] < y[i] ? -1 : 1;
    return la < lb ? -1 : (la > lb ? 1 : 0);
}
__device__ bool name_eq(const ReadBatchView &rb, uint32_t a, uint32_t b) {
    uint32_t la = rb.qn_len[a];
    if (la != rb.qn_len[b]) return false;
    const char *x = rb.text + rb.line_off[a], *y = rb.text + rb.line_off[b];
    for (uint32_t i = 0; i < la; i++) if (x[i] != y[i]) return false;
    return true;
}

__global__ void __launch_bounds__(256) pair_runs_k(ReadBatchView rb, uint32_t *__restrict__ perm, const uint32_t *__restrict__ hlo,
                                                    uint32_t n, uint32_t *__restrict__ mate, unsigned long long *__restrict__ stats) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t npairs = 0;
    if (i < n) {
        const uint32_t lo = hlo[i];
        bool head = (i == 0) || hlo[i - 1] != lo;
        if (head) {
            uint32_t j = i + 1;
            while (j < n && hlo[j] == lo) j++;
            uint32_t run = j - i;
            if (run == 2) {
                uint32_t a = perm[i], b = perm[i + 1];
                if (rb.status[a] != REC_BLANK && rb.status[b] != REC_BLANK && name_eq(rb, a, b)) { mate[a] = b; mate[b] = a; npairs = 1; }
            } else if (run > 2) {
                // rare: supplementary alignments sharing a QNAME, or a 64-bit hash collision.  Order by whole line
                // (match_maker.cpp:60) with an in-place insertion sort, then pair adjacent equal names greedily (:62-76).
                for (uint32_t k = i + 1; k < j; k++) {
                    uint32_t v = perm[k]; uint32_t q = k;
                    while (q > i && line_cmp(rb, perm[q - 1], v) > 0) { perm[q] = perm[q - 1]; q--; }
                    perm[q] = v;
                }
                for (uint32_t k = i; k + 1 < j; k++) {
                    uint32_t a = perm[k], b = perm[k + 1];
                    if (rb.status[a] != REC_BLANK && rb.status[b] != REC_BLANK && name_eq(rb, a, b)) { mate[a] = b; mate[b] = a; npairs++; k++; }
                }
            }
        }
    }
    // one atomic per warp
    for (int d = 16; d >= 1; d >>= 1) npairs += __shfl_xor_sync(0xffffffffu, npairs, d);
    if ((threadIdx.x & 31) == 0 && npairs) atomicAdd(&stats[ST_PAIRS], (unsigned long long)npairs);
}

// one thread per record: slots of two byte-equal names are pairs; members of fuller slots go to the slow list
__global__ void __launch_bounds__(256) pair_resolve_k(ReadBatchView rb, const Slot *__restrict__ tab, const uint32_t *__restrict__ slot_of,
                                                       unsigned long long hash_mask, uint32_t *__restrict__ mate, uint32_t *__restrict__ slow_key,
                                                       uint32_t *__restrict__ slow_id, uint32_t *__restrict__ n_slow,
                                                       unsigned long long *__restrict__ stats) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t npairs = 0;
    if (r < rb.n && slot_of[r] != EMPTY) {
        const Slot s = tab[slot_of[r]];
        if (s.cnt == 2) {
            if (r == s.mn && name_eq(rb, s.mn, s.mx)) { mate[s.mn] = s.mx; mate[s.mx] = s.mn; npairs = 1; }
        } else if (s.cnt > 2) {
            const uint32_t k = atomicAdd(n_slow, 1u);
            slow_key[k] = slot_of[r]; slow_id[k] = r;                // the slot index identifies the 64-bit key uniquely
        }
    }
    for (int d = 16; d >= 1; d >>= 1) npairs += __shfl_xor_sync(0xffffffffu, npairs, d);
    if ((threadIdx.x & 31) == 0 && npairs) atomicAdd(&stats[ST_PAIRS], (unsigned long long)npairs);
}

}  // namespace

int build_mates(wgbs_ctx *ctx, const ReadBatch &rb, bool paired, Temps &T, uint32_t **mate_out, unsigned long long *d_stats) {
    const uint32_t n = rb.n;
    uint32_t *mate;
    RC_TRY(T.alloc(&mate, n));
    if (n) LAUNCH(ctx, fill_u32_k, grid_for(n, 256), 256, 0, mate, (size_t)n, NONE);
    *mate_out = mate;
    if (!paired || n < 2) { LAUNCH_CHECK(); return 0; }
    // hash table: slots = next power of two >= 2n
    size_t nslots = 1; while (nslots < (size_t)n * 2) nslots <<= 1;
    unsigned long long hash_mask = ~0ull;
    if (const char *e = getenv("WGBS_PAIR_HASH_BITS")) { int b = atoi(e); if (b > 0 && b < 64) hash_mask = (1ull << b) - 1; }   // test hook: force collisions
    Slot *tab; uint32_t *slot_of, *slow_key, *slow_id, *n_slow = ctx->d_flags + 16;
    RC_TRY(T.alloc(&tab, nslots)); RC_TRY(T.alloc(&slot_of, n)); RC_TRY(T.alloc(&slow_key, n)); RC_TRY(T.alloc(&slow_id, n));
    CUDA_TRY(cudaMemsetAsync(n_slow, 0, 4, ctx->stream));
    LAUNCH(ctx, slots_init_k, grid_for(nslots, 256), 256, 0, tab, nslots);
    LAUNCH(ctx, pair_insert_k, grid_for(n, 256), 256, 0, view_of(rb), tab, (uint32_t)(nslots - 1), hash_mask, slot_of);
    LAUNCH(ctx, pair_resolve_k, grid_for(n, 256), 256, 0, view_of(rb), tab, slot_of, hash_mask, mate, slow_key, slow_id, n_slow, d_stats);
    uint32_t m = 0;
    CUDA_TRY(cudaMemcpyAsync(&m, n_slow, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (m >= 2) {
        // rare path: order the crowded slots' members by hash, then whole-line order + greedy pairing inside every equal-hash run
        uint32_t *ka, *va;
        RC_TRY(T.alloc(&ka, m)); RC_TRY(T.alloc(&va, m));
        uint32_t *k = slow_key, *v = slow_id;
        RC_TRY(radix_sort_pairs(ctx, &k, &v, &ka, &va, m));
        ReadBatchView vw = view_of(rb);
        if (rb.bam) {
            // BAM batch: the whole-line order is the order of the SAM lines samtools would print; format just these few records
            if (!rb.side) return wgbs_set_err("build_mates: BAM batch without its record table");
            char *stext; uint32_t *soff, *slen;
            RC_TRY(bam_side_lines(ctx, *rb.side, v, m, n, T, &stext, &soff, &slen));      // v: the m ids (sorted by slot)
            vw.text = stext; vw.line_off = soff; vw.line_len = slen;          // QNAME starts the line: name_eq keeps working
        }
        LAUNCH(ctx, pair_runs_k, grid_for(m, 256), 256, 0, vw, v, k, m, mate, d_stats);
    }
    LAUNCH_CHECK();
    return 0;
}
