// pair.cu -- template pairing: which two records are mates.
//
// Replaces the reference's `match_maker` (pipeline_wgbs/match_maker.cpp:48-183) and patter's adjacent-QNAME pairing
// (pipeline_wgbs/patter.cpp:395-412).  match_maker sorts buffered SAM lines as whole strings, pairs adjacent lines
// with equal QNAME greedily and lets everything else through as singles; for a coordinate-sorted single-chromosome
// stream with consistent PNEXT that is exactly "group records by QNAME; within a group, in whole-line order, pair
// greedily".  Here: sort record ids by the QNAME hash (one stable 32-bit radix sort), then one thread per
// hash run: runs of 1 are singles, runs of 2 with byte-equal names are a pair, anything else (3+ records, or a hash
// collision) is ordered by whole-line bytes by that thread and paired greedily.  (Only the low 32 hash bits are sorted
// on; names are always compared byte for byte, so collisions cost time, never correctness.)
//
// Output: mate[r] = record id of r's mate or NONE.  The template's slot is the smaller record id of the two, so
// templates stay in coordinate order.
#include "reads.cuh"
#include "sort.cuh"

namespace {

constexpr uint32_t NONE = 0xffffffffu;

__global__ void __launch_bounds__(256) fill_u32_k(uint32_t *p, size_t n, uint32_t v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
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

}  // namespace

int build_mates(wgbs_ctx *ctx, const ReadBatch &rb, bool paired, Temps &T, uint32_t **mate_out, unsigned long long *d_stats) {
    const uint32_t n = rb.n;
    uint32_t *mate;
    RC_TRY(T.alloc(&mate, n));
    if (n) LAUNCH(ctx, fill_u32_k, grid_for(n, 256), 256, 0, mate, (size_t)n, NONE);
    *mate_out = mate;
    if (!paired || n < 2) { LAUNCH_CHECK(); return 0; }
    // sort record ids by the LOW 32 bits of the QNAME hash only (4 digit passes): equal-key runs then hold the mates plus
    // the occasional 32-bit collision, which pair_runs_k separates by comparing the name bytes themselves.
    uint32_t *k0, *v0, *k1, *v1;
    RC_TRY(T.alloc(&k0, n)); RC_TRY(T.alloc(&v0, n)); RC_TRY(T.alloc(&k1, n)); RC_TRY(T.alloc(&v1, n));
    CUDA_TRY(cudaMemcpyAsync(k0, rb.hash_lo, (size_t)n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    RC_TRY(fill_iota(ctx, v0, n));
    uint32_t *k = k0, *v = v0, *ka = k1, *va = v1;
    RC_TRY(radix_sort_pairs(ctx, &k, &v, &ka, &va, n));
    LAUNCH(ctx, pair_runs_k, grid_for(n, 256), 256, 0, view_of(rb), v, k, n, mate, d_stats);
    LAUNCH_CHECK();
    return 0;
}
