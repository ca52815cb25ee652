// collapse.cu -- the collapse step and pat text formatting.
//
// Replaces, byte for byte on the uncompressed text, the reference's
//     sort -k2,2n -k3,3 | uniq -c | awk -v OFS='\t' '{print $2,$3,$4,$1}'          (python/bam2pat.py:99-106, C locale)
// Order = (CpG index numeric, pattern bytes with a shorter prefix first).  With symbol codes '.'<'C'<'H'<'T' = 0..3
// packed MSB-first and zero padding, that order is the numeric order of (idx, word0, word1, ...): patterns never end
// in '.', so zero padding is unambiguous.  One stable 32-bit radix sort on (idx - idx_min, first few symbols) orders
// everything except ties between longer patterns, which fix_ties_k settles run by run.
#include <algorithm>

#include "reads.cuh"
#include "sort.cuh"

namespace {

__global__ void __launch_bounds__(256) idx_range_k(const uint32_t *__restrict__ idx, size_t n, uint32_t *__restrict__ mnmx) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    // idx has int32 semantics (`sort -k2,2n`): x ^ 0x80000000 maps it order-preservingly onto uint32
    uint32_t lo = i < n ? (idx[i] ^ 0x80000000u) : 0xffffffffu, hi = i < n ? (idx[i] ^ 0x80000000u) : 0u;
    lo = __reduce_min_sync(0xffffffffu, lo); hi = __reduce_max_sync(0xffffffffu, hi);
    __shared__ uint32_t slo[8], shi[8];
    if ((threadIdx.x & 31) == 0) { slo[threadIdx.x >> 5] = lo; shi[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {                                          // one atomic pair per CTA
        for (int w = 1; w < 8; w++) { lo = min(lo, slo[w]); hi = max(hi, shi[w]); }
        atomicMin(&mnmx[0], lo); atomicMax(&mnmx[1], hi);
    }
}
// ONE 32-bit sort key per record: (idx - idx_min) in the high bits, the first `nsym` pattern symbols below it
__global__ void __launch_bounds__(256) make_key_k(PatsView P, uint32_t idx_min, uint32_t nsym, uint32_t *__restrict__ keys, uint32_t *__restrict__ perm) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    const uint32_t w0 = P.len[i] ? P.pool[P.off[i]] : 0u;
    const uint32_t rel = (P.idx[i] ^ 0x80000000u) - idx_min;       // idx_min is in the same biased representation
    keys[i] = nsym ? ((rel << (2 * nsym)) | (w0 >> (32 - 2 * nsym))) : rel;
    perm[i] = (uint32_t)i;
}
// Records are radix-sorted on ONE 32-bit key = (idx - idx_min, first nsym symbols).  Whatever order remains to be decided
// lies inside runs of equal key that contain a pattern longer than nsym symbols: one thread orders such a run in place by
// the full patterns (zero padded = shorter prefix first).  Runs are short (templates starting at one CpG with the same first
// calls), so an insertion sort is enough; it is stable, like `sort`'s last-resort comparison needs.
__device__ __forceinline__ int pat_cmp(const PatsView &P, uint32_t a, uint32_t b) {
    const uint32_t na = (P.len[a] + 15) >> 4, nb = (P.len[b] + 15) >> 4, m = max(na, nb);
    const uint32_t *x = P.pool + P.off[a], *y = P.pool + P.off[b];
    for (uint32_t k = 0; k < m; k++) {
        const uint32_t u = k < na ? x[k] : 0u, v = k < nb ? y[k] : 0u;
        if (u != v) return u < v ? -1 : 1;
    }
    return 0;
}
// --long: `sort`'s last-resort comparison sees the whole line "chr \t idx \t pattern \t qname": after idx and pattern, the read name
__device__ __forceinline__ int name_cmp(const PatsView &P, uint32_t a, uint32_t b) {
    const unsigned char *x = (const unsigned char *)P.names + P.name_off[a], *y = (const unsigned char *)P.names + P.name_off[b];
    const uint32_t la = P.name_len[a], lb = P.name_len[b], m = min(la, lb);
    for (uint32_t k = 0; k < m; k++) if (x[k] != y[k]) return x[k] < y[k] ? -1 : 1;
    return la < lb ? -1 : (la > lb ? 1 : 0);
}
// full comparison of two records of one run.  mode 0: patterns never end in '.', zero padding orders them; 1 (--long): ties go to
// the read name; 2 (cview output: a clipped pattern may end in '.'): equal zero-padded words are ordered shorter first, as
// `sort -k3,3` orders "C" before "C."
__device__ __forceinline__ int rec_cmp(const PatsView &P, uint32_t a, uint32_t b, int by_name) {
    int c = pat_cmp(P, a, b);
    if (c == 0 && by_name == 1) c = name_cmp(P, a, b);
    if (c == 0 && by_name == 2) c = P.len[a] < P.len[b] ? -1 : (P.len[a] > P.len[b] ? 1 : 0);
    return c;
}
// Runs up to this length are ordered by ONE thread (insertion sort: at 30x WGBS a run is the few templates that start at one CpG
// with the same first calls); longer ones -- amplicon / targeted / RRBS data pile 1e5..1e6 templates onto one start CpG -- are
// queued for fix_long_runs_k, which sorts each with a whole CTA in O(n log^2 n) like the reference's `sort` stays O(n log n).
constexpr uint32_t TIES_SERIAL_MAX = 64;
__global__ void __launch_bounds__(128) fix_ties_k(PatsView P, uint32_t *__restrict__ perm, const uint32_t *__restrict__ key /*sorted*/, uint32_t nsym, int by_name,
                                                   uint2 *__restrict__ long_runs, uint32_t *__restrict__ n_long) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    const uint32_t k0 = key[i];
    if (i > 0 && key[i - 1] == k0) return;                          // only the first position of a run works
    size_t j = i + 1; bool any_long = P.len[perm[i]] > nsym;
    while (j < P.n && key[j] == k0) { any_long |= P.len[perm[j]] > nsym; j++; }
    if ((!any_long && !by_name) || j - i < 2) return;
    if (j - i > TIES_SERIAL_MAX) { long_runs[atomicAdd(n_long, 1u)] = make_uint2((uint32_t)i, (uint32_t)(j - i)); return; }
    for (size_t k = i + 1; k < j; k++) {
        const uint32_t v = perm[k]; size_t q = k;
        while (q > i) {
            if (rec_cmp(P, perm[q - 1], v, by_name) <= 0) break;
            perm[q] = perm[q - 1]; q--;
        }
        perm[q] = v;
    }
}
// One CTA per queued run (CTAs stride over the queue): bitonic sort of perm[start .. start + len) in global memory, positions past
// the run's end acting as +infinity.  Not stable -- records that compare equal are identical in everything the output shows
// (index, pattern, and in --long the name), so their order cannot be observed.
__global__ void __launch_bounds__(1024) fix_long_runs_k(PatsView P, uint32_t *__restrict__ perm, int by_name, const uint2 *__restrict__ long_runs,
                                                         const uint32_t *__restrict__ n_long) {
    const uint32_t nq = *n_long;
    for (uint32_t r = blockIdx.x; r < nq; r += gridDim.x) {
        uint32_t *a = perm + long_runs[r].x; const uint32_t n = long_runs[r].y;
        uint32_t np2 = 1; while (np2 < n) np2 <<= 1;
        // the all-ascending form of the network: the first step of a merge stage pairs position p with its mirror inside the
        // 2k-block (p ^ (k - 1)), the others pair p with p ^ j; virtual +infinity elements then never have to move
        for (uint32_t k = 2; k <= np2; k <<= 1) {
            for (uint32_t j = k >> 1; j > 0; j >>= 1) {
                const bool mirror = j == (k >> 1);
                for (uint32_t t = threadIdx.x; t < np2 / 2; t += blockDim.x) {
                    const uint32_t lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));                  // position without bit j
                    const uint32_t hi = mirror ? (lo ^ (k - 1)) : (lo | j);
                    if (hi >= n) continue;                                                       // partner is +infinity
                    const uint32_t x = a[lo], y = a[hi];
                    if (rec_cmp(P, x, y, by_name) > 0) { a[lo] = y; a[hi] = x; }
                }
                __syncthreads();
            }
        }
        __syncthreads();
    }
}
__global__ void __launch_bounds__(256) iota_k(uint32_t *__restrict__ p, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (uint32_t)i;
}
__global__ void __launch_bounds__(256) permute_long_k(PatsView P, const uint32_t *__restrict__ perm, uint32_t *__restrict__ o_idx, uint32_t *__restrict__ o_len,
                                                       uint32_t *__restrict__ o_off, uint32_t *__restrict__ o_cnt, uint32_t *__restrict__ o_noff,
                                                       uint32_t *__restrict__ o_nlen) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    const uint32_t r = perm[i];
    o_idx[i] = P.idx[r]; o_len[i] = P.len[r]; o_off[i] = P.off[r]; o_cnt[i] = P.count[r]; o_noff[i] = P.name_off[r]; o_nlen[i] = P.name_len[r];
}
__device__ __forceinline__ bool same_rec(const PatsView &P, uint32_t a, uint32_t b) {
    if (P.idx[a] != P.idx[b] || P.len[a] != P.len[b]) return false;
    const uint32_t nw = (P.len[a] + 15) >> 4;
    const uint32_t *x = P.pool + P.off[a], *y = P.pool + P.off[b];
    for (uint32_t k = 0; k < nw; k++) if (x[k] != y[k]) return false;
    return true;
}
__global__ void __launch_bounds__(256) head_flags_k(PatsView P, const uint32_t *__restrict__ perm, uint32_t *__restrict__ head) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    head[i] = (i == 0 || !same_rec(P, perm[i - 1], perm[i])) ? 1u : 0u;
}
__global__ void __launch_bounds__(256) emit_unique_k(PatsView P, const uint32_t *__restrict__ perm, const uint32_t *__restrict__ head,
                                                      const uint32_t *__restrict__ dst, uint32_t *__restrict__ o_idx, uint32_t *__restrict__ o_len,
                                                      uint32_t *__restrict__ o_off, uint32_t *__restrict__ o_cnt) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n || !head[i]) return;
    uint32_t r = perm[i];
    uint32_t c = P.count[r];
    for (size_t j = i + 1; j < P.n && !head[j]; j++) c += P.count[perm[j]];       // uniq -c
    uint32_t k = dst[i];
    o_idx[k] = P.idx[r]; o_len[k] = P.len[r]; o_off[k] = P.off[r]; o_cnt[k] = c;
}

// ---- text ----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ndigits_i32(int32_t v) {
    uint32_t n = v < 0 ? 1 : 0; uint32_t u = v < 0 ? (uint32_t)(-(int64_t)v) : (uint32_t)v;
    uint32_t d = 1; while (u >= 10) { u /= 10; d++; }
    return n + d;
}
__device__ __forceinline__ char *put_i32(char *p, int32_t v) {
    uint32_t u = v < 0 ? (uint32_t)(-(int64_t)v) : (uint32_t)v;
    if (v < 0) *p++ = '-';
    char tmp[10]; int k = 0;
    do { tmp[k++] = (char)('0' + u % 10); u /= 10; } while (u);
    while (k) *p++ = tmp[--k];
    return p;
}
__global__ void __launch_bounds__(256) line_len_k(PatsView P, uint32_t chrom_len, uint32_t *__restrict__ ll) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    ll[i] = chrom_len + 1 + ndigits_i32((int32_t)P.idx[i]) + 1 + P.len[i] + 1 + ndigits_i32((int32_t)P.count[i]) + 1;
}
__global__ void __launch_bounds__(256) line_write_k(PatsView P, const char *__restrict__ chrom, uint32_t chrom_len,
                                                     const uint64_t *__restrict__ loff, char *__restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    char *p = out + loff[i];
    for (uint32_t k = 0; k < chrom_len; k++) *p++ = chrom[k];
    *p++ = '\t';
    p = put_i32(p, (int32_t)P.idx[i]);
    *p++ = '\t';
    const uint32_t L = P.len[i];
    const uint32_t *wp = P.pool + P.off[i];
    for (uint32_t b = 0; b < L; b += 16) {
        uint32_t w = *wp++, m = min(16u, L - b);
        for (uint32_t k = 0; k < m; k++) *p++ = sym_char((w >> (30 - 2 * k)) & 3u);
    }
    *p++ = '\t';
    p = put_i32(p, (int32_t)P.count[i]);
    *p++ = '\n';
}

__global__ void __launch_bounds__(256) line_len_long_k(PatsView P, uint32_t chrom_len, uint32_t *__restrict__ ll) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    ll[i] = chrom_len + 1 + ndigits_i32((int32_t)P.idx[i]) + 1 + P.len[i] + 1 + 1 + 1 + P.name_len[i] + 1;      // ... \t 1 \t qname \n
}
__global__ void __launch_bounds__(256) line_write_long_k(PatsView P, const char *__restrict__ chrom, uint32_t chrom_len,
                                                          const uint64_t *__restrict__ loff, char *__restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    char *p = out + loff[i];
    for (uint32_t k = 0; k < chrom_len; k++) *p++ = chrom[k];
    *p++ = '\t';
    p = put_i32(p, (int32_t)P.idx[i]);
    *p++ = '\t';
    const uint32_t L = P.len[i];
    const uint32_t *wp = P.pool + P.off[i];
    for (uint32_t b = 0; b < L; b += 16) {
        uint32_t w = *wp++, m = min(16u, L - b);
        for (uint32_t k = 0; k < m; k++) *p++ = sym_char((w >> (30 - 2 * k)) & 3u);
    }
    *p++ = '\t'; *p++ = '1'; *p++ = '\t';
    const char *nm = P.names + P.name_off[i];
    for (uint32_t k = 0; k < P.name_len[i]; k++) *p++ = nm[k];
    *p++ = '\n';
}


}  // namespace

// mode: WGBS_COLLAPSE_* (include/wgbs_b200.h)
static int collapse_impl(wgbs_ctx *ctx, wgbs_pats *P, int mode) {
    const bool long_mode = mode == WGBS_COLLAPSE_LONG;
    RC_TRY(wgbs_ctx_activate(ctx));
    if (!P) return wgbs_set_err("null pats");
    if (long_mode && P->n && !P->names) return wgbs_set_err("wgbs_collapse_long: records carry no read names (set opts.keep_names in the pileup)");
    const size_t n = P->n;
    if (n < 1) return 0;
    if (n >= 0xffffffffull) return wgbs_set_err("wgbs_collapse: too many records");
    Temps T(ctx);
    uint32_t *k0, *v0, *k1, *v1;
    RC_TRY(T.alloc(&k0, n)); RC_TRY(T.alloc(&v0, n)); RC_TRY(T.alloc(&k1, n)); RC_TRY(T.alloc(&v1, n));
    uint32_t *k = k0, *v = v0, *ka = k1, *va = v1;
    PatsView pv = view_of(P);
    // key layout from the index range of this batch
    uint32_t *d_mm = ctx->d_flags + 12; const uint32_t init[2] = {0xffffffffu, 0u}; uint32_t mm[2];
    CUDA_TRY(cudaMemcpyAsync(d_mm, init, 8, cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH(ctx, idx_range_k, grid_for(n, 256), 256, 0, P->idx, n, d_mm);
    CUDA_TRY(cudaMemcpyAsync(mm, d_mm, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    uint32_t range = mm[1] - mm[0], ibits = 0; while (ibits < 32 && (range >> ibits)) ibits++;
    const uint32_t nsym = (32 - ibits) / 2 > 16 ? 16 : (32 - ibits) / 2;
    if (mode == WGBS_COLLAPSE_ADJACENT) {
        LAUNCH(ctx, iota_k, grid_for(n, 256), 256, 0, v, n);          // no sort: only adjacent equal records merge (collapse_pat.pl on unsorted input)
    } else {
        LAUNCH(ctx, make_key_k, grid_for(n, 256), 256, 0, pv, mm[0], nsym, k, v);
        RC_TRY(radix_sort_pairs(ctx, &k, &v, &ka, &va, n));
        // longer patterns (and names): order inside equal-key runs; runs too long for one thread are queued and sorted by whole CTAs
        const int by = long_mode ? 1 : (mode == WGBS_COLLAPSE_DOTTED ? 2 : 0);
        uint2 *long_runs; uint32_t *n_long = ctx->d_flags + 14;
        RC_TRY(T.alloc(&long_runs, n / TIES_SERIAL_MAX + 1));
        CUDA_TRY(cudaMemsetAsync(n_long, 0, 4, ctx->stream));
        LAUNCH(ctx, fix_ties_k, grid_for(n, 128), 128, 0, pv, v, k, nsym, by, long_runs, n_long);
        LAUNCH(ctx, fix_long_runs_k, (unsigned)std::min<size_t>(n / TIES_SERIAL_MAX + 1, (size_t)ctx->sm_count), 1024, 0, pv, v, by, long_runs, n_long);
    }
    if (long_mode) {
        // no uniq in --long: the records are only re-ordered
        uint32_t *o_idx, *o_len, *o_off, *o_cnt, *o_noff, *o_nlen;
        int rc2;
        if ((rc2 = dalloc(ctx, &o_idx, n)) < 0 || (rc2 = dalloc(ctx, &o_len, n)) < 0 || (rc2 = dalloc(ctx, &o_off, n + 1)) < 0 || (rc2 = dalloc(ctx, &o_cnt, n)) < 0 ||
            (rc2 = dalloc(ctx, &o_noff, n + 1)) < 0 || (rc2 = dalloc(ctx, &o_nlen, n)) < 0) return rc2;
        LAUNCH(ctx, permute_long_k, grid_for(n, 256), 256, 0, pv, v, o_idx, o_len, o_off, o_cnt, o_noff, o_nlen);
        LAUNCH_CHECK();
        dfree(ctx, P->idx); dfree(ctx, P->len); dfree(ctx, P->off); dfree(ctx, P->count); dfree(ctx, P->name_off); dfree(ctx, P->name_len);
        P->idx = o_idx; P->len = o_len; P->off = o_off; P->count = o_cnt; P->name_off = o_noff; P->name_len = o_nlen;
        return 0;
    }
    // run-length: heads, destinations, counts
    uint32_t *head, *dst;
    RC_TRY(T.alloc(&head, n)); RC_TRY(T.alloc(&dst, n + 1));
    LAUNCH(ctx, head_flags_k, grid_for(n, 256), 256, 0, pv, v, head);
    RC_TRY(scan_u32_u32(ctx, head, dst, n));
    uint32_t nu = 0;
    CUDA_TRY(cudaMemcpyAsync(&nu, dst + n, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    uint32_t *o_idx, *o_len, *o_off, *o_cnt;
    int rc;
    if ((rc = dalloc(ctx, &o_idx, nu)) < 0) return rc;
    if ((rc = dalloc(ctx, &o_len, nu)) < 0) { dfree(ctx, o_idx); return rc; }
    if ((rc = dalloc(ctx, &o_off, (size_t)nu + 1)) < 0) { dfree(ctx, o_idx); dfree(ctx, o_len); return rc; }
    if ((rc = dalloc(ctx, &o_cnt, nu)) < 0) { dfree(ctx, o_idx); dfree(ctx, o_len); dfree(ctx, o_off); return rc; }
    LAUNCH(ctx, emit_unique_k, grid_for(n, 256), 256, 0, pv, v, head, dst, o_idx, o_len, o_off, o_cnt);
    LAUNCH_CHECK();
    dfree(ctx, P->idx); dfree(ctx, P->len); dfree(ctx, P->off); dfree(ctx, P->count);
    P->idx = o_idx; P->len = o_len; P->off = o_off; P->count = o_cnt; P->n = nu;
    return 0;
}

extern "C" int wgbs_collapse(wgbs_ctx *ctx, wgbs_pats *P) { return collapse_impl(ctx, P, WGBS_COLLAPSE_SORTED); }
extern "C" int wgbs_collapse_long(wgbs_ctx *ctx, wgbs_pats *P) { return collapse_impl(ctx, P, WGBS_COLLAPSE_LONG); }
extern "C" int wgbs_collapse_ex(wgbs_ctx *ctx, wgbs_pats *P, int mode) {
    if (mode < WGBS_COLLAPSE_SORTED || mode > WGBS_COLLAPSE_ADJACENT) return wgbs_set_err("wgbs_collapse_ex: unknown mode %d", mode);
    return collapse_impl(ctx, P, mode);
}

// Write "chrom \t idx \t pattern \t count \n" per record, in record order (reference docs/pat_format.md:3-47).
// out == NULL: only *nbytes is computed.  out may be host or device.
extern "C" int wgbs_pats_format(wgbs_ctx *ctx, const wgbs_pats *P, const char *chrom, char *out, size_t cap, size_t *nbytes) {
    RC_TRY(wgbs_ctx_activate(ctx));
    if (!P || !chrom) return wgbs_set_err("wgbs_pats_format: null argument");
    const size_t n = P->n;
    const uint32_t cl = (uint32_t)strlen(chrom);
    Temps T(ctx);
    uint32_t *ll; uint64_t *loff;
    RC_TRY(T.alloc(&ll, n)); RC_TRY(T.alloc(&loff, n + 1));
    if (n) LAUNCH(ctx, line_len_k, grid_for(n, 256), 256, 0, view_of(P), cl, ll);
    RC_TRY(scan_u32_u64(ctx, ll, loff, n));
    uint64_t total = 0;
    CUDA_TRY(cudaMemcpyAsync(&total, loff + n, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (nbytes) *nbytes = (size_t)total;
    if (!out) return 0;
    if (cap < total) return wgbs_set_err("wgbs_pats_format: buffer too small (%zu < %llu)", cap, (unsigned long long)total);
    char *dchrom; RC_TRY(T.alloc(&dchrom, (size_t)cl + 1));
    RC_TRY(copy_any(ctx, dchrom, chrom, cl + 1));
    char *dout = out;
    if (!is_device_ptr(out)) RC_TRY(T.alloc(&dout, (size_t)total));
    if (n) { LAUNCH(ctx, line_write_k, grid_for(n, 256), 256, 0, view_of(P), dchrom, cl, loff, dout); LAUNCH_CHECK(); }
    if (dout != out) RC_TRY(copy_any(ctx, out, dout, (size_t)total));
    return 0;
}

extern "C" int wgbs_pats_format_long(wgbs_ctx *ctx, const wgbs_pats *P, const char *chrom, char *out, size_t cap, size_t *nbytes) {
    RC_TRY(wgbs_ctx_activate(ctx));
    if (!P || !chrom) return wgbs_set_err("wgbs_pats_format_long: null argument");
    const size_t n = P->n;
    if (n && !P->names) return wgbs_set_err("wgbs_pats_format_long: records carry no read names (set opts.keep_names in the pileup)");
    const uint32_t cl = (uint32_t)strlen(chrom);
    Temps T(ctx);
    uint32_t *ll; uint64_t *loff;
    RC_TRY(T.alloc(&ll, n)); RC_TRY(T.alloc(&loff, n + 1));
    if (n) LAUNCH(ctx, line_len_long_k, grid_for(n, 256), 256, 0, view_of(P), cl, ll);
    RC_TRY(scan_u32_u64(ctx, ll, loff, n));
    uint64_t total = 0;
    CUDA_TRY(cudaMemcpyAsync(&total, loff + n, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (nbytes) *nbytes = (size_t)total;
    if (!out) return 0;
    if (cap < total) return wgbs_set_err("wgbs_pats_format_long: buffer too small (%zu < %llu)", cap, (unsigned long long)total);
    char *dchrom; RC_TRY(T.alloc(&dchrom, (size_t)cl + 1));
    RC_TRY(copy_any(ctx, dchrom, chrom, cl + 1));
    char *dout = out;
    if (!is_device_ptr(out)) RC_TRY(T.alloc(&dout, (size_t)total));
    if (n) { LAUNCH(ctx, line_write_long_k, grid_for(n, 256), 256, 0, view_of(P), dchrom, cl, loff, dout); LAUNCH_CHECK(); }
    if (dout != out) RC_TRY(copy_any(ctx, out, dout, (size_t)total));
    return 0;
}
