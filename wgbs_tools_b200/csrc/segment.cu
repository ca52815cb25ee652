// segment.cu -- the segment_betas dynamic-programming segmenter.
//
// Reference: src/segment_betas/segmentor.cpp:60-159 (dp), :50-58 (traceback), :164-190 (read_beta_file validation).
// For a chunk of n sites and K beta files the reference fills, row by row, cost[i][j] = sum_k ll_k(i, i+j) (block of
// j+1 sites starting at i; -inf when the block is longer than max_bp) and folds it into
//     M[i+1] = max_{k in [i+1-max_cpg, i]} M[k] + cost[k][i-k]      (strict '>' scanning k upwards: lowest k wins ties).
//
// Here, per wave of chunks:
//   seg_prefix_*   uint32 running sums of meth / cover per dataset (differences are exact integers < 2^24, which is
//                  what the reference's float accumulators hold)
//   seg_window_k   per END site e: how many starts i are admissible (max_cpg and max_bp; dists are non-decreasing,
//                  so the admissible starts are a contiguous run ending at e)  -> cell offsets (end-major layout)
//   seg_cost_k     one thread per cell (e, j): the k-loop in file order with the reference's exact float/double mix and
//                  bit-exact glibc log2f/log2 (glibc_log2.cuh).  The double sum over datasets is order dependent, so one
//                  thread owns a cell's whole k-loop.
//   seg_dp_k       one WARP per chunk: sequential over e, parallel max over the admissible block lengths, M in a
//                  shared-memory ring, next step's cells prefetched.  End-major layout: the candidates of step e are contiguous.
//   seg_trace_k    traceback (segmentor.cpp:50-58)
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "glibc_log2.cuh"

namespace {

// ---- running sums ---------------------------------------------------------------------------------------------------
constexpr int PF_T = 256, PF_I = 8, PF_TILE = PF_T * PF_I;

__device__ __forceinline__ uint2 warp_incl_scan2(uint2 v) {
    const unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t a = __shfl_up_sync(0xffffffffu, v.x, d), b = __shfl_up_sync(0xffffffffu, v.y, d);
        if (lane >= (unsigned)d) { v.x += a; v.y += b; }
    }
    return v;
}
// tile sums: grid (ntiles, K)
__global__ void __launch_bounds__(PF_T) seg_prefix_sums_k(const uint8_t *const *__restrict__ betas, uint32_t s0, uint32_t ns,
                                                          uint32_t ntiles, uint2 *__restrict__ tsum, uint32_t *__restrict__ bad) {
    const uint8_t *b = betas[blockIdx.y] + (size_t)s0 * 2;
    uint32_t base = blockIdx.x * PF_TILE;
    uint2 s = make_uint2(0, 0);
#pragma unroll
    for (int i = 0; i < PF_I; i++) {
        uint32_t k = base + i * PF_T + threadIdx.x;
        if (k < ns) { uint32_t m = b[2 * (size_t)k], c = b[2 * (size_t)k + 1]; s.x += m; s.y += c; if (m > c) atomicOr(bad, 1u); }   // segmentor.cpp:181-187
    }
    __shared__ uint2 ws[PF_T / 32];
    s.x = __reduce_add_sync(0xffffffffu, s.x); s.y = __reduce_add_sync(0xffffffffu, s.y);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) { uint2 t = make_uint2(0, 0); for (int i = 0; i < PF_T / 32; i++) { t.x += ws[i].x; t.y += ws[i].y; } tsum[(size_t)blockIdx.y * ntiles + blockIdx.x] = t; }
}
// exclusive scan of tile sums per dataset: grid (K), one CTA loops
__global__ void __launch_bounds__(256) seg_prefix_tiles_k(uint2 *__restrict__ tsum, uint32_t ntiles) {
    uint2 *t = tsum + (size_t)blockIdx.x * ntiles;
    __shared__ uint2 ws[8]; __shared__ uint2 carry;
    if (threadIdx.x == 0) carry = make_uint2(0, 0);
    __syncthreads();
    for (uint32_t base = 0; base < ntiles; base += 256) {
        uint32_t k = base + threadIdx.x;
        uint2 v = k < ntiles ? t[k] : make_uint2(0, 0);
        uint2 inc = warp_incl_scan2(v);
        if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = inc;
        __syncthreads();
        uint2 off = carry;
        for (unsigned w = 0; w < (threadIdx.x >> 5); w++) { off.x += ws[w].x; off.y += ws[w].y; }
        if (k < ntiles) t[k] = make_uint2(off.x + inc.x - v.x, off.y + inc.y - v.y);
        __syncthreads();
        if (threadIdx.x == 255) carry = make_uint2(off.x + inc.x, off.y + inc.y);
        __syncthreads();
    }
}
// P[k][s] for s in [0, ns]: grid (ntiles, K).  Thread t owns PF_I consecutive sites.
__global__ void __launch_bounds__(PF_T) seg_prefix_write_k(const uint8_t *const *__restrict__ betas, uint32_t s0, uint32_t ns,
                                                           uint32_t ntiles, const uint2 *__restrict__ tsum, uint32_t *__restrict__ Pm,
                                                           uint32_t *__restrict__ Pt) {
    const uint8_t *b = betas[blockIdx.y] + (size_t)s0 * 2;
    uint32_t base = blockIdx.x * PF_TILE + threadIdx.x * PF_I;
    uint32_t m[PF_I], c[PF_I]; uint2 s = make_uint2(0, 0);
#pragma unroll
    for (int i = 0; i < PF_I; i++) { uint32_t k = base + i; m[i] = k < ns ? b[2 * (size_t)k] : 0; c[i] = k < ns ? b[2 * (size_t)k + 1] : 0; s.x += m[i]; s.y += c[i]; }
    uint2 inc = warp_incl_scan2(s);
    __shared__ uint2 ws[PF_T / 32];
    if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = inc;
    __syncthreads();
    uint2 off = tsum[(size_t)blockIdx.y * ntiles + blockIdx.x];
    for (unsigned w = 0; w < (threadIdx.x >> 5); w++) { off.x += ws[w].x; off.y += ws[w].y; }
    off.x += inc.x - s.x; off.y += inc.y - s.y;
    uint32_t *pm = Pm + (size_t)blockIdx.y * (ns + 1), *pt = Pt + (size_t)blockIdx.y * (ns + 1);
#pragma unroll
    for (int i = 0; i < PF_I; i++) {
        uint32_t k = base + i;
        if (k <= ns) { pm[k] = off.x; pt[k] = off.y; }
        off.x += m[i]; off.y += c[i];
    }
}

// ---- admissible window per end site ---------------------------------------------------------------------------------
struct Chunk { uint32_t start, n; };   // relative to the wave's first site
__global__ void __launch_bounds__(256) seg_window_k(const uint32_t *__restrict__ dists, const uint32_t *__restrict__ chunk_of,
                                                     const Chunk *__restrict__ chunks, uint32_t ns, uint32_t max_cpg, uint32_t max_bp,
                                                     uint32_t *__restrict__ W, uint32_t *__restrict__ bad) {
    uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ns) return;
    if (chunk_of[e] == 0xffffffffu) { W[e] = 0; return; }       // site in a gap between chunks of this wave
    const Chunk c = chunks[chunk_of[e]];
    uint32_t lo = c.start;
    if (e + 1 - c.start > max_cpg) lo = e + 1 - max_cpg;
    if (e > c.start && dists[e] < dists[e - 1]) atomicOr(bad, 2u);   // the port assumes loci ascend inside a chunk
    // first i in [lo, e] with dists[e] - dists[i] <= max_bp
    const uint32_t de = dists[e];
    uint32_t a = lo, b = e;
    while (a < b) { uint32_t m = (a + b) >> 1; if (de - dists[m] > max_bp) a = m + 1; else b = m; }
    W[e] = e - a + 1;
}

// ---- cost of one cell ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) seg_cost_k(const uint32_t *__restrict__ Pm, const uint32_t *__restrict__ Pt, uint32_t ns, int K,
                                                   const uint64_t *__restrict__ coff, uint64_t ncells, float ps, double *__restrict__ cost) {
    uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    // end site e: last e with coff[e] <= c
    uint32_t lo = 0, hi = ns;
    while (hi - lo > 1) { uint32_t m = (lo + hi) >> 1; if (coff[m] <= c) lo = m; else hi = m; }
    const uint32_t e = lo, j = (uint32_t)(c - coff[e]), i = e - j;
    const float ps2 = __fmul_rn(2.0f, ps);
    double ll_sum = 0.0;
    for (int k = 0; k < K; k++) {
        const uint32_t *pm = Pm + (size_t)k * (ns + 1), *pt = Pt + (size_t)k * (ns + 1);
        const float nm = (float)(pm[e + 1] - pm[i]);
        const float nt = (float)(pt[e + 1] - pt[i]);
        if (nt == 0.0f) continue;
        const float p = __fdiv_rn(__fadd_rn(nm, ps), __fadd_rn(nt, ps2));            // segmentor.cpp:127
        float ll = 0.0f;
        if (p > 0.0f) ll = __fadd_rn(ll, __fmul_rn(nm, glibc_log2f(p)));            // :129-131 (float)
        if (p < 1.0f) {                                                              // :132-134 (double, rounded back to float)
            const double t = __dmul_rn((double)__fsub_rn(nt, nm), glibc_log2(__dsub_rn(1.0, (double)p)));
            ll = __double2float_rn(__dadd_rn((double)ll, t));
        }
        ll_sum = __dadd_rn(ll_sum, (double)ll);
    }
    cost[c] = ll_sum != 0.0 ? ll_sum : 0.0;                                          // :137
}

// ---- DP: one WARP per chunk --------------------------------------------------------------------------------------------
// The recurrence is a sequential chain over the chunk's sites; one step costs (shared-memory M read + 5 shuffle rounds), so
// the latency of everything else is taken off the chain: W / cell offsets of 32 consecutive steps sit in registers (one per
// lane, broadcast by shuffle), and the first 32 cost cells of step x+1 are loaded while step x is being reduced.  No block
// barrier anywhere; many chunks share an SM.
constexpr int DP_T = 32;
// (The argmax of a step by three REDUX.MAX warp reductions on order-preserving integer keys instead of the five shuffle rounds was
// measured on a B200: 31.2 ms against 32.4 ms per 60 000-site chunk -- the reduction is not what the chain waits for; not kept.)
__global__ void __launch_bounds__(DP_T) seg_dp_k(const Chunk *__restrict__ chunks, const uint32_t *__restrict__ W, const uint64_t *__restrict__ coff,
                                                  const double *__restrict__ cost, uint32_t ring, int32_t *__restrict__ Tb /* per site+chunk */,
                                                  const uint64_t *__restrict__ toff) {
    extern __shared__ double M[];                 // ring (power of two > max_cpg)
    const Chunk ch = chunks[blockIdx.x];
    int32_t *T = Tb + toff[blockIdx.x];           // n+1 entries
    const uint32_t mask = ring - 1, lane = threadIdx.x;
    if (lane == 0) { M[0] = 0.0; T[0] = 0; }
    __syncwarp();
    for (uint32_t xb = 0; xb < ch.n; xb += 32) {
        const bool have = xb + lane < ch.n;
        const uint32_t myW = have ? W[ch.start + xb + lane] : 0u;
        const unsigned long long myB = have ? coff[ch.start + xb + lane] : 0ull;
        uint32_t w = __shfl_sync(0xffffffffu, myW, 0);
        unsigned long long base = __shfl_sync(0xffffffffu, myB, 0);
        double c_cur = lane < w ? cost[base + lane] : 0.0;
        const uint32_t steps = min(32u, ch.n - xb);
        for (uint32_t s = 0; s < steps; s++) {
            const uint32_t x = xb + s;            // = i in the reference's loop (site index inside the chunk)
            uint32_t wn = 0; unsigned long long bn = 0; double c_next = 0.0;
            if (s + 1 < steps) {                  // software prefetch of the next step's first 32 cells
                wn = __shfl_sync(0xffffffffu, myW, s + 1); bn = __shfl_sync(0xffffffffu, myB, s + 1);
                if (lane < wn) c_next = cost[bn + lane];
            }
            double best = -INFINITY; uint32_t bj = 0;
            if (lane < w) { best = M[(x - lane) & mask] + c_cur; bj = lane; }
            for (uint32_t j = lane + 32; j < w; j += 32) {
                const double v = M[(x - j) & mask] + cost[base + j];
                if (v >= best) { best = v; bj = j; }                          // lowest k = largest j wins ties
            }
            {
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) {
                    const double ov = __shfl_xor_sync(0xffffffffu, best, d); const uint32_t oj = __shfl_xor_sync(0xffffffffu, bj, d);
                    if (ov > best || (ov == best && oj > bj)) { best = ov; bj = oj; }
                }
            }
            if (lane == 0) { M[(x + 1) & mask] = best; T[x + 1] = (int32_t)(x - bj); }
            __syncwarp();
            w = wn; base = bn; c_cur = c_next;
        }
    }
}

__global__ void seg_trace_k(const Chunk *__restrict__ chunks, uint32_t nchunks, const int32_t *__restrict__ Tb, const uint64_t *__restrict__ toff,
                            int32_t *__restrict__ borders, int32_t *__restrict__ nborders) {
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nchunks) return;
    const int32_t *T = Tb + toff[c];
    int32_t *out = borders + toff[c];               // same capacity n+1
    int32_t i = (int32_t)chunks[c].n, nb = 0;
    out[nb++] = i;
    while (i > 0) { i = max(0, T[i]); out[nb++] = i; }
    for (int32_t a = 0, b = nb - 1; a < b; a++, b--) { int32_t t = out[a]; out[a] = out[b]; out[b] = t; }
    nborders[c] = nb;
}

__global__ void seg_chunk_of_k(const Chunk *__restrict__ chunks, uint32_t nchunks, uint32_t *__restrict__ chunk_of) {
    uint32_t c = blockIdx.x;
    const Chunk ch = chunks[c];
    for (uint32_t s = threadIdx.x; s < ch.n; s += blockDim.x) chunk_of[ch.start + s] = c;
}

// cells of every chunk of the call (what seg_window_k will find, summed per chunk): lets the host pack waves by the real scratch
// need instead of the worst case n * max_cpg -- with max_bp limiting blocks to a few dozen sites that is ~40x more chunks
// (= concurrent DP warps) per wave.  One CTA per chunk; chunks are absolute here.
__global__ void __launch_bounds__(256) seg_chunk_cells_k(const uint32_t *__restrict__ dists, const Chunk *__restrict__ chunks, uint32_t max_cpg,
                                                          uint32_t max_bp, unsigned long long *__restrict__ cells) {
    const Chunk c = chunks[blockIdx.x];
    unsigned long long sum = 0;
    for (uint32_t k = threadIdx.x; k < c.n; k += blockDim.x) {
        const uint32_t e = c.start + k;
        uint32_t lo = c.start;
        if (e + 1 - c.start > max_cpg) lo = e + 1 - max_cpg;
        const uint32_t de = dists[e];
        uint32_t a = lo, b = e;
        while (a < b) { uint32_t m = (a + b) >> 1; if (de - dists[m] > max_bp) a = m + 1; else b = m; }
        sum += e - a + 1;
    }
    __shared__ unsigned long long ws[8];
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) { unsigned long long t = 0; for (int i = 0; i < 8; i++) t += ws[i]; cells[blockIdx.x] = t; }
}

__global__ void __launch_bounds__(256) glibc_log2_probe_k(const float *__restrict__ p, size_t n, float *__restrict__ l2f, double *__restrict__ l2) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (l2f) l2f[i] = glibc_log2f(p[i]);
    if (l2) l2[i] = glibc_log2(__dsub_rn(1.0, (double)p[i]));
}

}  // namespace

// numerics self-test hook: out_log2f[i] = log2f(p[i]); out_log2_1mp[i] = log2(1.0 - (double)p[i]), as glibc computes them
extern "C" int wgbs_glibc_log2_probe(wgbs_ctx *ctx, const float *p, size_t n, float *out_log2f, double *out_log2_1mp) {
    RC_TRY(wgbs_ctx_activate(ctx));
    Temps T(ctx);
    const void *dp; bool o; RC_TRY(to_device(ctx, p, n * 4, &dp, &o)); if (o) T.v.push_back((void *)dp);
    float *d1 = nullptr; double *d2 = nullptr;
    if (out_log2f) { if (is_device_ptr(out_log2f)) d1 = out_log2f; else RC_TRY(T.alloc(&d1, n)); }
    if (out_log2_1mp) { if (is_device_ptr(out_log2_1mp)) d2 = out_log2_1mp; else RC_TRY(T.alloc(&d2, n)); }
    if (n) { LAUNCH(ctx, glibc_log2_probe_k, grid_for(n, 256), 256, 0, (const float *)dp, n, d1, d2); LAUNCH_CHECK(); }
    if (out_log2f && d1 != out_log2f) RC_TRY(copy_any(ctx, out_log2f, d1, n * 4));
    if (out_log2_1mp && d2 != out_log2_1mp) RC_TRY(copy_any(ctx, out_log2_1mp, d2, n * 8));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int wgbs_segment(wgbs_ctx *ctx, const uint8_t *const *betas, int K, const uint32_t *dists, size_t nsites,
                            const wgbs_chunk *chunks, int nchunks, int max_cpg, uint32_t max_bp, float pseudo,
                            int32_t *borders, int32_t *nborders) {
    RC_TRY(wgbs_ctx_activate(ctx));
    if (K < 1 || !betas || !dists || !chunks || !borders || !nborders) return wgbs_set_err("wgbs_segment: bad argument");
    if (max_cpg < 1) return wgbs_set_err("wgbs_segment: max_cpg must be >= 1");
    if ((uint64_t)max_cpg * 255ull >= (1ull << 24)) return wgbs_set_err("wgbs_segment: max_cpg too large for exact float32 block sums (limit 65793)");
    uint32_t ring = 1; while (ring <= (uint32_t)max_cpg) ring <<= 1;
    const size_t smem = (size_t)ring * sizeof(double);
    if (smem > 200 * 1024) return wgbs_set_err("wgbs_segment: max_cpg %d needs a %zu KiB DP ring; the shared-memory DP supports max_cpg < 16384", max_cpg, smem >> 10);
    if (nsites >= 0x7fffffffull) return wgbs_set_err("wgbs_segment: too many sites in one call");
    std::vector<wgbs_chunk> hc(chunks, chunks + nchunks);
    uint64_t tot_out = 0;
    std::vector<uint64_t> out_off(nchunks + 1, 0);
    for (int c = 0; c < nchunks; c++) {
        if (hc[c].n == 0 || (uint64_t)hc[c].start + hc[c].n > nsites) return wgbs_set_err("wgbs_segment: chunk %d out of range", c);
        out_off[c] = tot_out; tot_out += (uint64_t)hc[c].n + 1;
    }
    out_off[nchunks] = tot_out;
    Temps T(ctx);
    // device-resident inputs: K beta arrays packed [K][nsites][2], dists
    // betas already resident in HBM are used in place; host arrays are copied once per call
    std::vector<const uint8_t *> hptr(K);
    for (int k = 0; k < K; k++) {
        const void *q; bool ob; RC_TRY(to_device(ctx, betas[k], nsites * 2, &q, &ob)); if (ob) T.v.push_back((void *)q);
        hptr[k] = (const uint8_t *)q;
    }
    const uint8_t **dbeta; RC_TRY(T.alloc(&dbeta, (size_t)K));
    RC_TRY(copy_any(ctx, dbeta, hptr.data(), (size_t)K * sizeof(void *)));
    const void *dd; bool od; RC_TRY(to_device(ctx, dists, nsites * 4, &dd, &od)); if (od) T.v.push_back((void *)dd);
    const uint32_t *ddists = (const uint32_t *)dd;
    uint32_t *bad = ctx->d_flags + 2;
    CUDA_TRY(cudaMemsetAsync(bad, 0, 4, ctx->stream));
    CUDA_TRY(cudaFuncSetAttribute(seg_dp_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));

    // waves of consecutive chunks bounded by a scratch budget
    const uint64_t CELL_BUDGET = 600ull << 20;          // cells (8 B each)  -> <= 4.7 GiB of cost
    const uint64_t PREFIX_BUDGET = 6ull << 30;          // bytes of running sums
    // the cells of a chunk are budgeted by what its windows really hold (one small kernel + one read-back per call), not by
    // n * min(max_cpg, n): with max_bp 2000 a site has ~25 admissible block lengths, not 1000, so ~400 chunks fit one wave instead of
    // 10 (measured on a B200: same borders, 64 chunks of K = 10 in one wave)
    std::vector<unsigned long long> exact_cells;
    if (nchunks > 1) {
        Chunk *dall; unsigned long long *dcells;
        RC_TRY(T.alloc(&dall, (size_t)nchunks)); RC_TRY(T.alloc(&dcells, (size_t)nchunks));
        RC_TRY(copy_any(ctx, dall, hc.data(), (size_t)nchunks * sizeof(wgbs_chunk)));
        LAUNCH(ctx, seg_chunk_cells_k, (unsigned)nchunks, 256, 0, ddists, dall, (uint32_t)max_cpg, max_bp, dcells);
        LAUNCH_CHECK();
        exact_cells.resize(nchunks);
        RC_TRY(copy_any(ctx, exact_cells.data(), dcells, (size_t)nchunks * 8));
    }
    int c0 = 0;
    while (c0 < nchunks) {
        // a wave must cover a contiguous site range: extend while chunks are adjacent/ascending and budgets hold
        int c1 = c0; uint64_t sites = 0, worst_cells = 0;
        uint32_t s_begin = hc[c0].start, s_end = hc[c0].start;
        while (c1 < nchunks) {
            if (c1 > c0 && hc[c1].start < s_end) break;                       // overlapping / unordered: new wave
            uint64_t ns_new = (uint64_t)hc[c1].start + hc[c1].n - s_begin;
            uint64_t cells_new = worst_cells + (exact_cells.empty() ? (uint64_t)hc[c1].n * (uint64_t)std::min<uint64_t>(max_cpg, hc[c1].n) : (uint64_t)exact_cells[c1]);
            if (c1 > c0 && (ns_new * 8ull * K > PREFIX_BUDGET || cells_new > CELL_BUDGET)) break;
            sites = ns_new; worst_cells = cells_new; s_end = hc[c1].start + hc[c1].n; c1++;
        }
        const uint32_t ns = (uint32_t)sites, nw = (uint32_t)(c1 - c0);
        Temps W(ctx);
        std::vector<wgbs_chunk> rel(nw); std::vector<uint64_t> toff(nw + 1, 0);
        for (uint32_t q = 0; q < nw; q++) { rel[q].start = hc[c0 + q].start - s_begin; rel[q].n = hc[c0 + q].n; toff[q + 1] = toff[q] + rel[q].n + 1; }
        Chunk *dch; uint64_t *dtoff; uint32_t *chunk_of, *Wd, *Pm, *Pt; uint64_t *coff; uint2 *tsum;
        RC_TRY(W.alloc(&dch, nw)); RC_TRY(W.alloc(&dtoff, nw + 1)); RC_TRY(W.alloc(&chunk_of, ns)); RC_TRY(W.alloc(&Wd, ns)); RC_TRY(W.alloc(&coff, (size_t)ns + 1));
        RC_TRY(copy_any(ctx, dch, rel.data(), nw * sizeof(wgbs_chunk))); RC_TRY(copy_any(ctx, dtoff, toff.data(), (nw + 1) * 8));
        const uint32_t ntiles = (ns + 1 + PF_TILE - 1) / PF_TILE;
        RC_TRY(W.alloc(&Pm, (size_t)K * (ns + 1))); RC_TRY(W.alloc(&Pt, (size_t)K * (ns + 1))); RC_TRY(W.alloc(&tsum, (size_t)K * ntiles));
        CUDA_TRY(cudaMemsetAsync(Wd, 0, (size_t)ns * 4, ctx->stream));
        CUDA_TRY(cudaMemsetAsync(chunk_of, 0xff, (size_t)ns * 4, ctx->stream));
        LAUNCH(ctx, seg_chunk_of_k, nw, 256, 0, dch, nw, chunk_of);
        dim3 pg(ntiles, K);
        LAUNCH(ctx, seg_prefix_sums_k, pg, PF_T, 0, dbeta, s_begin, ns, ntiles, tsum, bad);
        LAUNCH(ctx, seg_prefix_tiles_k, K, 256, 0, tsum, ntiles);
        LAUNCH(ctx, seg_prefix_write_k, pg, PF_T, 0, dbeta, s_begin, ns, ntiles, tsum, Pm, Pt);
        LAUNCH(ctx, seg_window_k, grid_for(ns, 256), 256, 0, ddists + s_begin, chunk_of, dch, ns, (uint32_t)max_cpg, max_bp, Wd, bad);
        RC_TRY(scan_u32_u64(ctx, Wd, coff, ns));
        uint64_t ncells = 0; uint32_t hbad = 0;
        CUDA_TRY(cudaMemcpyAsync(&ncells, coff + ns, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(&hbad, bad, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        if (hbad & 1) return wgbs_set_err("invalid data: a beta file has meth > cover (segmentor.cpp:181-187)");
        if (hbad & 2) return wgbs_set_err("wgbs_segment: CpG loci (dists) must be non-decreasing inside a chunk");
        double *cost; int32_t *Tb, *dbord, *dnb;
        RC_TRY(W.alloc(&cost, ncells)); RC_TRY(W.alloc(&Tb, toff[nw])); RC_TRY(W.alloc(&dbord, toff[nw])); RC_TRY(W.alloc(&dnb, nw));
        if (ncells) {
            uint64_t g = (ncells + 255) / 256;
            if (g > 0x7fffffffull) return wgbs_set_err("wgbs_segment: wave too large");
            LAUNCH(ctx, seg_cost_k, (unsigned)g, 256, 0, Pm, Pt, ns, K, coff, ncells, pseudo, cost);
        }
        LAUNCH(ctx, seg_dp_k, nw, DP_T, smem, dch, Wd, coff, cost, ring, Tb, dtoff);
        LAUNCH(ctx, seg_trace_k, grid_for(nw, 64), 64, 0, dch, nw, Tb, dtoff, dbord, dnb);
        LAUNCH_CHECK();
        // borders: the caller's buffer is laid out like ours (n_c + 1 slots per chunk, chunk order)
        RC_TRY(copy_any(ctx, borders + out_off[c0], dbord, toff[nw] * 4));
        RC_TRY(copy_any(ctx, nborders + c0, dnb, nw * 4));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        c0 = c1;
    }
    return 0;
}
