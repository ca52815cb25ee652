// pat.cu -- pat text -> device records, pat2beta (+trim), homog.
//
// Reference behaviour restated here (paths relative to the reference's src/):
//   pat2beta/stdin2beta.cpp:59-93   Pat2Beta::proc_line      -> pat2beta_k
//   python/utils_wgbs.py:277-290    trim_to_uint8            -> trim_k
//   homog/homog.cpp:154-260         update_m2 / proc_line    -> homog_k
#include <stdlib.h>

#include "common.cuh"
#include "pats.cuh"
#include "lines.cuh"


namespace {

// ------------------------------------------------------------------------------------------------------------------
// one thread per pat line: "chr \t idx \t pattern \t count [\t ...]"
// ------------------------------------------------------------------------------------------------------------------
// std::stoi semantics: optional blanks, sign, >=1 digit; returns false (== throw) otherwise / on int overflow
__device__ __forceinline__ bool parse_int_field(const char *__restrict__ t, uint32_t s, uint32_t e, int32_t *out) {
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

__global__ void __launch_bounds__(256) pat_lines_k(const char *__restrict__ text, uint32_t n, const uint32_t *__restrict__ nlpos,
                                                    uint32_t n_nl, uint32_t n_lines, uint32_t *__restrict__ idx,
                                                    uint32_t *__restrict__ len, uint32_t *__restrict__ count,
                                                    uint32_t *__restrict__ pstart, uint32_t *__restrict__ words,
                                                    uint32_t *__restrict__ err) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_lines) return;
    uint32_t s = i == 0 ? 0 : nlpos[i - 1] + 1;
    uint32_t e = i < n_nl ? nlpos[i] : n;
    uint32_t o_idx = 0, o_len = 0, o_cnt = 0, o_ps = s;
    if (e > s) {
        uint32_t tab[4]; int nt = 0;
        for (uint32_t p = s; p < e && nt < 4; p++) if (text[p] == '\t') tab[nt++] = p;
        if (nt < 3) { atomicOr(err, 1u); }
        else {
            uint32_t cend = nt >= 4 ? tab[3] : e;
            int32_t vi, vc;
            if (!parse_int_field(text, tab[0] + 1, tab[1], &vi) || !parse_int_field(text, tab[2] + 1, cend, &vc)) atomicOr(err, 2u);
            else { o_idx = (uint32_t)vi; o_cnt = (uint32_t)vc; o_len = tab[2] - tab[1] - 1; o_ps = tab[1] + 1; }
        }
    }
    idx[i] = o_idx; len[i] = o_len; count[i] = o_cnt; pstart[i] = o_ps; words[i] = (o_len + 15) >> 4;
}

__global__ void __launch_bounds__(256) pat_pack_k(const char *__restrict__ text, uint32_t n_rec, const uint32_t *__restrict__ len,
                                                   const uint32_t *__restrict__ pstart, const uint32_t *__restrict__ off,
                                                   uint32_t *__restrict__ pool) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rec) return;
    uint32_t L = len[i], p = pstart[i], o = off[i];
    for (uint32_t b = 0; b < L; b += 16) {
        uint32_t w = 0, m = min(16u, L - b);
        for (uint32_t k = 0; k < m; k++) w |= sym_code(text[p + b + k]) << (30 - 2 * k);
        pool[o++] = w;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// (A two-pass tile parser -- 16 KiB tiles parked in shared memory by LDG.128 or one TMA bulk copy, newline / tab masks from SIMD
// byte compares -- was measured on a B200 against the four kernels above: 0.77 / 0.76 ms against 0.51 ms for 160 MB of pat text.  Not kept.)


// ------------------------------------------------------------------------------------------------------------------
// pat2beta: CTA-local shared-memory window over consecutive (idx-sorted) records, flushed with one red.add per
// touched site; symbols falling outside the window (unsorted input, sparse coverage) go straight to global.
// ------------------------------------------------------------------------------------------------------------------
constexpr int P2B_T = 256, P2B_RPT = 4, P2B_REC = P2B_T * P2B_RPT, P2B_W = 4096;

__global__ void __launch_bounds__(P2B_T) pat2beta_k(PatsView P, int32_t start, int32_t nsites, int32_t *__restrict__ mc) {
    __shared__ int32_t sm[2 * P2B_W];  // interleaved (meth, cover) like the output
    const size_t r0 = (size_t)blockIdx.x * P2B_REC;
    for (int i = threadIdx.x; i < 2 * P2B_W; i += P2B_T) sm[i] = 0;
    // window base: site offset of the first record in this CTA
    const int64_t base = (int64_t)(int32_t)P.idx[r0] - start;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < P2B_RPT; j++) {
        size_t r = r0 + (size_t)j * P2B_T + threadIdx.x;
        if (r >= P.n) break;
        const uint32_t L = P.len[r];
        const int32_t cnt = (int32_t)P.count[r];
        int64_t k = (int64_t)(int32_t)P.idx[r] - start;   // site offset of symbol 0
        if (L == 0 || k >= nsites || k + (int64_t)L <= 0) continue;   // stdin2beta.cpp:72-75
        const uint32_t *wp = P.pool + P.off[r];
        for (uint32_t b = 0; b < L; b += 16, k += 16) {
            uint32_t w = *wp++;
            while (w) {
                int lz = __clz(w) >> 1;                     // symbol position of the next non-'.' symbol
                uint32_t code = (w >> (30 - 2 * lz)) & 3u;
                w &= ~(3u << (30 - 2 * lz));
                int64_t site = k + lz;
                if (site < 0 || site >= nsites) continue;
                int64_t wofs = site - base;
                int add_m = (code != SYM_T) ? cnt : 0;      // 'C' or 'H'
                if (wofs >= 0 && wofs < P2B_W) {
                    atomicAdd(&sm[2 * wofs + 1], cnt);
                    if (add_m) atomicAdd(&sm[2 * wofs], add_m);
                } else {
                    atomicAdd(&mc[2 * site + 1], cnt);
                    if (add_m) atomicAdd(&mc[2 * site], add_m);
                }
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * P2B_W; i += P2B_T) {
        int32_t v = sm[i];
        if (v) { int64_t g = 2 * base + i; atomicAdd(&mc[g], v); }
    }
}

template <typename OutT>
__global__ void __launch_bounds__(256) trim_k(const int2 *__restrict__ mc, size_t n, int32_t maxv, OutT *__restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int2 v = mc[i];
    int64_t m = v.x, c = v.y;
    if (c > maxv) {
        // utils_wgbs.py:286-287: float64 divide, then multiply, then truncate toward zero
        double q = __dmul_rn(__ddiv_rn((double)m, (double)c), (double)maxv);
        m = (int64_t)q;
        c = maxv;
    }
    out[2 * i] = (OutT)m;
    out[2 * i + 1] = (OutT)c;
}

// ------------------------------------------------------------------------------------------------------------------
// homog
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) homog_cut_k(PatsView P, int32_t last_end, unsigned long long *cut) {
    size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= P.n) return;
    // homog.cpp:218-223: the first record starting at/after the LAST block's end stops the run
    if ((int32_t)P.idx[r] >= last_end) atomicMin(cut, (unsigned long long)r);
}

// count C|H and T among symbols [a, b) of a record
__device__ __forceinline__ void count_ct(const uint32_t *__restrict__ wp, uint32_t a, uint32_t b, int *nC, int *nT) {
    int c = 0, t = 0;
    for (uint32_t wi = a >> 4; wi <= (b - 1) >> 4; wi++) {
        uint32_t w = wp[wi];
        uint32_t lo_s = wi == (a >> 4) ? (a & 15) : 0, hi_s = wi == ((b - 1) >> 4) ? ((b - 1) & 15) + 1 : 16;  // symbols [lo_s, hi_s)
        uint32_t mask = (hi_s - lo_s == 16) ? 0xffffffffu : (((1u << (2 * (hi_s - lo_s))) - 1u) << (32 - 2 * hi_s));
        w &= mask;
        uint32_t lo = w & 0x55555555u, hi = (w >> 1) & 0x55555555u;
        t += __popc(hi & lo);
        c += __popc(hi ^ lo);
    }
    *nC = c; *nT = t;
}

__global__ void __launch_bounds__(256) homog_k(PatsView P, const int32_t *__restrict__ bs, const int32_t *__restrict__ be,
                                                const int32_t *__restrict__ pmax, int32_t nb, const float *__restrict__ range,
                                                int nbins, int min_cpgs, int inclusive, const unsigned long long *__restrict__ cut,
                                                int32_t *__restrict__ out) {
    size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= P.n || r >= *cut) return;
    const int32_t rs = (int32_t)P.idx[r];
    const int32_t L = (int32_t)P.len[r];
    const int32_t re = rs + L - 1;
    // first block with start > read_end
    int lo = 0, hi = nb;
    while (lo < hi) { int m = (lo + hi) >> 1; if (bs[m] <= re) lo = m + 1; else hi = m; }
    const int ub = lo;
    // first block whose running max end exceeds read_start (blocks before it all end at/before the read)
    lo = 0; hi = ub;
    while (lo < hi) { int m = (lo + hi) >> 1; if (pmax[m] <= rs) lo = m + 1; else hi = m; }
    const uint32_t *wp = P.pool + P.off[r];
    const int32_t cnt = (int32_t)P.count[r];
    for (int bi = lo; bi < ub; bi++) {
        int32_t os = max(rs, bs[bi]), oe = min(rs + L, be[bi]);
        if (os >= oe) continue;
        int nC, nT;
        if (inclusive) { if (L < min_cpgs) continue; count_ct(wp, 0, (uint32_t)L, &nC, &nT); }
        else { if (oe - os < min_cpgs) continue; count_ct(wp, (uint32_t)(os - rs), (uint32_t)(oe - rs), &nC, &nT); }
        if (nC + nT < min_cpgs) continue;
        float meth = __fdiv_rn((float)nC, (float)(nC + nT));
        if (meth < range[0]) continue;
        int bin = 0;
        for (bin = 0; bin < nbins; bin++) if ((meth >= range[bin]) && (meth < range[bin + 1])) break;
        if (bin == nbins) bin--;
        atomicAdd(&out[(size_t)bi * nbins + bin], cnt);
    }
}

__global__ void set_u64_k(unsigned long long *p, unsigned long long v) { *p = v; }

}  // namespace

// ==================================================================================================================
// C ABI
// ==================================================================================================================

extern "C" int wgbs_pats_from_text(wgbs_ctx *ctx, const char *text, size_t nbytes, wgbs_pats **out) {
    RC_TRY(wgbs_ctx_activate(ctx));
    if (!out) return wgbs_set_err("wgbs_pats_from_text: out is null");
    *out = nullptr;
    if (nbytes >= 0xfffffff0ull) return wgbs_set_err("wgbs_pats_from_text: text must be < 4 GiB per call (got %zu); split on line boundaries", nbytes);
    Temps T(ctx);
    const void *dtext_v = nullptr; bool owned = false;
    RC_TRY(to_device(ctx, text, nbytes, &dtext_v, &owned));
    const char *dtext = (const char *)dtext_v;
    if (owned) T.v.push_back((void *)dtext);
    const uint32_t n = (uint32_t)nbytes;
    // 1. newline positions
    uint32_t *nlpos = nullptr, n_nl = 0, n_lines = 0;
    RC_TRY(find_lines(ctx, dtext, nbytes, T, &nlpos, &n_nl, &n_lines));
    // 2. per-line fields
    wgbs_pats *P = new wgbs_pats();
    P->n = n_lines;
    uint32_t *pstart, *words;
    int rc = 0;
    if ((rc = dalloc(ctx, &P->idx, n_lines)) < 0 || (rc = dalloc(ctx, &P->len, n_lines)) < 0 || (rc = dalloc(ctx, &P->count, n_lines)) < 0 ||
        (rc = dalloc(ctx, &P->off, (size_t)n_lines + 1)) < 0) { wgbs_pats_free(ctx, P); return rc; }
    if ((rc = T.alloc(&pstart, n_lines)) < 0 || (rc = T.alloc(&words, n_lines)) < 0) { wgbs_pats_free(ctx, P); return rc; }
    uint32_t *err = ctx->d_flags;
    CUDA_TRY(cudaMemsetAsync(err, 0, 4, ctx->stream));
    if (n_lines) LAUNCH(ctx, pat_lines_k, grid_for(n_lines, 256), 256, 0, dtext, n, nlpos, n_nl, n_lines, P->idx, P->len, P->count, pstart, words, err);
    if ((rc = scan_u32_u32(ctx, words, P->off, n_lines)) < 0) { wgbs_pats_free(ctx, P); return rc; }
    uint32_t herr = 0, total = 0;
    CUDA_TRY(cudaMemcpyAsync(&herr, err, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(&total, P->off + n_lines, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (herr) {
        wgbs_pats_free(ctx, P);
        return wgbs_set_err(herr & 1 ? "failed parsing pat: too few columns in file" : "failed parsing pat: non-numeric CpG index or count");
    }
    P->pool_words = total;
    if ((rc = dalloc(ctx, &P->pool, total)) < 0) { wgbs_pats_free(ctx, P); return rc; }
    if (n_lines) LAUNCH(ctx, pat_pack_k, grid_for(n_lines, 256), 256, 0, dtext, n_lines, P->len, pstart, P->off, P->pool);
    LAUNCH_CHECK();
    *out = P;
    return 0;
}

extern "C" int wgbs_pats_count(const wgbs_pats *P, uint64_t *n_records, uint64_t *n_pool_words) {
    if (!P) return wgbs_set_err("null pats");
    if (n_records) *n_records = P->n;
    if (n_pool_words) *n_pool_words = P->pool_words;
    return 0;
}

extern "C" int wgbs_pats_download(wgbs_ctx *ctx, const wgbs_pats *P, uint32_t *idx, uint32_t *len, uint32_t *count, uint32_t *off, uint32_t *pool) {
    RC_TRY(wgbs_ctx_activate(ctx));
    if (!P) return wgbs_set_err("null pats");
    if (idx) RC_TRY(copy_any(ctx, idx, P->idx, P->n * 4));
    if (len) RC_TRY(copy_any(ctx, len, P->len, P->n * 4));
    if (count) RC_TRY(copy_any(ctx, count, P->count, P->n * 4));
    if (off) RC_TRY(copy_any(ctx, off, P->off, P->n * 4));
    if (pool) RC_TRY(copy_any(ctx, pool, P->pool, P->pool_words * 4));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" void wgbs_pats_free(wgbs_ctx *ctx, wgbs_pats *P) {
    if (!P || !ctx) return;
    cudaSetDevice(ctx->device);
    dfree(ctx, P->idx); dfree(ctx, P->len); dfree(ctx, P->count); dfree(ctx, P->off); dfree(ctx, P->pool);
    dfree(ctx, P->name_off); dfree(ctx, P->name_len); dfree(ctx, P->names);
    delete P;
}

extern "C" int wgbs_pat2beta(wgbs_ctx *ctx, const wgbs_pats *P, uint32_t start, uint32_t end, int32_t *meth_cov, int zero_first) {
    RC_TRY(wgbs_ctx_activate(ctx));
    if (!P) return wgbs_set_err("null pats");
    if (end < start) return wgbs_set_err("wgbs_pat2beta: end < start");
    if (!is_device_ptr(meth_cov)) return wgbs_set_err("wgbs_pat2beta: meth_cov must be a device pointer");
    const size_t ns = (size_t)end - start;
    if (zero_first) CUDA_TRY(cudaMemsetAsync(meth_cov, 0, ns * 2 * sizeof(int32_t), ctx->stream));
    if (P->n && ns) {
        LAUNCH(ctx, pat2beta_k, grid_for(P->n, P2B_T, P2B_RPT), P2B_T, 0, view_of(P), (int32_t)start, (int32_t)ns, meth_cov);
        LAUNCH_CHECK();
    }
    return 0;
}

extern "C" int wgbs_trim(wgbs_ctx *ctx, const int32_t *meth_cov, size_t n, int nbits, void *out) {
    RC_TRY(wgbs_ctx_activate(ctx));
    if (nbits != 8 && nbits != 16) return wgbs_set_err("wgbs_trim: nbits must be 8 or 16");
    Temps T(ctx);
    const void *dmc = nullptr; bool owned = false;
    RC_TRY(to_device(ctx, meth_cov, n * 8, &dmc, &owned));
    if (owned) T.v.push_back((void *)dmc);
    const size_t obytes = n * 2 * (nbits / 8);
    void *dout = out;
    if (!is_device_ptr(out)) RC_TRY(T.alloc((char **)&dout, obytes));
    if (n) {
        if (nbits == 8) LAUNCH(ctx, trim_k<uint8_t>, grid_for(n, 256), 256, 0, (const int2 *)dmc, n, 255, (uint8_t *)dout);
        else LAUNCH(ctx, trim_k<uint16_t>, grid_for(n, 256), 256, 0, (const int2 *)dmc, n, 65535, (uint16_t *)dout);
        LAUNCH_CHECK();
    }
    if (dout != out) RC_TRY(copy_any(ctx, out, dout, obytes));
    return 0;
}

extern "C" int wgbs_pat2beta_text(wgbs_ctx *ctx, const char *text, size_t nbytes, uint32_t start, uint32_t end, int nbits,
                                   void *beta_out, int32_t *meth_cov_out) {
    RC_TRY(wgbs_ctx_activate(ctx));
    if (end < start) return wgbs_set_err("wgbs_pat2beta_text: end < start");
    wgbs_pats *P = nullptr;
    RC_TRY(wgbs_pats_from_text(ctx, text, nbytes, &P));
    Temps T(ctx);
    const size_t ns = (size_t)end - start;
    int32_t *mc = nullptr;
    int rc = T.alloc(&mc, ns * 2);
    if (rc == 0) rc = wgbs_pat2beta(ctx, P, start, end, mc, 1);
    if (rc == 0 && beta_out) rc = wgbs_trim(ctx, mc, ns, nbits, beta_out);
    if (rc == 0 && meth_cov_out) rc = copy_any(ctx, meth_cov_out, mc, ns * 8);
    wgbs_pats_free(ctx, P);
    if (rc == 0) CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return rc;
}

extern "C" int wgbs_homog(wgbs_ctx *ctx, const wgbs_pats *P, const int32_t *bstart, const int32_t *bend, size_t nblocks,
                          const float *range, int nbins, int min_cpgs, int inclusive, int32_t *out) {
    RC_TRY(wgbs_ctx_activate(ctx));
    if (!P) return wgbs_set_err("null pats");
    if (nbins < 1 || nbins > 64) return wgbs_set_err("wgbs_homog: nbins must be in [1,64]");
    if (nblocks > 0x7fffffffull) return wgbs_set_err("wgbs_homog: too many blocks");
    Temps T(ctx);
    // host copies of the block borders (validation + running max of ends; O(B) host logic, not the hot path)
    std::vector<int32_t> hs(nblocks), he(nblocks), hp(nblocks);
    RC_TRY(copy_any(ctx, hs.data(), bstart, nblocks * 4));
    RC_TRY(copy_any(ctx, he.data(), bend, nblocks * 4));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    int32_t run = INT32_MIN;
    for (size_t i = 0; i < nblocks; i++) {
        if (he[i] <= hs[i]) return wgbs_set_err("Invalid block: endCpG <= startCpG");       // homog.cpp:92-95
        if (hs[i] < 1) return wgbs_set_err("Invalid block: startCpG < 1");                  // homog.cpp:96-98
        if (i && hs[i] < hs[i - 1]) return wgbs_set_err("wgbs_homog: blocks must be sorted by startCpG (apply --sort_blocks on the host)");
        run = he[i] > run ? he[i] : run; hp[i] = run;
    }
    const void *dbs, *dbe, *drange; bool o1, o2, o3;
    RC_TRY(to_device(ctx, bstart, nblocks * 4, &dbs, &o1)); if (o1) T.v.push_back((void *)dbs);
    RC_TRY(to_device(ctx, bend, nblocks * 4, &dbe, &o2)); if (o2) T.v.push_back((void *)dbe);
    RC_TRY(to_device(ctx, range, (size_t)(nbins + 1) * 4, &drange, &o3)); if (o3) T.v.push_back((void *)drange);
    int32_t *dpm; RC_TRY(T.alloc(&dpm, nblocks));
    RC_TRY(copy_any(ctx, dpm, hp.data(), nblocks * 4));
    const size_t ob = nblocks * (size_t)nbins * 4;
    int32_t *dout = out;
    if (!is_device_ptr(out)) RC_TRY(T.alloc(&dout, nblocks * (size_t)nbins));
    CUDA_TRY(cudaMemsetAsync(dout, 0, ob ? ob : 4, ctx->stream));
    unsigned long long *cut; RC_TRY(T.alloc(&cut, 1));
    LAUNCH(ctx, set_u64_k, 1, 1, 0, cut, ~0ull);
    if (P->n && nblocks) {
        LAUNCH(ctx, homog_cut_k, grid_for(P->n, 256), 256, 0, view_of(P), he[nblocks - 1], cut);
        LAUNCH(ctx, homog_k, grid_for(P->n, 256), 256, 0, view_of(P), (const int32_t *)dbs, (const int32_t *)dbe, (const int32_t *)dpm,
               (int32_t)nblocks, (const float *)drange, nbins, min_cpgs, inclusive, cut, dout);
    }
    LAUNCH_CHECK();
    if (dout != out) RC_TRY(copy_any(ctx, out, dout, ob));
    return 0;
}
