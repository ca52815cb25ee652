// view.cu -- the two direct consumers of pat / beta files that sit next to the hot path:
//   * cview      (reference src/cview/cview.cpp:87-167): select / clip pat records by blocks, --strict --strip --no_gaps --min_cpgs
//   * beta_to_blocks (reference src/python/beta_to_blocks.py:101-126): per-block sums of a beta file (np.add.reduceat) + trim
#include <vector>

#include "common.cuh"
#include "cview_core.cuh"
#include "pats.cuh"

namespace {

struct CountEmit {
    uint32_t pieces = 0, words = 0;
    __device__ __forceinline__ void operator()(int32_t, uint32_t, uint32_t len) { pieces++; words += (len + 15) >> 4; }
};
__global__ void __launch_bounds__(256) cview_count_k(PatsView P, CviewParams prm, uint32_t *__restrict__ npieces, uint32_t *__restrict__ nwords) {
    const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= P.n) return;
    CountEmit ce;
    cview_record(prm, (int32_t)P.idx[r], P.len[r], P.pool + P.off[r], ce);
    npieces[r] = ce.pieces; nwords[r] = ce.words;
}

// 16 symbols starting at symbol a of a record of nw words (zero = '.' beyond the record)
__device__ __forceinline__ uint32_t take16(const uint32_t *__restrict__ wp, uint32_t nw, uint32_t a) {
    const uint32_t j = a >> 4, sh = 2 * (a & 15);
    const uint32_t hi = j < nw ? wp[j] : 0u, lo = (j + 1 < nw) ? wp[j + 1] : 0u;
    return sh ? ((hi << sh) | (lo >> (32 - sh))) : hi;
}
struct WriteEmit {
    const uint32_t *wp; uint32_t nw, cnt;
    uint32_t piece, word;                                            // running output positions
    uint32_t *o_idx, *o_len, *o_cnt, *o_off, *o_pool;
    __device__ __forceinline__ void operator()(int32_t start, uint32_t a, uint32_t len) {
        o_idx[piece] = (uint32_t)start; o_len[piece] = len; o_cnt[piece] = cnt; o_off[piece] = word;
        for (uint32_t k = 0; k < len; k += 16) {
            uint32_t w = take16(wp, nw, a + k);
            const uint32_t m = len - k;
            if (m < 16) w &= ~0u << (32 - 2 * m);                    // zero padding behind the last symbol (collapse relies on it)
            o_pool[word++] = w;
        }
        piece++;
    }
};
__global__ void __launch_bounds__(256) cview_write_k(PatsView P, CviewParams prm, const uint32_t *__restrict__ piece_off, const uint32_t *__restrict__ word_off,
                                                      uint32_t *__restrict__ o_idx, uint32_t *__restrict__ o_len, uint32_t *__restrict__ o_cnt,
                                                      uint32_t *__restrict__ o_off, uint32_t *__restrict__ o_pool) {
    const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= P.n) return;
    if (piece_off[r + 1] == piece_off[r]) return;
    WriteEmit we{P.pool + P.off[r], (P.len[r] + 15) >> 4, P.count[r], piece_off[r], word_off[r], o_idx, o_len, o_cnt, o_off, o_pool};
    cview_record(prm, (int32_t)P.idx[r], P.len[r], P.pool + P.off[r], we);
}
__global__ void set_last_off_k(uint32_t *o_off, uint32_t n, const uint32_t *word_off_total) { o_off[n] = *word_off_total; }

// ---- beta_to_blocks -------------------------------------------------------------------------------------------------
// 8 lanes per block: the lanes stride over the block's sites (2 values per site), 64-bit partial sums, shuffle reduce.
template <typename InT>
__global__ void __launch_bounds__(256) beta_blocks_k(const InT *__restrict__ beta, size_t nsites, const int32_t *__restrict__ bs, const int32_t *__restrict__ be,
                                                      size_t nblocks, long long *__restrict__ sums) {
    const size_t g = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const unsigned lane = threadIdx.x & 7;
    unsigned long long m = 0, c = 0;
    if (g < nblocks) {
        // rows startCpG-1 .. endCpG-2 (numpy slice data[startCpG-1 : endCpG-1], clamped to the file like a slice is)
        long long a = (long long)bs[g] - 1, b = (long long)be[g] - 1;
        if (a < 0) a = 0;
        if (b > (long long)nsites) b = (long long)nsites;
        for (long long i = a + lane; i < b; i += 8) { m += beta[2 * i]; c += beta[2 * i + 1]; }
    }
    for (int d = 4; d >= 1; d >>= 1) { m += __shfl_xor_sync(0xffffffffu, m, d, 8); c += __shfl_xor_sync(0xffffffffu, c, d, 8); }
    if (g < nblocks && lane == 0) { sums[2 * g] = (long long)m; sums[2 * g + 1] = (long long)c; }
}
template <typename OutT>
__global__ void __launch_bounds__(256) trim64_k(const long long *__restrict__ mc, size_t n, long long maxv, OutT *__restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    long long m = mc[2 * i], c = mc[2 * i + 1];
    if (c > maxv) {                                                  // utils_wgbs.py:286-287: float64 divide, multiply, truncate
        const double q = __dmul_rn(__ddiv_rn((double)m, (double)c), (double)maxv);
        m = (long long)q; c = maxv;
    }
    out[2 * i] = (OutT)m; out[2 * i + 1] = (OutT)c;
}

}  // namespace

extern "C" int wgbs_cview(wgbs_ctx *ctx, const wgbs_pats *in, const int32_t *bstart, const int32_t *bend, size_t nblocks,
                          const int32_t *pre_lo, const int32_t *pre_hi, size_t npre, int strict, int strip, int no_gaps, int min_cpgs,
                          wgbs_pats **out) {
    RC_TRY(wgbs_ctx_activate(ctx));
    if (!in || !out) return wgbs_set_err("wgbs_cview: null argument");
    *out = nullptr;
    if (in->names) return wgbs_set_err("wgbs_cview: records with read names (--long) are not supported");
    if (nblocks == 0) return wgbs_set_err("Error while loading blocks. 0 blocks found.");                    // cview.cpp:80-82
    if (nblocks > 0x7ffffff0ull || npre > 0x7ffffff0ull || in->n >= 0xfffffff0ull) return wgbs_set_err("wgbs_cview: too many blocks / records");
    if ((!bstart || !bend) || (npre && (!pre_lo || !pre_hi))) return wgbs_set_err("wgbs_cview: null block / range array");
    Temps T(ctx);
    // host copies for validation and the running maximum of ends (O(blocks) host logic)
    std::vector<int32_t> hs(nblocks), he(nblocks), hp(nblocks), pl(npre), ph(npre);
    RC_TRY(copy_any(ctx, hs.data(), bstart, nblocks * 4)); RC_TRY(copy_any(ctx, he.data(), bend, nblocks * 4));
    if (npre) { RC_TRY(copy_any(ctx, pl.data(), pre_lo, npre * 4)); RC_TRY(copy_any(ctx, ph.data(), pre_hi, npre * 4)); }
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    int32_t run = INT32_MIN;
    for (size_t i = 0; i < nblocks; i++) {
        if (he[i] <= hs[i]) return wgbs_set_err("Invalid block: endCpG <= startCpG");                         // cview.cpp:68-71
        if (hs[i] < 1) return wgbs_set_err("Invalid block: startCpG < 1");                                    // cview.cpp:72-73
        if (i && hs[i] < hs[i - 1]) return wgbs_set_err("wgbs_cview: blocks must be sorted by startCpG (cview sorts them with `sort -k1,1n`)");
        if (strict && i && hs[i] < he[i - 1]) return wgbs_set_err("wgbs_cview: --strict needs non-overlapping blocks (the reference aborts on them)");
        run = he[i] > run ? he[i] : run; hp[i] = run;
    }
    for (size_t i = 0; i < npre; i++)
        if (ph[i] < pl[i] || (i && pl[i] <= ph[i - 1])) return wgbs_set_err("wgbs_cview: pre-selection ranges must be sorted and disjoint");
    int32_t *dbs, *dbe, *dpm, *dpl = nullptr, *dph = nullptr;
    RC_TRY(T.alloc(&dbs, nblocks)); RC_TRY(T.alloc(&dbe, nblocks)); RC_TRY(T.alloc(&dpm, nblocks));
    RC_TRY(copy_any(ctx, dbs, hs.data(), nblocks * 4)); RC_TRY(copy_any(ctx, dbe, he.data(), nblocks * 4)); RC_TRY(copy_any(ctx, dpm, hp.data(), nblocks * 4));
    if (npre) {
        RC_TRY(T.alloc(&dpl, npre)); RC_TRY(T.alloc(&dph, npre));
        RC_TRY(copy_any(ctx, dpl, pl.data(), npre * 4)); RC_TRY(copy_any(ctx, dph, ph.data(), npre * 4));
    }
    CviewParams prm{dbs, dbe, dpm, (int32_t)nblocks, dpl, dph, (int32_t)npre, strict, strip, no_gaps, min_cpgs};
    const size_t n = in->n;
    uint32_t *np_, *nw_, *poff, *woff;
    RC_TRY(T.alloc(&np_, n)); RC_TRY(T.alloc(&nw_, n)); RC_TRY(T.alloc(&poff, n + 1)); RC_TRY(T.alloc(&woff, n + 1));
    const PatsView pv = view_of(in);
    if (n) LAUNCH(ctx, cview_count_k, grid_for(n, 256), 256, 0, pv, prm, np_, nw_);
    RC_TRY(scan_u32_u32(ctx, np_, poff, n)); RC_TRY(scan_u32_u32(ctx, nw_, woff, n));
    uint32_t tot[2] = {0, 0};
    CUDA_TRY(cudaMemcpyAsync(&tot[0], poff + n, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(&tot[1], woff + n, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    wgbs_pats *O = new wgbs_pats();
    O->n = tot[0]; O->pool_words = tot[1];
    int rc = 0;
    if ((rc = dalloc(ctx, &O->idx, O->n)) < 0 || (rc = dalloc(ctx, &O->len, O->n)) < 0 || (rc = dalloc(ctx, &O->count, O->n)) < 0 ||
        (rc = dalloc(ctx, &O->off, O->n + 1)) < 0 || (rc = dalloc(ctx, &O->pool, O->pool_words)) < 0) { wgbs_pats_free(ctx, O); return rc; }
    if (n) LAUNCH(ctx, cview_write_k, grid_for(n, 256), 256, 0, pv, prm, poff, woff, O->idx, O->len, O->count, O->off, O->pool);
    LAUNCH(ctx, set_last_off_k, 1, 1, 0, O->off, (uint32_t)O->n, woff + n);
    LAUNCH_CHECK();
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));                    // the temporaries above are read by the kernels
    *out = O;
    return 0;
}

extern "C" int wgbs_beta_to_blocks(wgbs_ctx *ctx, const void *beta, int in_bits, size_t nsites, const int32_t *bstart, const int32_t *bend,
                                   size_t nblocks, int out_bits, void *out_trimmed, int64_t *out_sums) {
    RC_TRY(wgbs_ctx_activate(ctx));
    if (in_bits != 8 && in_bits != 16) return wgbs_set_err("wgbs_beta_to_blocks: in_bits must be 8 (.beta/.bin) or 16 (.lbeta)");
    if (out_trimmed && out_bits != 8 && out_bits != 16) return wgbs_set_err("wgbs_beta_to_blocks: out_bits must be 8 or 16");
    if ((nsites && !beta) || (nblocks && (!bstart || !bend))) return wgbs_set_err("wgbs_beta_to_blocks: null argument");
    if (nblocks == 0) return 0;
    Temps T(ctx);
    const void *dbeta = nullptr, *dbs = nullptr, *dbe = nullptr; bool own = false;
    RC_TRY(to_device(ctx, beta, nsites * 2 * (in_bits / 8), &dbeta, &own)); if (own) T.v.push_back((void *)dbeta);
    RC_TRY(to_device(ctx, bstart, nblocks * 4, &dbs, &own)); if (own) T.v.push_back((void *)dbs);
    RC_TRY(to_device(ctx, bend, nblocks * 4, &dbe, &own)); if (own) T.v.push_back((void *)dbe);
    long long *dsum = (long long *)out_sums;
    if (!out_sums || !is_device_ptr(out_sums)) RC_TRY(T.alloc(&dsum, nblocks * 2));
    const unsigned grid = grid_for(nblocks * 8, 256);
    if (in_bits == 8) LAUNCH(ctx, beta_blocks_k<uint8_t>, grid, 256, 0, (const uint8_t *)dbeta, nsites, (const int32_t *)dbs, (const int32_t *)dbe, nblocks, dsum);
    else LAUNCH(ctx, beta_blocks_k<uint16_t>, grid, 256, 0, (const uint16_t *)dbeta, nsites, (const int32_t *)dbs, (const int32_t *)dbe, nblocks, dsum);
    if (out_trimmed) {
        const size_t ob = nblocks * 2 * (out_bits / 8);
        void *dout = out_trimmed;
        if (!is_device_ptr(out_trimmed)) RC_TRY(T.alloc((char **)&dout, ob));
        if (out_bits == 8) LAUNCH(ctx, trim64_k<uint8_t>, grid_for(nblocks, 256), 256, 0, dsum, nblocks, 255ll, (uint8_t *)dout);
        else LAUNCH(ctx, trim64_k<uint16_t>, grid_for(nblocks, 256), 256, 0, dsum, nblocks, 65535ll, (uint16_t *)dout);
        LAUNCH_CHECK();
        if (dout != out_trimmed) RC_TRY(copy_any(ctx, out_trimmed, dout, ob));
    }
    LAUNCH_CHECK();
    if (out_sums && (void *)dsum != (void *)out_sums) RC_TRY(copy_any(ctx, out_sums, dsum, nblocks * 16));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return 0;
}
