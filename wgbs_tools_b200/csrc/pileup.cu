// pileup.cu -- the bam->pat pileup: per-record CIGAR walk + CpG-index lookup + C/T call, mate merge, template filter.
//
// Reference behaviour restated (paths relative to the reference's src/pipeline_wgbs/):
//   patter_utils.cpp:209-251  clean_CIGAR        -> cig_validate / CigCursor (no adjusted string is materialised)
//   patter_utils.cpp:163-168  is_bottom          -> is_bottom
//   patter.cpp:96-184         is_cpg / compareSeqToRef -> pileup_call_k
//   patter_utils.cpp:292-342  merge_PE (+strip)  -> merge_templates_k
//   patter.cpp:247-290        proc2lines (min_cpg filter, counters) -> merge_templates_k
//   patter.cpp:14-42,81-93    CpG dictionary: unordered_map + bool conv[chromlen]  -> sorted uint32 loci[] + binary search
//
// One thread per record.  CIGAR and SEQ bytes are read in place from the SAM text buffer.  Calls are written as 2-bit
// symbols into a word pool (layout: include/wgbs_b200.h).
#include "pileup_dev.cuh"

int build_mates(wgbs_ctx *ctx, const ReadBatch &rb, bool paired, Temps &T, uint32_t **mate_out, unsigned long long *d_stats);
int pileup_records(wgbs_ctx *ctx, const wgbs_index *ix, const ReadBatch &rb, const wgbs_pileup_opts *opts, Temps &T, wgbs_pats **out,
                   uint64_t *stats_out, int32_t *mbias_out);
// np.cu (MM/ML mode)
int np_measure(wgbs_ctx *ctx, const ReadBatch &rb, const uint32_t *loci, uint32_t nloci, PileupOpts o, uint32_t *r_lo, uint32_t *r_ncand,
               uint32_t *words, unsigned long long *d_stats);
int np_call(wgbs_ctx *ctx, const ReadBatch &rb, const uint32_t *loci, uint32_t first_idx, PileupOpts o, const uint32_t *r_lo,
            const uint32_t *r_ncand, const uint32_t *off, uint32_t *pool, uint32_t pool_words, int32_t *r_idx, uint32_t *r_len,
            unsigned long long *d_stats);

namespace {

// ---- pass 1: validity, reference span, candidate CpG range ---------------------------------------------------------
__global__ void __launch_bounds__(256) pileup_measure_k(ReadBatchView rb, const uint32_t *__restrict__ loci, uint32_t nloci, PileupOpts o,
                                                         uint32_t *__restrict__ r_lo, uint32_t *__restrict__ r_ncand,
                                                         uint32_t *__restrict__ words, unsigned long long *__restrict__ stats) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t inval = 0;
    if (r < rb.n) {
        uint32_t lo = 0, nc = 0;
        uint8_t st = rb.status[r];
        if (st == REC_INVALID || st == REC_BADINT) inval = 1;
        else if (st == REC_OK) {
            int64_t span = 0;
            bool ok = rec_cig_validate(rb, r, &span);
            if (!ok) inval = 1;
            else {
                int64_t pos = rb.pos[r];
                lo = lower_bound_u32(loci, nloci, pos);
                uint32_t hi = lower_bound_u32(loci, nloci, pos + span);
                nc = hi > lo ? hi - lo : 0;
            }
        }
        r_lo[r] = lo; r_ncand[r] = inval ? NONE : nc;
        words[r] = inval ? 0 : (nc + 15) >> 4;
    }
    for (int d = 16; d >= 1; d >>= 1) inval += __shfl_xor_sync(0xffffffffu, inval, d);
    if ((threadIdx.x & 31) == 0 && inval) atomicAdd(&stats[ST_INVALID], (unsigned long long)inval);
}

// merged-slot size of a template head (bounded by MAX_PE_PAT_LEN symbols: longer merges are dropped)
__global__ void __launch_bounds__(256) template_words_k(uint32_t n, const uint32_t *__restrict__ mate, const uint32_t *__restrict__ r_lo,
                                                         const uint32_t *__restrict__ r_ncand, uint32_t *__restrict__ words_t) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    uint32_t m = mate[r], w = 0;
    if (m != NONE && r < m && r_ncand[r] != NONE && r_ncand[m] != NONE && r_ncand[r] && r_ncand[m]) {
        uint32_t lo = min(r_lo[r], r_lo[m]), hi = max(r_lo[r] + r_ncand[r], r_lo[m] + r_ncand[m]);
        uint32_t u = min(hi - lo, (uint32_t)MAX_PE_PAT_LEN);
        w = (u + 15) >> 4;
    }
    words_t[r] = w;
}

// ---- pass 2: the calls ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pileup_call_k(ReadBatchView rb, const uint32_t *__restrict__ loci, uint32_t first_idx, PileupOpts o,
                                                      const uint32_t *__restrict__ r_lo, const uint32_t *__restrict__ r_ncand,
                                                      const uint32_t *__restrict__ off, uint32_t *__restrict__ pool,
                                                      int32_t *__restrict__ r_idx, uint32_t *__restrict__ r_len,
                                                      unsigned long long *__restrict__ stats, int32_t *__restrict__ mbias) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t empty = 0;
    if (r < rb.n && rb.status[r] == REC_OK && r_ncand[r] != NONE) {
        const uint32_t nc = r_ncand[r], lo = r_lo[r];
        const int64_t pos = rb.pos[r];
        const int flag = rb.flag[r];
        const bool bottom = is_bottom(flag, o.paired);
        int64_t span; rec_cig_validate(rb, r, &span);
        CigCursor cc; cc.init(rb, r);
        const uint32_t seq_off = rb.seq_off[r];
        uint32_t *wp = pool + off[r];
        int32_t first = -1, last = -1;      // candidate ordinals of the first / last called ('C'/'T') site
        uint32_t w = 0; int32_t nsym = 0;   // symbols emitted since `first`
        // M-bias table slot of this read (patter.cpp:116-129): [OT|OB][mate 0|1][1000 positions][meth|unmeth], or none
        int32_t *mb = nullptr;
        if (mbias) {
            int mate = 0; bool skip = false;
            if (o.paired) {
                if ((flag & 0x53) == 0x53) mate = 0; else if ((flag & 0xA3) == 0xA3) mate = 1;
                else if ((flag & 0x63) == 0x63) mate = 0; else if ((flag & 0x93) == 0x93) mate = 1; else skip = true;
            }
            // bottom reads index positions from the far end: the first step (mj = len-1) already trips the MAX_READ_LEN guard (:143)
            if (bottom && span - 1 >= 1000) skip = true;
            if (!skip) mb = mbias + ((bottom ? 1 : 0) * 2 + mate) * 2000;
        }
        for (uint32_t j = 0; j < nc; j++) {
            const int64_t i = (int64_t)loci[lo + j] - pos;     // offset of the CpG's C in the reference-projected read
            // adj[i] and adj[i+1]
            char c0 = 0, c1 = 0;
            if (cc.seek(i)) c0 = cc.op == 'M' ? rec_base(rb, seq_off, cc.q0 + (i - cc.r0)) : 'N';
            if (i + 1 < span && cc.seek(i + 1)) c1 = cc.op == 'M' ? rec_base(rb, seq_off, cc.q0 + (i + 1 - cc.r0)) : 'N';
            uint32_t code = SYM_DOT;
            int64_t jx;
            if (!bottom) {      // OT: C/T at the C position, G must follow (patter.cpp:96-103 is_cpg, shift 0)
                jx = i;
                if ((i < span - 1) && c1 == 'G') code = c0 == 'T' ? SYM_T : (c0 == 'C' ? SYM_C : SYM_DOT);
            } else {            // OB: G/A at the G position, C must precede (shift 1)
                jx = i + 1;
                if (c0 == 'C') code = c1 == 'A' ? SYM_T : (c1 == 'G' ? SYM_C : SYM_DOT);
            }
            if (mb && code != SYM_DOT) {                                       // counted before the clip (patter.cpp:152-165)
                const int64_t mj = bottom ? span - i - 1 : i;
                if (mj < 1000) atomicAdd(&mb[2 * mj + (code == SYM_T ? 1 : 0)], 1);
            }
            if (!((jx >= o.clip) && (jx < span - o.clip))) code = SYM_DOT;     // patter.cpp:169-172
            if (first < 0) { if (code == SYM_DOT) continue; first = (int32_t)j; }
            if (code != SYM_DOT) last = (int32_t)j;
            w |= code << (30 - 2 * (nsym & 15));
            if ((++nsym & 15) == 0) { *wp++ = w; w = 0; }
        }
        if (nsym & 15) *wp = w;
        if (first < 0) { r_idx[r] = 0; r_len[r] = 0; empty = 1; }
        else { r_idx[r] = (int32_t)(first_idx + lo + (uint32_t)first); r_len[r] = (uint32_t)(last - first + 1); }
    } else if (r < rb.n) { r_idx[r] = 0; r_len[r] = 0; }
    for (int d = 16; d >= 1; d >>= 1) empty += __shfl_xor_sync(0xffffffffu, empty, d);
    if ((threadIdx.x & 31) == 0 && empty) atomicAdd(&stats[ST_EMPTY], (unsigned long long)empty);
}

// ---- mate merge + min_cpg filter ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t get_sym(const uint32_t *__restrict__ wp, uint32_t k) { return (wp[k >> 4] >> (30 - 2 * (k & 15))) & 3u; }

__global__ void __launch_bounds__(256) merge_templates_k(uint32_t n, const uint8_t *__restrict__ status, const uint32_t *__restrict__ mate,
                                                          PileupOpts o, const int32_t *__restrict__ r_idx, const uint32_t *__restrict__ r_len,
                                                          const uint32_t *__restrict__ off /*[2n+1]*/, uint32_t *__restrict__ pool,
                                                          uint32_t *__restrict__ t_idx, uint32_t *__restrict__ t_len, uint32_t *__restrict__ t_off,
                                                          uint32_t *__restrict__ t_valid, unsigned long long *__restrict__ stats) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t nshort = 0;
    if (r < n) {
        uint32_t valid = 0, oi = 0, ol = 0, oo = 0;
        const uint32_t m = mate[r];
        if (status[r] != REC_BLANK && (m == NONE || r < m)) {          // r is a template head
            const uint32_t la = r_len[r], lb = m == NONE ? 0 : r_len[m];
            if (la && !lb) { oi = (uint32_t)r_idx[r]; ol = la; oo = off[r]; valid = 1; }
            else if (lb && !la) { oi = (uint32_t)r_idx[m]; ol = lb; oo = off[m]; valid = 1; }
            else if (la && lb) {
                // merge_PE: l1 = the mate that starts first
                uint32_t a = r, b = m;
                if (r_idx[a] > r_idx[b]) { a = m; b = r; }
                const int32_t s1 = r_idx[a], s2 = r_idx[b];
                const uint32_t l1 = r_len[a], l2 = r_len[b];
                const int32_t lastp = max(s1 + (int32_t)l1, s2 + (int32_t)l2);
                const int32_t tot = lastp - s1;
                if (tot <= MAX_PE_PAT_LEN) {                             // else: "merged read is too long" -> dropped, not counted
                    const uint32_t *pa = pool + off[a], *pb = pool + off[b];
                    uint32_t *po = pool + off[n + r];
                    const uint32_t d = (uint32_t)(s2 - s1);
                    int32_t first = -1, lastk = -1;
                    // first pass: locate the first / last non-'.' merged symbol (conflicts become '.')
                    for (int32_t k = 0; k < tot; k++) {
                        uint32_t x = (uint32_t)k < l1 ? get_sym(pa, k) : 0;
                        uint32_t y = ((uint32_t)k >= d && (uint32_t)k - d < l2) ? get_sym(pb, k - d) : 0;
                        uint32_t z = x == 0 ? y : (y == 0 || y == x ? x : 0);
                        if (z) { if (first < 0) first = k; lastk = k; }
                    }
                    if (first >= 0) {
                        uint32_t w = 0; int32_t ns = 0;
                        for (int32_t k = first; k <= lastk; k++) {
                            uint32_t x = (uint32_t)k < l1 ? get_sym(pa, k) : 0;
                            uint32_t y = ((uint32_t)k >= d && (uint32_t)k - d < l2) ? get_sym(pb, k - d) : 0;
                            uint32_t z = x == 0 ? y : (y == 0 || y == x ? x : 0);
                            w |= z << (30 - 2 * (ns & 15));
                            if ((++ns & 15) == 0) { *po++ = w; w = 0; }
                        }
                        if (ns & 15) *po = w;
                        oi = (uint32_t)(s1 + first); ol = (uint32_t)(lastk - first + 1); oo = off[n + r]; valid = 1;
                    }
                }
            }
            if (valid && (int32_t)ol < o.min_cpg) { valid = 0; nshort = 1; }   // patter.cpp:264-267
        }
        t_idx[r] = oi; t_len[r] = ol; t_off[r] = oo; t_valid[r] = valid;
    }
    for (int d = 16; d >= 1; d >>= 1) nshort += __shfl_xor_sync(0xffffffffu, nshort, d);
    if ((threadIdx.x & 31) == 0 && nshort) atomicAdd(&stats[ST_SHORT], (unsigned long long)nshort);
}

// --long: template k keeps the QNAME of its head record (tokens1[0] in patter.cpp:274-276; both mates share it)
__global__ void __launch_bounds__(256) name_len_k(uint32_t n, const uint32_t *__restrict__ t_valid, const uint32_t *__restrict__ dst,
                                                   const uint32_t *__restrict__ qn_len, uint32_t *__restrict__ o_len) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n && t_valid[r]) o_len[dst[r]] = qn_len[r];
}
__global__ void __launch_bounds__(256) name_copy_k(uint32_t n, const uint32_t *__restrict__ t_valid, const uint32_t *__restrict__ dst, ReadBatchView rb,
                                                    const uint32_t *__restrict__ o_off, char *__restrict__ names) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n || !t_valid[r]) return;
    const char *src = rb.text + rb.line_off[r]; char *d = names + o_off[dst[r]];
    for (uint32_t k = 0; k < rb.qn_len[r]; k++) d[k] = src[k];
}

__global__ void __launch_bounds__(256) compact_templates_k(uint32_t n, const uint32_t *__restrict__ t_valid, const uint32_t *__restrict__ dst,
                                                            const uint32_t *__restrict__ t_idx, const uint32_t *__restrict__ t_len,
                                                            const uint32_t *__restrict__ t_off, uint32_t *__restrict__ o_idx,
                                                            uint32_t *__restrict__ o_len, uint32_t *__restrict__ o_off, uint32_t *__restrict__ o_cnt) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n || !t_valid[r]) return;
    uint32_t k = dst[r];
    o_idx[k] = t_idx[r]; o_len[k] = t_len[r]; o_off[k] = t_off[r]; o_cnt[k] = 1;
}

}  // namespace

// ====================================================================================================================
// C ABI
// ====================================================================================================================
extern "C" int wgbs_index_load(wgbs_ctx *ctx, const uint32_t *loci, size_t n, uint32_t first_idx, wgbs_index **out) {
    RC_TRY(wgbs_ctx_activate(ctx));
    if (!out) return wgbs_set_err("wgbs_index_load: out is null");
    if (n >= 0x7fffffffull) return wgbs_set_err("wgbs_index_load: too many loci");
    wgbs_index *ix = new wgbs_index();
    ix->n = (uint32_t)n; ix->first_idx = first_idx;
    int rc = dalloc(ctx, &ix->loci, n);
    if (rc == 0) rc = copy_any(ctx, ix->loci, loci, n * 4);
    if (rc < 0) { dfree(ctx, ix->loci); delete ix; return rc; }
    *out = ix;
    return 0;
}
extern "C" void wgbs_index_free(wgbs_ctx *ctx, wgbs_index *ix) {
    if (!ix || !ctx) return;
    cudaSetDevice(ctx->device);
    dfree(ctx, ix->loci);
    delete ix;
}

extern "C" int wgbs_pileup_sam(wgbs_ctx *ctx, const wgbs_index *ix, const char *sam, size_t nbytes, const wgbs_pileup_opts *opts,
                               wgbs_pats **out, uint64_t *stats_out) {
    return wgbs_pileup_sam_mbias(ctx, ix, sam, nbytes, opts, out, stats_out, nullptr);
}

extern "C" int wgbs_pileup_sam_mbias(wgbs_ctx *ctx, const wgbs_index *ix, const char *sam, size_t nbytes, const wgbs_pileup_opts *opts,
                                     wgbs_pats **out, uint64_t *stats_out, int32_t *mbias_out) {
    RC_TRY(wgbs_ctx_activate(ctx));
    if (!ix || !opts || !out) return wgbs_set_err("wgbs_pileup_sam: null argument");
    *out = nullptr;
    Temps T(ctx);
    const void *dv = nullptr; bool owned = false;
    RC_TRY(to_device(ctx, sam, nbytes, &dv, &owned));
    const char *dtext = (const char *)dv;
    if (owned) T.v.push_back((void *)dtext);

    // first_line (patter.cpp:324-350): paired iff FLAG&1 of the first line; MM/ML mode iff requested or the first line
    // carries an MM tag (the tokenizer checks that and only then records tag spans).
    ReadBatch rb;
    RC_TRY(sam_tokenize(ctx, dtext, nbytes, opts->nanopore ? 1 : -1, T, &rb));
    return pileup_records(ctx, ix, rb, opts, T, out, stats_out, mbias_out);
}

// everything after the records exist: mates, calls, merge, filter.  rb: a tokenized SAM text, or a BAM batch (bamdev.cu)
int pileup_records(wgbs_ctx *ctx, const wgbs_index *ix, const ReadBatch &rb, const wgbs_pileup_opts *opts, Temps &T, wgbs_pats **out,
                   uint64_t *stats_out, int32_t *mbias_out) {
    const uint32_t n = rb.n;
    unsigned long long *d_stats;
    RC_TRY(T.alloc(&d_stats, ST_N));
    CUDA_TRY(cudaMemsetAsync(d_stats, 0, ST_N * 8, ctx->stream));

    PileupOpts o;
    o.min_cpg = opts->min_cpg; o.clip = opts->clip; o.nanopore = (opts->nanopore || rb.mm_off != nullptr) ? 1 : 0; o.combine_mods = opts->combine_mods;
    o.np_thresh = opts->np_thresh; o.cpc_call = opts->cpc_call ? opts->cpc_call : 'C';
    o.paired = opts->paired;
    {
        // host-side peek at the head of the batch (tiny D2H): first non-blank record
        const uint32_t PEEK = 64;
        uint32_t m = n < PEEK ? n : PEEK;
        std::vector<uint8_t> st(m); std::vector<int32_t> fl(m);
        if (m) {
            CUDA_TRY(cudaMemcpyAsync(st.data(), rb.status, m, cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_TRY(cudaMemcpyAsync(fl.data(), rb.flag, m * 4, cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        }
        uint32_t f = 0; while (f < m && st[f] == REC_BLANK) f++;
        if (f < m) {
            if ((st[f] == REC_INVALID || st[f] == REC_BADINT) && o.paired < 0) return wgbs_set_err("Invalid first line (cannot determine paired/single end)");
            if (o.paired < 0) o.paired = ((uint16_t)fl[f]) & 1;
        } else if (o.paired < 0) o.paired = 0;
        if (o.paired && o.nanopore) return wgbs_set_err("Unrecognized bam format: paired end and nanopore");   // patter.cpp:341-343
    }

    uint32_t *mate = nullptr;
    RC_TRY(build_mates(ctx, rb, o.paired != 0, T, &mate, d_stats));


    uint32_t *r_lo, *r_ncand, *words, *off;
    RC_TRY(T.alloc(&r_lo, n)); RC_TRY(T.alloc(&r_ncand, n)); RC_TRY(T.alloc(&words, (size_t)2 * n)); RC_TRY(T.alloc(&off, (size_t)2 * n + 1));
    if (n) {
        if (o.nanopore) RC_TRY(np_measure(ctx, rb, ix->loci, ix->n, o, r_lo, r_ncand, words, d_stats));
        else LAUNCH(ctx, pileup_measure_k, grid_for(n, 256), 256, 0, view_of(rb), ix->loci, ix->n, o, r_lo, r_ncand, words, d_stats);
        LAUNCH(ctx, template_words_k, grid_for(n, 256), 256, 0, n, mate, r_lo, r_ncand, words + n);
    }
    RC_TRY(scan_u32_u32(ctx, words, off, (size_t)2 * n));
    uint32_t pool_words = 0;
    CUDA_TRY(cudaMemcpyAsync(&pool_words, off + (size_t)2 * n, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));

    wgbs_pats *P = new wgbs_pats();
    int rc = dalloc(ctx, &P->pool, pool_words);
    if (rc < 0) { delete P; return rc; }
    P->pool_words = pool_words;
    int32_t *d_mbias = nullptr;
    if (mbias_out) {
        if (is_device_ptr(mbias_out)) d_mbias = mbias_out; else if ((rc = T.alloc(&d_mbias, 8000)) < 0) { wgbs_pats_free(ctx, P); return rc; }
        CUDA_TRY(cudaMemsetAsync(d_mbias, 0, 8000 * sizeof(int32_t), ctx->stream));
    }
    int32_t *r_idx; uint32_t *r_len, *t_idx, *t_len, *t_off, *t_valid, *dst;
    if ((rc = T.alloc(&r_idx, n)) < 0 || (rc = T.alloc(&r_len, n)) < 0 || (rc = T.alloc(&t_idx, n)) < 0 || (rc = T.alloc(&t_len, n)) < 0 ||
        (rc = T.alloc(&t_off, n)) < 0 || (rc = T.alloc(&t_valid, n)) < 0 || (rc = T.alloc(&dst, (size_t)n + 1)) < 0) { wgbs_pats_free(ctx, P); return rc; }
    if (n) {
        if (o.nanopore) { int rc2 = np_call(ctx, rb, ix->loci, ix->first_idx, o, r_lo, r_ncand, off, P->pool, pool_words, r_idx, r_len, d_stats); if (rc2 < 0) { wgbs_pats_free(ctx, P); return rc2; } }
        else LAUNCH(ctx, pileup_call_k, grid_for(n, 256), 256, 0, view_of(rb), ix->loci, ix->first_idx, o, r_lo, r_ncand, off, P->pool, r_idx, r_len, d_stats, d_mbias);
        LAUNCH(ctx, merge_templates_k, grid_for(n, 256), 256, 0, n, rb.status, mate, o, r_idx, r_len, off, P->pool, t_idx, t_len, t_off, t_valid, d_stats);
    }
    if ((rc = scan_u32_u32(ctx, t_valid, dst, n)) < 0) { wgbs_pats_free(ctx, P); return rc; }
    uint32_t n_out = 0;
    unsigned long long hstats[ST_N];
    CUDA_TRY(cudaMemcpyAsync(&n_out, dst + n, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(hstats, d_stats, sizeof hstats, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    P->n = n_out;
    if ((rc = dalloc(ctx, &P->idx, n_out)) < 0 || (rc = dalloc(ctx, &P->len, n_out)) < 0 || (rc = dalloc(ctx, &P->count, n_out)) < 0 ||
        (rc = dalloc(ctx, &P->off, (size_t)n_out + 1)) < 0) { wgbs_pats_free(ctx, P); return rc; }
    if (n) LAUNCH(ctx, compact_templates_k, grid_for(n, 256), 256, 0, n, t_valid, dst, t_idx, t_len, t_off, P->idx, P->len, P->off, P->count);
    if (opts->keep_names && n_out) {
        uint32_t *nl_tmp;
        if ((rc = dalloc(ctx, &P->name_len, n_out)) < 0 || (rc = dalloc(ctx, &P->name_off, (size_t)n_out + 1)) < 0) { wgbs_pats_free(ctx, P); return rc; }
        (void)nl_tmp;
        LAUNCH(ctx, name_len_k, grid_for(n, 256), 256, 0, n, t_valid, dst, rb.qn_len, P->name_len);
        if ((rc = scan_u32_u32(ctx, P->name_len, P->name_off, n_out)) < 0) { wgbs_pats_free(ctx, P); return rc; }
        uint32_t nb = 0;
        CUDA_TRY(cudaMemcpyAsync(&nb, P->name_off + n_out, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        if ((rc = dalloc(ctx, &P->names, nb)) < 0) { wgbs_pats_free(ctx, P); return rc; }
        P->names_bytes = nb;
        LAUNCH(ctx, name_copy_k, grid_for(n, 256), 256, 0, n, t_valid, dst, view_of(rb), P->name_off, P->names);
    }
    LAUNCH_CHECK();
    if (mbias_out && d_mbias != mbias_out) { if ((rc = copy_any(ctx, mbias_out, d_mbias, 8000 * sizeof(int32_t))) < 0) { wgbs_pats_free(ctx, P); return rc; } }
    if (stats_out) {
        stats_out[0] = n;                       // lines (patter counts blank lines too)
        stats_out[1] = hstats[ST_PAIRS]; stats_out[2] = hstats[ST_EMPTY]; stats_out[3] = hstats[ST_SHORT];
        stats_out[4] = hstats[ST_INVALID]; stats_out[5] = (uint64_t)o.paired; stats_out[6] = (uint64_t)o.nanopore; stats_out[7] = n_out;
    }
    *out = P;
    return 0;
}
