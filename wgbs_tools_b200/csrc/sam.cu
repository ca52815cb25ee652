// sam.cu -- SAM text (the stdin of the reference's `patter`, pipeline_wgbs/patter.cpp:381-416) -> ReadBatch.
//
// Replaces line2tokens (pipeline_wgbs/patter_utils.cpp:9-18) and the per-field std::stoi calls; also produces the QNAME
// hash that template pairing sorts on.  Flat data-parallel design (see below): the text is read twice with aligned
// 16-byte loads, every tab learns its ordinal inside its line from a segmented scan.
#include "lines.cuh"
#include "reads.cuh"

namespace {

constexpr int NTAB = 11;   // tab ordinals 0..10 delimit the 11 mandatory fields

__device__ __forceinline__ uint64_t fmix64(uint64_t k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
    return k;
}

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

// does the field starting at p carry an MM / ML tag?  1 = "MM:Z:" | "Mm:Z:", 2 = "ML:B:C" | "Ml:B:C"
__device__ __forceinline__ int tag_kind(const char *__restrict__ t, uint32_t p, uint32_t e) {
    if (p + 5 > e || t[p] != 'M' || t[p + 2] != ':') return 0;
    const char c1 = t[p + 1];
    if ((c1 == 'M' || c1 == 'm') && t[p + 3] == 'Z' && t[p + 4] == ':') return 1;
    if ((c1 == 'L' || c1 == 'l') && p + 6 <= e && t[p + 3] == 'B' && t[p + 4] == ':' && t[p + 5] == 'C') return 2;
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// Flat single-pass tokenizer.  The text is cut into 16 KiB tiles (one CTA each, handed out by an atomic ticket); a thread
// owns 64 consecutive bytes.  tk_scan_k reads every byte ONCE:
//   - byte-compare masks for '\t' and '\n' (aligned LDG.128, coalesced; masks are re-distributed inside each warp through
//     shared memory so that a lane ends up with its 4 consecutive chunks),
//   - a segmented scan gives every tab its ordinal inside its line and every newline its line number:
//       state = (newlines so far, tabs since the last newline);  combine(l, r) = (l.nl + r.nl, r.nl ? r.tail : l.tail + r.tail)
//     lanes -> warp shuffles, warps -> shared memory (one __syncthreads), tiles -> decoupled look-back over a status word
//     per tile (flag | newline count | tab tail), so no second pass over the text is needed,
//   - newline offsets go to nlpos[line], the first 11 tab offsets of a line to ftab[line*11 + ordinal]; a tab that starts
//     a tag field (ordinal >= 10) is checked for MM / ML.
// tk_records_k then turns one line per thread into a ReadBatch record (FLAG / POS, field spans, QNAME hash, tag spans).
// The line count is not known before the pass: output arrays are sized for an average line of >= 96 bytes, and the pass
// is repeated with the exact size in the (pathological) case that this was not enough.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int TK2_T = 256, TK2_CPT = 4, TK2_TILE = TK2_T * TK2_CPT * 16;   // 16 KiB
static_assert(TK2_CPT == 4, "a lane reads its 4 masks as one uint4");
constexpr unsigned long long TS_AGG = 1ull << 62, TS_INC = 2ull << 62;

struct Seg { uint32_t nl, tail; };
__device__ __forceinline__ Seg seg_combine(Seg l, Seg r) { Seg o; o.nl = l.nl + r.nl; o.tail = r.nl ? r.tail : l.tail + r.tail; return o; }
__device__ __forceinline__ unsigned long long seg_pack(Seg s, unsigned long long flag) { return flag | ((unsigned long long)s.nl << 16) | (s.tail > 0xffffu ? 0xffffu : s.tail); }
__device__ __forceinline__ Seg seg_unpack(unsigned long long w) { Seg s; s.nl = (uint32_t)(w >> 16); s.tail = (uint32_t)(w & 0xffffu); return s; }

__global__ void __launch_bounds__(TK2_T) tk_scan_k(const char *__restrict__ text, size_t n, uint32_t line_cap, int want_tags,
                                                    unsigned long long *__restrict__ status, unsigned int *__restrict__ ticket,
                                                    uint32_t *__restrict__ nlpos, uint32_t *__restrict__ ftab,
                                                    uint32_t *__restrict__ mm_off, uint32_t *__restrict__ ml_off, uint32_t *__restrict__ totals) {
    __shared__ uint32_t sm_mask[TK2_T * TK2_CPT];      // (nl << 16) | tab per chunk; warp w owns [w*128, w*128+128)
    __shared__ Seg ws[TK2_T / 32];
    __shared__ Seg s_start;
    __shared__ unsigned s_tile;
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const unsigned tile = s_tile;
    const size_t warp0 = (size_t)tile * TK2_TILE + (size_t)w * (32 * TK2_CPT * 16);   // first byte of this warp's 2 KiB
    // ---- masks, coalesced: round c, lane l -> chunk c*32 + l of the warp
#pragma unroll
    for (int c = 0; c < TK2_CPT; c++) {
        const size_t p = warp0 + (size_t)(c * 32 + lane) * 16;
        uint32_t tm = 0, nm = 0;
        if (p < n) {
            const uint4 v = (p + 16 <= n) ? *reinterpret_cast<const uint4 *>(text + p) : load16_guard(text, p, n);
            tm = eq_mask16(v, '\t'); nm = eq_mask16(v, '\n');
            if (n - p < 16) { const uint32_t ok = (1u << (n - p)) - 1; tm &= ok; nm &= ok; }
        }
        sm_mask[w * 128 + c * 32 + lane] = (nm << 16) | tm;
    }
    __syncwarp();
    const uint4 q = *reinterpret_cast<const uint4 *>(&sm_mask[w * 128 + lane * TK2_CPT]);   // this lane's 4 consecutive chunks
    const uint32_t mk[4] = {q.x, q.y, q.z, q.w};
    // ---- lane state, warp scan, warp aggregate
    Seg mine; mine.nl = 0; mine.tail = 0;
#pragma unroll
    for (int c = 0; c < TK2_CPT; c++) {
        const uint32_t nm = mk[c] >> 16, tm = mk[c] & 0xffffu;
        if (nm) { mine.nl += __popc(nm); mine.tail = __popc(tm & ~((2u << (31 - __clz(nm))) - 1)); }   // tabs after the chunk's last newline
        else mine.tail += __popc(tm);
    }
    Seg inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        Seg o; o.nl = __shfl_up_sync(0xffffffffu, inc.nl, d); o.tail = __shfl_up_sync(0xffffffffu, inc.tail, d);
        if (lane >= (unsigned)d) inc = seg_combine(o, inc);
    }
    Seg ex; ex.nl = __shfl_up_sync(0xffffffffu, inc.nl, 1); ex.tail = __shfl_up_sync(0xffffffffu, inc.tail, 1);
    if (lane == 0) { ex.nl = 0; ex.tail = 0; }
    if (lane == 31) ws[w] = inc;
    __syncthreads();
    Seg pre; pre.nl = 0; pre.tail = 0;
    Seg tot; tot.nl = 0; tot.tail = 0;
#pragma unroll
    for (int i = 0; i < TK2_T / 32; i++) { if (i < (int)w) pre = seg_combine(pre, ws[i]); tot = seg_combine(tot, ws[i]); }
    // ---- tile start state: decoupled look-back by warp 0, 32 predecessor tiles per probe.  Lane j holds tile-1-j; an
    // ordered warp scan combines them (earlier tiles on the left) up to the nearest tile that already published an
    // inclusive state.
    if (w == 0) {
        volatile unsigned long long *st = status;
        Seg start; start.nl = 0; start.tail = 0;
        if (tile == 0) { if (lane == 0) st[0] = seg_pack(tot, TS_INC); }
        else {
            if (lane == 0) st[tile] = seg_pack(tot, TS_AGG);
            Seg acc; acc.nl = 0; acc.tail = 0;           // combined state of the tiles already folded in (nearest ones)
            long long top = (long long)tile - 1;
            while (true) {
                const long long j = top - lane;
                unsigned long long x = TS_INC;           // before tile 0: the empty inclusive state
                if (j >= 0) x = st[j];
                while (__any_sync(0xffffffffu, (x >> 62) == 0)) { if ((x >> 62) == 0) x = st[j]; }
                const unsigned incm = __ballot_sync(0xffffffffu, (x >> 62) == 2);
                Seg v = seg_unpack(x & ~(3ull << 62));
                // inclusive ordered scan: lane j <- S(top-j) o ... o S(top)
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    Seg o; o.nl = __shfl_up_sync(0xffffffffu, v.nl, d); o.tail = __shfl_up_sync(0xffffffffu, v.tail, d);
                    if (lane >= (unsigned)d) v = seg_combine(v, o);
                }
                const int L = incm ? __ffs(incm) - 1 : 31;
                Seg win; win.nl = __shfl_sync(0xffffffffu, v.nl, L); win.tail = __shfl_sync(0xffffffffu, v.tail, L);
                acc = seg_combine(win, acc);
                if (incm) break;
                top -= 32;
            }
            start = acc;
            if (lane == 0) st[tile] = seg_pack(seg_combine(start, tot), TS_INC);
        }
        if (lane == 0) {
            s_start = start;
            if ((size_t)(tile + 1) * TK2_TILE >= n) { const Seg e = seg_combine(start, tot); totals[0] = e.nl; }   // last tile: newline total
        }
    }
    __syncthreads();
    const Seg stt = seg_combine(seg_combine(s_start, pre), ex);   // state just before this lane's first byte
    uint32_t line = stt.nl, ord = stt.tail;
    const size_t p0 = warp0 + (size_t)lane * (TK2_CPT * 16);
#pragma unroll
    for (int c = 0; c < TK2_CPT; c++) {
        const uint32_t nm = mk[c] >> 16;
        uint32_t m = (mk[c] & 0xffffu) | nm;
        while (m) {
            const int b = __ffs(m) - 1; m &= m - 1;
            const uint32_t x = (uint32_t)(p0 + (size_t)c * 16 + b);
            if ((nm >> b) & 1u) {
                if (line < line_cap) {
                    nlpos[line] = x;
                    for (uint32_t k = ord; k < NTAB; k++) ftab[(size_t)line * NTAB + k] = 0xffffffffu;   // fewer than 11 tabs: mark the rest absent
                }
                line++; ord = 0;
            }
            else {
                if (line < line_cap) {
                    if (ord < NTAB) ftab[(size_t)line * NTAB + ord] = x;
                    if (want_tags && ord >= 10) {                      // field index ord+1 >= 11: a tag
                        const int k = tag_kind(text, x + 1, (uint32_t)n);
                        if (k == 1) atomicMax(&mm_off[line], x + 6); else if (k == 2) atomicMax(&ml_off[line], x + 7);   // last occurrence wins
                    }
                }
                ord++;
            }
        }
    }
    // a last line without a trailing newline is closed by the lane that owns the final byte
    if (n && p0 <= n - 1 && n - 1 < p0 + TK2_CPT * 16 && text[n - 1] != '\n' && line < line_cap)
        for (uint32_t k = ord; k < NTAB; k++) ftab[(size_t)line * NTAB + k] = 0xffffffffu;
}

// per-(byte, position) mixing summed over the name
__device__ __forceinline__ uint64_t name_byte_mix(uint32_t byte, uint32_t pos) {
    uint64_t x = ((uint64_t)(byte | (pos << 8)) + 1) * 0x9e3779b97f4a7c15ULL;
    x ^= x >> 29; x *= 0xbf58476d1ce4e5b9ULL; x ^= x >> 32;
    return x;
}

__global__ void __launch_bounds__(256) tk_records_k(const char *__restrict__ text, uint32_t n, const uint32_t *__restrict__ nlpos, uint32_t n_nl,
                                                     uint32_t n_lines, const uint32_t *__restrict__ ftab, int want_tags, ReadBatch rb) {
    const uint32_t line = blockIdx.x * blockDim.x + threadIdx.x;
    if (line >= n_lines) return;
    const uint32_t s = line == 0 ? 0 : nlpos[line - 1] + 1;
    const uint32_t e = line < n_nl ? nlpos[line] : n;
    uint32_t tb[NTAB];
#pragma unroll
    for (int i = 0; i < NTAB; i++) tb[i] = ftab[(size_t)line * NTAB + i];
    const uint32_t qend = tb[0] != 0xffffffffu ? tb[0] : e;
    uint64_t h = 0;
    for (uint32_t p = s; p < qend; p++) h += name_byte_mix((uint8_t)text[p], p - s);
    h = fmix64(h ^ (uint64_t)(qend - s));
    uint8_t st = REC_OK;
    int32_t flag = 0, pos = 0;
    uint32_t cig_off = s, cig_len = 0, seq_off = s, seq_len = 0;
    if (e == s) { st = REC_BLANK; h = fmix64(0x5851f42d4c957f2dULL ^ (uint64_t)line); }   // blank lines: unique keys, never paired
    else if (tb[9] == 0xffffffffu || (tb[10] == 0xffffffffu && tb[9] == e - 1)) st = REC_INVALID;   // line2tokens would give < 11 tokens
    else {
        if (!parse_i32(text, tb[0] + 1, tb[1], &flag) || !parse_i32(text, tb[2] + 1, tb[3], &pos)) st = REC_BADINT;
        cig_off = tb[4] + 1; cig_len = tb[5] - tb[4] - 1;
        seq_off = tb[8] + 1; seq_len = tb[9] - tb[8] - 1;
    }
    rb.line_off[line] = s; rb.line_len[line] = e - s; rb.qn_len[line] = qend - s;
    rb.flag[line] = flag; rb.pos[line] = pos;
    rb.cig_off[line] = cig_off; rb.cig_len[line] = cig_len; rb.seq_off[line] = seq_off; rb.seq_len[line] = seq_len;
    rb.hash_lo[line] = (uint32_t)h; rb.hash_hi[line] = (uint32_t)(h >> 32);
    rb.status[line] = st;
    if (want_tags) {
        // mm_off / ml_off hold the payload start (0: no tag); the payload runs to the next tab or the line end
        const uint32_t mo = rb.mm_off[line], lo = rb.ml_off[line];
        uint32_t z = mo; if (mo) while (z < e && text[z] != '\t') z++;
        rb.mm_len[line] = mo ? z - mo : 0;
        z = lo; if (lo) while (z < e && text[z] != '\t') z++;
        rb.ml_len[line] = lo ? z - lo : 0;
    }
}

}  // namespace

int sam_tokenize(wgbs_ctx *ctx, const char *dtext, size_t nbytes, bool want_tags, Temps &T, ReadBatch *out) {
    if (nbytes >= 0xfffffff0ull) return wgbs_set_err("SAM text must be < 4 GiB per call (got %zu); split on line boundaries", nbytes);
    const uint32_t ntiles = (uint32_t)((nbytes + TK2_TILE - 1) / TK2_TILE);
    ReadBatch rb;
    rb.text = dtext; rb.nbytes = (uint32_t)nbytes;
    uint32_t *nlpos = nullptr, *ftab = nullptr, *totals = ctx->d_flags + 8;
    unsigned long long *status = nullptr;
    if (ntiles) RC_TRY(T.alloc(&status, (size_t)ntiles + 1));
    uint32_t cap = (uint32_t)(nbytes / 96 + 1024);            // optimistic: average line >= 96 bytes (a 50 bp SAM record is ~130)
    uint32_t n_nl = 0; char last = '\n';
    std::vector<void *> sized;                                 // arrays sized by `cap`, re-made if the guess was too small
    for (int attempt = 0; attempt < 2 && ntiles; attempt++) {
        for (void *q : sized) { T.keep(q); dfree(ctx, q); }
        sized.clear();
        RC_TRY(T.alloc(&nlpos, cap)); sized.push_back(nlpos);
        RC_TRY(T.alloc(&ftab, (size_t)cap * NTAB)); sized.push_back(ftab);
        if (want_tags) {
            RC_TRY(T.alloc(&rb.mm_off, cap)); sized.push_back(rb.mm_off); RC_TRY(T.alloc(&rb.ml_off, cap)); sized.push_back(rb.ml_off);
            CUDA_TRY(cudaMemsetAsync(rb.mm_off, 0, (size_t)cap * 4, ctx->stream));
            CUDA_TRY(cudaMemsetAsync(rb.ml_off, 0, (size_t)cap * 4, ctx->stream));
        }
        CUDA_TRY(cudaMemsetAsync(status, 0, ((size_t)ntiles + 1) * 8, ctx->stream));
        LAUNCH(ctx, tk_scan_k, ntiles, TK2_T, 0, dtext, nbytes, cap, want_tags ? 1 : 0, status, (unsigned int *)(status + ntiles), nlpos, ftab,
               rb.mm_off, rb.ml_off, totals);
        CUDA_TRY(cudaMemcpyAsync(&n_nl, totals, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(&last, dtext + nbytes - 1, 1, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        if (n_nl + 1 <= cap) break;
        cap = n_nl + 1;                                       // many very short lines: repeat with the exact size
    }
    const uint32_t n_lines = n_nl + ((nbytes && last != '\n') ? 1 : 0);
    rb.n = n_lines;
    RC_TRY(T.alloc(&rb.line_off, n_lines)); RC_TRY(T.alloc(&rb.line_len, n_lines)); RC_TRY(T.alloc(&rb.qn_len, n_lines));
    RC_TRY(T.alloc(&rb.flag, n_lines)); RC_TRY(T.alloc(&rb.pos, n_lines));
    RC_TRY(T.alloc(&rb.cig_off, n_lines)); RC_TRY(T.alloc(&rb.cig_len, n_lines));
    RC_TRY(T.alloc(&rb.seq_off, n_lines)); RC_TRY(T.alloc(&rb.seq_len, n_lines));
    RC_TRY(T.alloc(&rb.hash_lo, n_lines)); RC_TRY(T.alloc(&rb.hash_hi, n_lines));
    RC_TRY(T.alloc(&rb.status, n_lines));
    if (want_tags) { RC_TRY(T.alloc(&rb.mm_len, n_lines)); RC_TRY(T.alloc(&rb.ml_len, n_lines)); if (!ntiles) { RC_TRY(T.alloc(&rb.mm_off, 1)); RC_TRY(T.alloc(&rb.ml_off, 1)); } }
    if (n_lines) {
        LAUNCH(ctx, tk_records_k, grid_for(n_lines, 256), 256, 0, dtext, (uint32_t)nbytes, nlpos, n_nl, n_lines, ftab, want_tags ? 1 : 0, rb);
        LAUNCH_CHECK();
    }
    *out = rb;
    return 0;
}
