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
// Flat tokenizer.  The text is cut into 16 KiB tiles; a thread owns 64 consecutive bytes (4 aligned LDG.128).
//   tk_count_k     per tile: (#newlines, #tabs after the tile's last newline [all tabs if it has none])
//   tk_tilescan_k  one CTA: running (newline count, tabs since the last newline) at every tile start
//   tk_mark_k      per tile again: every newline -> nlpos[line]; every tab -> its ordinal inside its line; the first 11 go
//                  to ftab[line*11 + ordinal]; tabs that start a tag field are checked for MM/ML
//   tk_records_k   one thread per line: FLAG / POS, field spans, QNAME hash, MM/ML spans
// The "tabs since the last newline" state is a segmented sum: combine(l, r) = (l.nl + r.nl, r.nl ? r.tail : l.tail + r.tail).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int TK2_T = 256, TK2_CPT = 4, TK2_TILE = TK2_T * TK2_CPT * 16;   // 16 KiB
static_assert(TK2_CPT == 4, "load_masks reads 4 masks per thread as one uint4");

struct Seg { uint32_t nl, tail; };
__device__ __forceinline__ Seg seg_combine(Seg l, Seg r) { Seg o; o.nl = l.nl + r.nl; o.tail = r.nl ? r.tail : l.tail + r.tail; return o; }

// masks of the 64 bytes a thread owns: bit b of m[c] <-> byte 16*c + b.
// Global loads are coalesced (consecutive lanes read consecutive 16-byte chunks, TK2_CPT rounds); the (tab, newline)
// masks are then transposed through shared memory so that each thread ends up with its 4 CONSECUTIVE chunks.
__device__ __forceinline__ void load_masks(const char *__restrict__ text, size_t n, size_t tile0, uint32_t *tabm, uint32_t *nlm) {
    __shared__ uint32_t sm_mask[TK2_T * TK2_CPT];      // (nl << 16) | tab per chunk, chunk order
#pragma unroll
    for (int c = 0; c < TK2_CPT; c++) {
        const uint32_t chunk = c * TK2_T + threadIdx.x;
        const size_t p = tile0 + (size_t)chunk * 16;
        uint32_t tm = 0, nm = 0;
        if (p < n) {
            const uint4 v = (p + 16 <= n) ? *reinterpret_cast<const uint4 *>(text + p) : load16_guard(text, p, n);
            tm = eq_mask16(v, '\t'); nm = eq_mask16(v, '\n');
            if (n - p < 16) { const uint32_t ok = (1u << (n - p)) - 1; tm &= ok; nm &= ok; }
        }
        sm_mask[chunk] = (nm << 16) | tm;
    }
    __syncthreads();
    const uint4 q = *reinterpret_cast<const uint4 *>(&sm_mask[threadIdx.x * TK2_CPT]);
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int c = 0; c < TK2_CPT; c++) { tabm[c] = w[c] & 0xffffu; nlm[c] = w[c] >> 16; }
    __syncthreads();
}
__device__ __forceinline__ Seg thread_seg(const uint32_t *tabm, const uint32_t *nlm) {
    Seg s; s.nl = 0; s.tail = 0;
#pragma unroll
    for (int c = 0; c < TK2_CPT; c++) {
        const uint32_t nm = nlm[c], tm = tabm[c];
        if (nm) { s.nl += __popc(nm); s.tail = __popc(tm & ~((2u << (31 - __clz(nm))) - 1)); }   // tabs after the chunk's last newline
        else s.tail += __popc(tm);
    }
    return s;
}
// block-wide EXCLUSIVE segmented scan of one Seg per thread (thread order = byte order); *total = all threads combined
__device__ __forceinline__ Seg block_seg_excl(Seg v, Seg *total) {
    __shared__ Seg ws[TK2_T / 32];
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    Seg inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        Seg o; o.nl = __shfl_up_sync(0xffffffffu, inc.nl, d); o.tail = __shfl_up_sync(0xffffffffu, inc.tail, d);
        if (lane >= (unsigned)d) inc = seg_combine(o, inc);
    }
    if (lane == 31) ws[w] = inc;
    Seg ex; ex.nl = __shfl_up_sync(0xffffffffu, inc.nl, 1); ex.tail = __shfl_up_sync(0xffffffffu, inc.tail, 1);
    if (lane == 0) { ex.nl = 0; ex.tail = 0; }
    __syncthreads();
    Seg pre; pre.nl = 0; pre.tail = 0;
    Seg tot; tot.nl = 0; tot.tail = 0;
#pragma unroll
    for (int i = 0; i < TK2_T / 32; i++) { if (i < (int)w) pre = seg_combine(pre, ws[i]); tot = seg_combine(tot, ws[i]); }
    *total = tot;
    __syncthreads();
    return seg_combine(pre, ex);
}

__global__ void __launch_bounds__(TK2_T) tk_count_k(const char *__restrict__ text, size_t n, uint2 *__restrict__ tile_seg) {
    uint32_t tabm[TK2_CPT], nlm[TK2_CPT];
    load_masks(text, n, (size_t)blockIdx.x * TK2_TILE, tabm, nlm);
    Seg tot;
    block_seg_excl(thread_seg(tabm, nlm), &tot);
    if (threadIdx.x == 0) tile_seg[blockIdx.x] = make_uint2(tot.nl, tot.tail);
}

// in place: tile_seg[t] <- state at the START of tile t;  tile_seg[ntiles] <- state at the end of the text
__global__ void __launch_bounds__(1024) tk_tilescan_k(uint2 *__restrict__ tile_seg, uint32_t ntiles) {
    __shared__ Seg ws[32];
    __shared__ Seg carry;
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) { carry.nl = 0; carry.tail = 0; }
    __syncthreads();
    for (uint32_t base = 0; base < ntiles; base += 1024) {
        const uint32_t t = base + threadIdx.x;
        Seg v; v.nl = 0; v.tail = 0;
        if (t < ntiles) { const uint2 q = tile_seg[t]; v.nl = q.x; v.tail = q.y; }
        Seg inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            Seg o; o.nl = __shfl_up_sync(0xffffffffu, inc.nl, d); o.tail = __shfl_up_sync(0xffffffffu, inc.tail, d);
            if (lane >= (unsigned)d) inc = seg_combine(o, inc);
        }
        if (lane == 31) ws[w] = inc;
        Seg ex; ex.nl = __shfl_up_sync(0xffffffffu, inc.nl, 1); ex.tail = __shfl_up_sync(0xffffffffu, inc.tail, 1);
        if (lane == 0) { ex.nl = 0; ex.tail = 0; }
        __syncthreads();
        Seg pre = carry;
        for (unsigned i = 0; i < w; i++) pre = seg_combine(pre, ws[i]);
        const Seg mine = seg_combine(pre, ex);
        if (t < ntiles) tile_seg[t] = make_uint2(mine.nl, mine.tail);
        __syncthreads();
        if (threadIdx.x == 1023) carry = seg_combine(mine, v);
        __syncthreads();
    }
    if (threadIdx.x == 0) tile_seg[ntiles] = make_uint2(carry.nl, carry.tail);
}

__global__ void __launch_bounds__(TK2_T) tk_mark_k(const char *__restrict__ text, size_t n, const uint2 *__restrict__ tile_start,
                                                    uint32_t n_lines, int want_tags, uint32_t *__restrict__ nlpos, uint32_t *__restrict__ ftab,
                                                    uint32_t *__restrict__ mm_off, uint32_t *__restrict__ ml_off) {
    uint32_t tabm[TK2_CPT], nlm[TK2_CPT];
    const size_t p0 = (size_t)blockIdx.x * TK2_TILE + (size_t)threadIdx.x * (TK2_CPT * 16);
    load_masks(text, n, (size_t)blockIdx.x * TK2_TILE, tabm, nlm);
    Seg tot;
    const Seg pre_local = block_seg_excl(thread_seg(tabm, nlm), &tot);
    const uint2 ts = tile_start[blockIdx.x];
    Seg g; g.nl = ts.x; g.tail = ts.y;
    const Seg st = seg_combine(g, pre_local);          // state just before this thread's first byte
    uint32_t line = st.nl, ord = st.tail;
#pragma unroll
    for (int c = 0; c < TK2_CPT; c++) {
        uint32_t m = tabm[c] | nlm[c];
        const uint32_t nm = nlm[c];
        while (m) {
            const int b = __ffs(m) - 1; m &= m - 1;
            const uint32_t x = (uint32_t)(p0 + (size_t)c * 16 + b);
            if ((nm >> b) & 1u) { nlpos[line] = x; line++; ord = 0; }
            else {
                if (line < n_lines) {
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
    uint2 *tile_seg;
    RC_TRY(T.alloc(&tile_seg, (size_t)ntiles + 1));
    uint32_t n_nl = 0; char last = '\n';
    if (ntiles) {
        LAUNCH(ctx, tk_count_k, ntiles, TK2_T, 0, dtext, nbytes, tile_seg);
        LAUNCH(ctx, tk_tilescan_k, 1, 1024, 0, tile_seg, ntiles);
        CUDA_TRY(cudaMemcpyAsync(&n_nl, &tile_seg[ntiles].x, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(&last, dtext + nbytes - 1, 1, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    }
    const uint32_t n_lines = n_nl + ((nbytes && last != '\n') ? 1 : 0);
    ReadBatch rb;
    rb.text = dtext; rb.nbytes = (uint32_t)nbytes; rb.n = n_lines;
    uint32_t *nlpos, *ftab;
    RC_TRY(T.alloc(&nlpos, n_nl)); RC_TRY(T.alloc(&ftab, (size_t)n_lines * NTAB));
    RC_TRY(T.alloc(&rb.line_off, n_lines)); RC_TRY(T.alloc(&rb.line_len, n_lines)); RC_TRY(T.alloc(&rb.qn_len, n_lines));
    RC_TRY(T.alloc(&rb.flag, n_lines)); RC_TRY(T.alloc(&rb.pos, n_lines));
    RC_TRY(T.alloc(&rb.cig_off, n_lines)); RC_TRY(T.alloc(&rb.cig_len, n_lines));
    RC_TRY(T.alloc(&rb.seq_off, n_lines)); RC_TRY(T.alloc(&rb.seq_len, n_lines));
    RC_TRY(T.alloc(&rb.hash_lo, n_lines)); RC_TRY(T.alloc(&rb.hash_hi, n_lines));
    RC_TRY(T.alloc(&rb.status, n_lines));
    if (want_tags) {
        RC_TRY(T.alloc(&rb.mm_off, n_lines)); RC_TRY(T.alloc(&rb.mm_len, n_lines));
        RC_TRY(T.alloc(&rb.ml_off, n_lines)); RC_TRY(T.alloc(&rb.ml_len, n_lines));
        CUDA_TRY(cudaMemsetAsync(rb.mm_off, 0, (size_t)n_lines * 4, ctx->stream));
        CUDA_TRY(cudaMemsetAsync(rb.ml_off, 0, (size_t)n_lines * 4, ctx->stream));
    }
    if (n_lines) {
        CUDA_TRY(cudaMemsetAsync(ftab, 0xff, (size_t)n_lines * NTAB * 4, ctx->stream));
        LAUNCH(ctx, tk_mark_k, ntiles, TK2_T, 0, dtext, nbytes, tile_seg, n_lines, want_tags ? 1 : 0, nlpos, ftab, rb.mm_off, rb.ml_off);
        LAUNCH(ctx, tk_records_k, grid_for(n_lines, 256), 256, 0, dtext, (uint32_t)nbytes, nlpos, n_nl, n_lines, ftab, want_tags ? 1 : 0, rb);
        LAUNCH_CHECK();
    }
    *out = rb;
    return 0;
}
