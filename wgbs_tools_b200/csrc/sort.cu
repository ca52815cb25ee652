// sort.cu -- stable LSD radix sort of (uint32 key, uint32 value) pairs, 8-bit digits, with digit-pass skipping.
//
// Used by: template pairing (sort records by 64-bit QNAME hash = two calls) and the collapse step, which replaces the
// reference's `sort -k2,2n -k3,3` (python/bam2pat.py:99) by a sequence of these sorts over successive key words.
//
// A pre-pass builds the global histogram of all four digits at once; a digit whose histogram has a single non-empty bin
// is skipped (this is what makes sorting zero-padded pattern words cheap).  Every remaining pass is one kernel that reads
// the keys once: warp match_any ranking + decoupled look-back for the tile offsets.
#include "common.cuh"
#include "sort.cuh"

namespace {

constexpr int RS_T = 512;             // threads per CTA
constexpr int RS_IPT = 16;            // items per thread
constexpr int RS_TILE = RS_T * RS_IPT;
constexpr int RS_WARPS = RS_T / 32;

__global__ void __launch_bounds__(RS_T) rs_global_hist_k(const uint32_t *__restrict__ keys, size_t n, uint32_t *__restrict__ ghist /*[4][256]*/) {
    __shared__ uint32_t h[4 * 256];
    for (int i = threadIdx.x; i < 1024; i += RS_T) h[i] = 0;
    __syncthreads();
    for (size_t i = (size_t)blockIdx.x * RS_T + threadIdx.x; i < n; i += (size_t)gridDim.x * RS_T) {
        uint32_t k = keys[i];
        atomicAdd(&h[k & 255], 1u); atomicAdd(&h[256 + ((k >> 8) & 255)], 1u);
        atomicAdd(&h[512 + ((k >> 16) & 255)], 1u); atomicAdd(&h[768 + (k >> 24)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 1024; i += RS_T) if (h[i]) atomicAdd(&ghist[i], h[i]);
}

// One pass = ONE kernel ("onesweep" style): per-tile digit counts are published and the exclusive offsets of a tile are
// obtained by looking back over the preceding tiles' published words (decoupled look-back, one thread per digit), so the
// keys are read exactly once per pass.  status[tile*256 + d] = flag(2 bits) << 30 | count : 1 = this tile's count,
// 2 = inclusive count over tiles 0..tile.  Tiles are handed out by an atomic ticket (forward progress).
constexpr uint32_t RS_AGG = 1u << 30, RS_INC = 2u << 30, RS_VAL = (1u << 30) - 1;

__global__ void __launch_bounds__(RS_T) rs_onesweep_k(const uint32_t *__restrict__ kin, const uint32_t *__restrict__ vin, size_t n, int shift,
                                                       const uint32_t *__restrict__ ghist /*[256] of this digit*/, uint32_t *__restrict__ status,
                                                       unsigned int *__restrict__ ticket, uint32_t *__restrict__ kout, uint32_t *__restrict__ vout) {
    __shared__ uint32_t cnt[RS_WARPS][256];   // per-warp digit counts -> per-warp output bases
    __shared__ uint32_t gbase[256];           // exclusive scan of the global histogram
    __shared__ uint32_t wtot[RS_WARPS];
    __shared__ unsigned s_tile;
    const unsigned w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_T) (&cnt[0][0])[i] = 0;
    // exclusive scan of the 256-bin global histogram (digit base offsets)
    {
        uint32_t h = threadIdx.x < 256 ? ghist[threadIdx.x] : 0, inc = h;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= (unsigned)d) inc += t; }
        if (lane == 31 && w < 8) wtot[w] = inc;
        __syncthreads();
        if (threadIdx.x < 256) {
            uint32_t off = 0;
            for (unsigned i = 0; i < w; i++) off += wtot[i];
            gbase[threadIdx.x] = off + inc - h;
        }
    }
    __syncthreads();
    const unsigned tile = s_tile;
    const size_t wbase = (size_t)tile * RS_TILE + (size_t)w * (RS_IPT * 32);
    uint32_t key[RS_IPT], val[RS_IPT];
    uint16_t rank[RS_IPT];
#pragma unroll
    for (int j = 0; j < RS_IPT; j++) {
        size_t i = wbase + (size_t)j * 32 + lane;
        bool ok = i < n;
        key[j] = ok ? kin[i] : 0xffffffffu;
        val[j] = ok ? vin[i] : 0;
    }
#pragma unroll
    for (int j = 0; j < RS_IPT; j++) {
        size_t i = wbase + (size_t)j * 32 + lane;
        bool ok = i < n;
        uint32_t d = ok ? ((key[j] >> shift) & 255) : 256;          // 256: inactive lanes form their own group
        uint32_t peers = __match_any_sync(0xffffffffu, d);
        uint32_t below = __popc(peers & ((1u << lane) - 1));
        int leader = __ffs(peers) - 1;
        uint32_t b = 0;
        if (ok && (int)lane == leader) { b = cnt[w][d]; cnt[w][d] = b + __popc(peers); }
        b = __shfl_sync(0xffffffffu, b, leader);
        rank[j] = (uint16_t)(b + below);
        __syncwarp();
    }
    __syncthreads();
    if (threadIdx.x < 256) {
        // thread d owns digit d: tile count, publish, look back, per-warp bases
        const uint32_t d = threadIdx.x;
        uint32_t tc = 0;
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ww++) tc += cnt[ww][d];
        volatile uint32_t *st = status;
        uint32_t excl = 0;
        if (tile == 0) st[d] = RS_INC | tc;
        else {
            st[(size_t)tile * 256 + d] = RS_AGG | tc;
            long long t = (long long)tile - 1;
            while (true) {
                uint32_t x;
                do { x = st[(size_t)t * 256 + d]; } while ((x >> 30) == 0);
                excl += x & RS_VAL;
                if ((x >> 30) == 2) break;
                t--;
            }
            st[(size_t)tile * 256 + d] = RS_INC | ((excl + tc) & RS_VAL);
        }
        uint32_t run = gbase[d] + excl;
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ww++) { uint32_t c = cnt[ww][d]; cnt[ww][d] = run; run += c; }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < RS_IPT; j++) {
        size_t i = wbase + (size_t)j * 32 + lane;
        if (i < n) {
            uint32_t d = (key[j] >> shift) & 255;
            uint32_t o = cnt[w][d] + rank[j];
            kout[o] = key[j];
            vout[o] = val[j];
        }
    }
}

__global__ void iota_k(uint32_t *p, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (uint32_t)i;
}

}  // namespace

int fill_iota(wgbs_ctx *ctx, uint32_t *p, size_t n) {
    if (n) { LAUNCH(ctx, iota_k, grid_for(n, 256), 256, 0, p, n); LAUNCH_CHECK(); }
    return 0;
}

// Sorts in place in the sense that on return *keys / *vals point at the buffers holding the sorted data
// (either the original pair or the alt pair).  Stable.  n < 2^32.
int radix_sort_pairs(wgbs_ctx *ctx, uint32_t **keys, uint32_t **vals, uint32_t **keys_alt, uint32_t **vals_alt, size_t n) {
    if (n < 2) return 0;
    if (n >= (1ull << 30)) return wgbs_set_err("radix_sort_pairs: n must be < 2^30 per call");
    Temps T(ctx);
    uint32_t *ghist = nullptr;
    RC_TRY(T.alloc(&ghist, 1024));
    CUDA_TRY(cudaMemsetAsync(ghist, 0, 1024 * 4, ctx->stream));
    unsigned hb = (unsigned)((n + RS_T * 8 - 1) / (RS_T * 8)); if (hb > (unsigned)ctx->sm_count * 8) hb = ctx->sm_count * 8;
    LAUNCH(ctx, rs_global_hist_k, hb, RS_T, 0, *keys, n, ghist);
    uint32_t hh[1024];
    CUDA_TRY(cudaMemcpyAsync(hh, ghist, sizeof hh, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    const uint32_t ntiles = (uint32_t)((n + RS_TILE - 1) / RS_TILE);
    uint32_t *status = nullptr;
    int active = 0;
    bool act[4];
    for (int p = 0; p < 4; p++) {
        act[p] = true;
        for (int b = 0; b < 256; b++) if (hh[p * 256 + b] == (uint32_t)n) { act[p] = false; break; }
        active += act[p];
    }
    if (!active) return 0;
    // one zeroed status table (+ ticket) per active pass, cleared with a single memset
    const size_t per = (size_t)ntiles * 256 + 4;
    RC_TRY(T.alloc(&status, per * active));
    CUDA_TRY(cudaMemsetAsync(status, 0, per * active * 4, ctx->stream));
    int q = 0;
    for (int p = 0; p < 4; p++) {
        if (!act[p]) continue;
        uint32_t *st = status + per * q++;
        LAUNCH(ctx, rs_onesweep_k, ntiles, RS_T, 0, *keys, *vals, n, p * 8, ghist + p * 256, st, (unsigned int *)(st + (size_t)ntiles * 256), *keys_alt, *vals_alt);
        uint32_t *t = *keys; *keys = *keys_alt; *keys_alt = t;
        t = *vals; *vals = *vals_alt; *vals_alt = t;
    }
    LAUNCH_CHECK();
    return 0;
}

extern "C" int wgbs_sort_pairs_u32(wgbs_ctx *ctx, uint32_t *keys, uint32_t *vals, size_t n) {
    RC_TRY(wgbs_ctx_activate(ctx));
    if (!is_device_ptr(keys) || !is_device_ptr(vals)) return wgbs_set_err("wgbs_sort_pairs_u32: device pointers required");
    Temps T(ctx);
    uint32_t *ka, *va;
    RC_TRY(T.alloc(&ka, n)); RC_TRY(T.alloc(&va, n));
    uint32_t *k = keys, *v = vals, *k2 = ka, *v2 = va;
    RC_TRY(radix_sort_pairs(ctx, &k, &v, &k2, &v2, n));
    if (k != keys) {
        CUDA_TRY(cudaMemcpyAsync(keys, k, n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(vals, v, n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    return 0;
}
