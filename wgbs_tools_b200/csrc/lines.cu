// lines.cu -- newline positions of a text buffer (shared by the pat-text and SAM-text front ends).
#include "lines.cuh"

namespace {
// ------------------------------------------------------------------------------------------------------------------
// text -> lines.  One warp-iteration covers 512 contiguous bytes (16 B per lane); a CTA tile is 8 warps x 4 iters.
// ------------------------------------------------------------------------------------------------------------------
constexpr int NL_T = 256, NL_ITERS = 4, NL_TILE = NL_T * 16 * NL_ITERS;  // 16 KiB

__global__ void __launch_bounds__(NL_T) nl_count_k(const char *__restrict__ text, size_t n, uint32_t *__restrict__ bcount) {
    const size_t tile0 = (size_t)blockIdx.x * NL_TILE;
    const unsigned w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t c = 0;
#pragma unroll
    for (int it = 0; it < NL_ITERS; it++) {
        size_t pos = tile0 + ((size_t)(w * NL_ITERS + it) * 32 + lane) * 16;
        if (pos < n) c += __popc(eq_mask16(load16_guard(text, pos, n), '\n'));
    }
    c = __reduce_add_sync(0xffffffffu, c);
    __shared__ uint32_t ws[NL_T / 32];
    if (lane == 0) ws[w] = c;
    __syncthreads();
    if (threadIdx.x == 0) { uint32_t s = 0; for (int i = 0; i < NL_T / 32; i++) s += ws[i]; bcount[blockIdx.x] = s; }
}

__global__ void __launch_bounds__(NL_T) nl_write_k(const char *__restrict__ text, size_t n, const uint32_t *__restrict__ boff,
                                                    uint32_t *__restrict__ nlpos) {
    const size_t tile0 = (size_t)blockIdx.x * NL_TILE;
    const unsigned w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t masks[NL_ITERS], cnt[NL_ITERS], wtot = 0;
#pragma unroll
    for (int it = 0; it < NL_ITERS; it++) {
        size_t pos = tile0 + ((size_t)(w * NL_ITERS + it) * 32 + lane) * 16;
        masks[it] = pos < n ? eq_mask16(load16_guard(text, pos, n), '\n') : 0;
        cnt[it] = __popc(masks[it]);
        wtot += cnt[it];
    }
    wtot = __reduce_add_sync(0xffffffffu, wtot);
    __shared__ uint32_t ws[NL_T / 32];
    if (lane == 0) ws[w] = wtot;
    __syncthreads();
    uint32_t base = boff[blockIdx.x];
    for (unsigned i = 0; i < w; i++) base += ws[i];
#pragma unroll
    for (int it = 0; it < NL_ITERS; it++) {
        uint32_t inc = cnt[it];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= (unsigned)d) inc += t; }
        uint32_t o = base + inc - cnt[it];
        size_t pos = tile0 + ((size_t)(w * NL_ITERS + it) * 32 + lane) * 16;
        uint32_t m = masks[it];
        while (m) { int b = __ffs(m) - 1; m &= m - 1; nlpos[o++] = (uint32_t)(pos + b); }
        base += __shfl_sync(0xffffffffu, inc, 31);
    }
}


}  // namespace

// nlpos[j] = byte offset of the j-th '\n'.  line i = [i ? nlpos[i-1]+1 : 0, i < n_nl ? nlpos[i] : nbytes).
int find_lines(wgbs_ctx *ctx, const char *dtext, size_t nbytes, Temps &T, uint32_t **nlpos_out, uint32_t *n_nl_out, uint32_t *n_lines_out) {
    unsigned ntiles = (unsigned)((nbytes + NL_TILE - 1) / NL_TILE); if (!ntiles) ntiles = 1;
    uint32_t *bcount, *boff, *nlpos;
    RC_TRY(T.alloc(&bcount, ntiles)); RC_TRY(T.alloc(&boff, ntiles + 1));
    LAUNCH(ctx, nl_count_k, ntiles, NL_T, 0, dtext, nbytes, bcount);
    RC_TRY(scan_u32_u32(ctx, bcount, boff, ntiles));
    uint32_t n_nl = 0; char last = '\n';
    CUDA_TRY(cudaMemcpyAsync(&n_nl, boff + ntiles, 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (nbytes) CUDA_TRY(cudaMemcpyAsync(&last, dtext + nbytes - 1, 1, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    RC_TRY(T.alloc(&nlpos, n_nl));
    LAUNCH(ctx, nl_write_k, ntiles, NL_T, 0, dtext, nbytes, boff, nlpos);
    LAUNCH_CHECK();
    *nlpos_out = nlpos; *n_nl_out = n_nl; *n_lines_out = n_nl + ((nbytes && last != '\n') ? 1 : 0);
    return 0;
}
