// ctx.cu -- context lifecycle, error string, allocation, staging copies, device-wide scan.
#include <stdarg.h>

#include "common.cuh"

thread_local std::string g_wgbs_err;

int wgbs_set_err(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_wgbs_err = buf;
    return -1;
}

extern "C" const char *wgbs_last_error(void) { return g_wgbs_err.c_str(); }
extern "C" int wgbs_abi_version(void) { return WGBS_B200_ABI_VERSION; }

int wgbs_ctx_activate(wgbs_ctx *ctx) {
    if (!ctx) return wgbs_set_err("null ctx");
    CUDA_TRY(cudaSetDevice(ctx->device));
    return 0;
}

extern "C" wgbs_ctx *wgbs_create(int device, void *stream) {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        wgbs_set_err("wgbs_create: no usable CUDA device (%s); this library has no CPU path", cudaGetErrorString(e));
        return nullptr;
    }
    if (device < 0 || device >= ndev) { wgbs_set_err("wgbs_create: device %d out of range (have %d)", device, ndev); return nullptr; }
    if (cudaSetDevice(device) != cudaSuccess) { wgbs_set_err("cudaSetDevice(%d) failed", device); return nullptr; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { wgbs_set_err("cudaGetDeviceProperties failed"); return nullptr; }
    if (prop.major < 10) {
        wgbs_set_err("wgbs_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
        return nullptr;
    }
    wgbs_ctx *ctx = new wgbs_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    if (stream) { ctx->stream = (cudaStream_t)stream; ctx->own_stream = false; }
    else {
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; wgbs_set_err("cudaStreamCreate failed"); return nullptr; }
        ctx->own_stream = true;
    }
    // keep freed blocks in the pool: allocation becomes a pointer bump after warm-up
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    if (cudaMalloc(&ctx->d_flags, 64 * sizeof(uint32_t)) != cudaSuccess) { delete ctx; wgbs_set_err("cudaMalloc failed"); return nullptr; }
    cudaMemsetAsync(ctx->d_flags, 0, 64 * sizeof(uint32_t), ctx->stream);
    return ctx;
}

extern "C" void wgbs_destroy(wgbs_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (int i = 0; i < 2; i++) if (ctx->pin[i]) cudaFreeHost(ctx->pin[i]);
    if (ctx->d_flags) cudaFree(ctx->d_flags);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" int wgbs_sync(wgbs_ctx *ctx) {
    RC_TRY(wgbs_ctx_activate(ctx));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return 0;
}

void prof_begin(wgbs_ctx *ctx, const char *name) {
    wgbs_ctx::ProfRec r; r.name = name;
    cudaEventCreate(&r.e0); cudaEventCreate(&r.e1);
    cudaEventRecord(r.e0, ctx->stream);
    ctx->prof_recs.push_back(r);
}
void prof_end(wgbs_ctx *ctx) { cudaEventRecord(ctx->prof_recs.back().e1, ctx->stream); }

extern "C" int wgbs_prof_enable(wgbs_ctx *ctx, int on) {
    RC_TRY(wgbs_ctx_activate(ctx));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    for (auto &r : ctx->prof_recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    ctx->prof_recs.clear();
    ctx->prof = on != 0;
    return 0;
}
// "kernel \t launches \t total_ms \n" per kernel since wgbs_prof_enable(ctx, 1); returns bytes written (truncated to cap)
extern "C" int wgbs_prof_report(wgbs_ctx *ctx, char *buf, size_t cap) {
    RC_TRY(wgbs_ctx_activate(ctx));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    std::vector<std::string> names; std::vector<double> ms; std::vector<long> cnt;
    for (auto &r : ctx->prof_recs) {
        float t = 0; cudaEventElapsedTime(&t, r.e0, r.e1);
        size_t k = 0; for (; k < names.size(); k++) if (names[k] == r.name) break;
        if (k == names.size()) { names.push_back(r.name); ms.push_back(0); cnt.push_back(0); }
        ms[k] += t; cnt[k]++;
    }
    std::string out;
    for (size_t k = 0; k < names.size(); k++) { char line[256]; snprintf(line, sizeof line, "%s\t%ld\t%.6f\n", names[k].c_str(), cnt[k], ms[k]); out += line; }
    size_t nb = out.size() < cap ? out.size() : (cap ? cap - 1 : 0);
    if (buf && cap) { memcpy(buf, out.data(), nb); buf[nb] = 0; }
    return (int)nb;
}

extern "C" uint64_t wgbs_launch_count(const wgbs_ctx *ctx) { return ctx ? ctx->launches : 0; }

int dmalloc(wgbs_ctx *ctx, void **p, size_t nbytes) {
    *p = nullptr;
    CUDA_TRY(cudaMallocAsync(p, nbytes ? nbytes : 16, ctx->stream));
    return 0;
}
int dfree(wgbs_ctx *ctx, void *p) {
    if (!p) return 0;
    CUDA_TRY(cudaFreeAsync(p, ctx->stream));
    return 0;
}

extern "C" int wgbs_dev_alloc(wgbs_ctx *ctx, size_t nbytes, void **dptr) {
    RC_TRY(wgbs_ctx_activate(ctx));
    return dmalloc(ctx, dptr, nbytes);
}
extern "C" int wgbs_dev_free(wgbs_ctx *ctx, void *dptr) {
    RC_TRY(wgbs_ctx_activate(ctx));
    return dfree(ctx, dptr);
}

bool is_device_ptr(const void *p) {
    if (!p) return false;
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}
static bool is_pinned_host(const void *p) {
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

static int ensure_pin(wgbs_ctx *ctx, int which, size_t n) {
    if (ctx->pin_cap[which] >= n) return 0;
    if (ctx->pin[which]) { cudaFreeHost(ctx->pin[which]); ctx->pin[which] = nullptr; ctx->pin_cap[which] = 0; }
    CUDA_TRY(cudaMallocHost(&ctx->pin[which], n));
    ctx->pin_cap[which] = n;
    return 0;
}

// Pageable host memory is staged through two pinned 32 MiB buffers so the memcpy into pinned memory overlaps the DMA.
static const size_t STAGE = 32u << 20;

int copy_any(wgbs_ctx *ctx, void *dst, const void *src, size_t nbytes) {
    if (!nbytes) return 0;
    bool sd = is_device_ptr(src), dd = is_device_ptr(dst);
    if (sd && dd) { CUDA_TRY(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDeviceToDevice, ctx->stream)); return 0; }
    if (!sd && !dd) { memcpy(dst, src, nbytes); return 0; }
    if (!sd && dd) {
        if (is_pinned_host(src)) {
            CUDA_TRY(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyHostToDevice, ctx->stream));
            CUDA_TRY(cudaStreamSynchronize(ctx->stream));
            return 0;
        }
        RC_TRY(ensure_pin(ctx, 0, STAGE)); RC_TRY(ensure_pin(ctx, 1, STAGE));
        cudaEvent_t ev[2]; CUDA_TRY(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming)); CUDA_TRY(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
        size_t off = 0; int b = 0; bool used[2] = {false, false};
        while (off < nbytes) {
            size_t c = nbytes - off < STAGE ? nbytes - off : STAGE;
            if (used[b]) cudaEventSynchronize(ev[b]);
            memcpy(ctx->pin[b], (const char *)src + off, c);
            cudaMemcpyAsync((char *)dst + off, ctx->pin[b], c, cudaMemcpyHostToDevice, ctx->stream);
            cudaEventRecord(ev[b], ctx->stream); used[b] = true;
            off += c; b ^= 1;
        }
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        cudaEventDestroy(ev[0]); cudaEventDestroy(ev[1]);
        CUDA_TRY(e);
        return 0;
    }
    // device -> host
    if (is_pinned_host(dst)) {
        CUDA_TRY(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        return 0;
    }
    RC_TRY(ensure_pin(ctx, 0, STAGE)); RC_TRY(ensure_pin(ctx, 1, STAGE));
    {
        size_t off = 0; int b = 0; size_t prev_off = 0, prev_c = 0; int prev_b = -1;
        cudaEvent_t ev[2]; CUDA_TRY(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming)); CUDA_TRY(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
        while (off < nbytes) {
            size_t c = nbytes - off < STAGE ? nbytes - off : STAGE;
            cudaMemcpyAsync(ctx->pin[b], (const char *)src + off, c, cudaMemcpyDeviceToHost, ctx->stream);
            cudaEventRecord(ev[b], ctx->stream);
            if (prev_b >= 0) { cudaEventSynchronize(ev[prev_b]); memcpy((char *)dst + prev_off, ctx->pin[prev_b], prev_c); }
            prev_b = b; prev_off = off; prev_c = c; off += c; b ^= 1;
        }
        cudaError_t e = cudaSuccess;
        if (prev_b >= 0) { e = cudaEventSynchronize(ev[prev_b]); memcpy((char *)dst + prev_off, ctx->pin[prev_b], prev_c); }
        cudaEventDestroy(ev[0]); cudaEventDestroy(ev[1]);
        CUDA_TRY(e);
    }
    return 0;
}

extern "C" int wgbs_memcpy(wgbs_ctx *ctx, void *dst, const void *src, size_t nbytes) {
    RC_TRY(wgbs_ctx_activate(ctx));
    return copy_any(ctx, dst, src, nbytes);
}

int to_device(wgbs_ctx *ctx, const void *p, size_t nbytes, const void **dptr, bool *owned) {
    if (is_device_ptr(p) || nbytes == 0) { *dptr = p; *owned = false; if (!nbytes && !is_device_ptr(p)) { void *q; RC_TRY(dmalloc(ctx, &q, 16)); *dptr = q; *owned = true; } return 0; }
    void *q = nullptr;
    RC_TRY(dmalloc(ctx, &q, nbytes));
    int rc = copy_any(ctx, q, p, nbytes);
    if (rc < 0) { dfree(ctx, q); return rc; }
    *dptr = q; *owned = true;
    return 0;
}

// ------------------------------------------------------------------------------------------------------------------
// device-wide exclusive scan (reduce / scan-of-partials / downsweep). 2048 items per CTA.
// ------------------------------------------------------------------------------------------------------------------
namespace {
constexpr int SCAN_T = 256, SCAN_I = 8, SCAN_TILE = SCAN_T * SCAN_I;

__device__ __forceinline__ uint64_t warp_incl_scan(uint64_t v) {
    const unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint64_t t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}
// block-wide exclusive scan of one value per thread; returns exclusive prefix, *total = block sum
__device__ __forceinline__ uint64_t block_excl_scan(uint64_t v, uint64_t *total) {
    __shared__ uint64_t wsum[32];
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    uint64_t inc = warp_incl_scan(v);
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
        uint64_t s = lane < nw ? wsum[lane] : 0;
        uint64_t si = warp_incl_scan(s);
        wsum[lane] = si - s;
        if (lane == nw - 1) *total = si;
    }
    __syncthreads();
    uint64_t r = wsum[w] + inc - v;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(SCAN_T) scan_reduce_k(const uint32_t *__restrict__ in, size_t n, uint64_t *__restrict__ bsum) {
    size_t base = (size_t)blockIdx.x * SCAN_TILE;
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_I; i++) {
        size_t k = base + (size_t)i * SCAN_T + threadIdx.x;
        if (k < n) s += in[k];
    }
    __shared__ uint64_t tot;
    block_excl_scan(s, &tot);
    if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}
__global__ void __launch_bounds__(1024) scan_partials_k(uint64_t *bsum, size_t nb, uint64_t *total_out) {
    __shared__ uint64_t carry, tot;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (size_t base = 0; base < nb; base += 1024) {
        size_t k = base + threadIdx.x;
        uint64_t v = k < nb ? bsum[k] : 0;
        uint64_t ex = block_excl_scan(v, &tot);
        if (k < nb) bsum[k] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry;
}
template <typename OutT>
__global__ void __launch_bounds__(SCAN_T) scan_down_k(const uint32_t *__restrict__ in, size_t n, const uint64_t *__restrict__ bsum, OutT *__restrict__ out) {
    // thread t owns SCAN_I consecutive items so the tile scan is a thread-serial scan + one block scan
    size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_I;
    uint32_t v[SCAN_I];
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_I; i++) { size_t k = base + i; v[i] = k < n ? in[k] : 0; s += v[i]; }
    __shared__ uint64_t tot;
    uint64_t ex = block_excl_scan(s, &tot) + bsum[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_I; i++) { size_t k = base + i; if (k < n) out[k] = (OutT)ex; ex += v[i]; }
}
template <typename OutT>
__global__ void scan_total_k(const uint64_t *total, OutT *out_n) { *out_n = (OutT)*total; }

template <typename OutT>
int scan_impl(wgbs_ctx *ctx, const uint32_t *in, OutT *out, size_t n) {
    size_t nb = (n + SCAN_TILE - 1) / SCAN_TILE; if (nb == 0) nb = 1;
    Temps T(ctx);
    uint64_t *bsum = nullptr, *total = nullptr;
    RC_TRY(T.alloc(&bsum, nb)); RC_TRY(T.alloc(&total, 1));
    LAUNCH(ctx, scan_reduce_k, (unsigned)nb, SCAN_T, 0, in, n, bsum);
    LAUNCH(ctx, scan_partials_k, 1, 1024, 0, bsum, nb, total);
    LAUNCH(ctx, scan_down_k<OutT>, (unsigned)nb, SCAN_T, 0, in, n, bsum, out);
    LAUNCH(ctx, scan_total_k<OutT>, 1, 1, 0, total, out + n);
    LAUNCH_CHECK();
    return 0;
}
}  // namespace

int scan_u32_u64(wgbs_ctx *ctx, const uint32_t *in, uint64_t *out, size_t n) { return scan_impl<uint64_t>(ctx, in, out, n); }
int scan_u32_u32(wgbs_ctx *ctx, const uint32_t *in, uint32_t *out, size_t n) { return scan_impl<uint32_t>(ctx, in, out, n); }
