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
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
    if (ctx->ev_copy) cudaEventDestroy(ctx->ev_copy);
    if (ctx->ev_comp) cudaEventDestroy(ctx->ev_comp);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" int wgbs_sync(wgbs_ctx *ctx) {
    RC_TRY(wgbs_ctx_activate(ctx));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// Streaming inputs: start the host->device copy of the NEXT batch on the context's copy stream while the kernels of the
// current batch run (PCIe is the bottleneck of the end-to-end path: 357 MB of SAM text per 1M reads).  The copy is ordered
// after everything already enqueued on the compute stream (the destination buffer may still be in use by it);
// wgbs_prefetch_wait makes the compute stream wait for the last prefetch.  host_src should be pinned memory
// (cudaHostAlloc / torch pin_memory): a pageable source makes cudaMemcpyAsync synchronous.
extern "C" int wgbs_prefetch(wgbs_ctx *ctx, void *dev_dst, const void *host_src, size_t nbytes) {
    RC_TRY(wgbs_ctx_activate(ctx));
    if (!dev_dst || !host_src) return wgbs_set_err("wgbs_prefetch: null argument");
    if (!is_device_ptr(dev_dst) || is_device_ptr(host_src)) return wgbs_set_err("wgbs_prefetch: dst must be device memory, src host memory");
    if (!ctx->copy_stream) {
        CUDA_TRY(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_copy, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_comp, cudaEventDisableTiming));
    }
    CUDA_TRY(cudaEventRecord(ctx->ev_comp, ctx->stream));
    CUDA_TRY(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_comp, 0));
    CUDA_TRY(cudaMemcpyAsync(dev_dst, host_src, nbytes, cudaMemcpyHostToDevice, ctx->copy_stream));
    CUDA_TRY(cudaEventRecord(ctx->ev_copy, ctx->copy_stream));
    return 0;
}
extern "C" int wgbs_prefetch_wait(wgbs_ctx *ctx) {
    RC_TRY(wgbs_ctx_activate(ctx));
    if (ctx->ev_copy) CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ctx->ev_copy, 0));
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

int ctx_scratch(wgbs_ctx *ctx, size_t nbytes, void **p) {
    if (nbytes > ctx->scratch_cap) {
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));              // kernels of earlier calls may still use the old block
        if (ctx->scratch) { cudaFree(ctx->scratch); ctx->scratch = nullptr; ctx->scratch_cap = 0; }
        const size_t cap = nbytes + nbytes / 8 + (1u << 20);
        CUDA_TRY(cudaMalloc(&ctx->scratch, cap));
        ctx->scratch_cap = cap;
    }
    *p = ctx->scratch;
    return 0;
}

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
// device-wide exclusive scan: ONE kernel, single pass over the data (decoupled look-back).  4096 items per CTA; tiles are
// handed out by an atomic ticket so a CTA only ever waits on tiles that are already resident or finished.
// status[tile] = flag(2 bits) << 62 | value : flag 1 = tile aggregate, 2 = inclusive prefix up to and including the tile.
// ------------------------------------------------------------------------------------------------------------------
namespace {
constexpr int SCAN_T = 256, SCAN_I = 16, SCAN_TILE = SCAN_T * SCAN_I;
constexpr unsigned long long ST_AGG = 1ull << 62, ST_INC = 2ull << 62, ST_VAL = (1ull << 62) - 1;

__device__ __forceinline__ uint64_t warp_incl_scan(uint64_t v) {
    const unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint64_t t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

template <typename OutT>
__global__ void __launch_bounds__(SCAN_T) scan_lookback_k(const uint32_t *__restrict__ in, size_t n, OutT *__restrict__ out,
                                                          unsigned long long *__restrict__ status, unsigned int *__restrict__ ticket) {
    __shared__ uint64_t wsum[SCAN_T / 32];
    __shared__ uint64_t s_prefix;
    __shared__ unsigned s_tile;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const unsigned tile = s_tile;
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const size_t base = (size_t)tile * SCAN_TILE + (size_t)threadIdx.x * SCAN_I;
    uint32_t v[SCAN_I];
    uint64_t s = 0;
    const bool vec_in = (((uintptr_t)in) & 15) == 0, vec_out = (((uintptr_t)out) & 15) == 0;
    if (base + SCAN_I <= n && vec_in) {
        const uint4 *p4 = reinterpret_cast<const uint4 *>(in + base);
#pragma unroll
        for (int i = 0; i < SCAN_I / 4; i++) { uint4 q = p4[i]; v[4 * i] = q.x; v[4 * i + 1] = q.y; v[4 * i + 2] = q.z; v[4 * i + 3] = q.w; }
    } else {
#pragma unroll
        for (int i = 0; i < SCAN_I; i++) v[i] = base + i < n ? in[base + i] : 0;
    }
#pragma unroll
    for (int i = 0; i < SCAN_I; i++) s += v[i];
    const uint64_t inc = warp_incl_scan(s);
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    uint64_t wexcl = 0, total = 0;
#pragma unroll
    for (int i = 0; i < SCAN_T / 32; i++) { uint64_t x = wsum[i]; if (i < (int)w) wexcl += x; total += x; }
    if (w == 0) {
        volatile unsigned long long *st = status;
        uint64_t prefix = 0;
        if (tile == 0) { if (lane == 0) st[0] = ST_INC | total; }
        else {
            if (lane == 0) st[tile] = ST_AGG | total;
            long long idx = (long long)tile - 1;
            while (true) {
                const long long j = idx - lane;
                unsigned long long x = ST_INC;                         // before tile 0: inclusive prefix 0
                if (j >= 0) x = st[j];
                // every lane must hold a published word before we decide (all lanes take part in the vote)
                while (__any_sync(0xffffffffu, (x >> 62) == 0)) { if ((x >> 62) == 0) x = st[j]; }
                const unsigned incm = __ballot_sync(0xffffffffu, (x >> 62) == 2);
                uint64_t val = x & ST_VAL;
                if (incm) {
                    const int L = __ffs(incm) - 1;                     // closest predecessor holding an inclusive prefix
                    if ((int)lane > L) val = 0;
                }
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) val += __shfl_xor_sync(0xffffffffu, val, d);
                prefix += val;
                if (incm) break;
                idx -= 32;
            }
            if (lane == 0) st[tile] = ST_INC | ((prefix + total) & ST_VAL);
        }
        if (lane == 0) s_prefix = prefix;
    }
    __syncthreads();
    uint64_t ex = s_prefix + wexcl + inc - s;
    if (base + SCAN_I <= n && sizeof(OutT) == 4 && vec_out) {
        uint32_t o[SCAN_I];
#pragma unroll
        for (int i = 0; i < SCAN_I; i++) { o[i] = (uint32_t)ex; ex += v[i]; }
        uint4 *q4 = reinterpret_cast<uint4 *>(out + base);
#pragma unroll
        for (int i = 0; i < SCAN_I / 4; i++) q4[i] = make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
    } else {
#pragma unroll
        for (int i = 0; i < SCAN_I; i++) { if (base + i < n) out[base + i] = (OutT)ex; ex += v[i]; }
    }
    if ((size_t)(tile + 1) * SCAN_TILE >= n && threadIdx.x == SCAN_T - 1) {
        // last tile: the grand total goes to out[n]
        out[n] = (OutT)(s_prefix + total);
    }
}

template <typename OutT>
int scan_impl(wgbs_ctx *ctx, const uint32_t *in, OutT *out, size_t n) {
    size_t nb = (n + SCAN_TILE - 1) / SCAN_TILE; if (nb == 0) nb = 1;
    Temps T(ctx);
    unsigned long long *status = nullptr;
    RC_TRY(T.alloc(&status, nb + 1));
    CUDA_TRY(cudaMemsetAsync(status, 0, (nb + 1) * 8, ctx->stream));
    LAUNCH(ctx, scan_lookback_k<OutT>, (unsigned)nb, SCAN_T, 0, in, n, out, status, (unsigned int *)(status + nb));
    LAUNCH_CHECK();
    return 0;
}
}  // namespace

int scan_u32_u64(wgbs_ctx *ctx, const uint32_t *in, uint64_t *out, size_t n) { return scan_impl<uint64_t>(ctx, in, out, n); }
int scan_u32_u32(wgbs_ctx *ctx, const uint32_t *in, uint32_t *out, size_t n) { return scan_impl<uint32_t>(ctx, in, out, n); }
