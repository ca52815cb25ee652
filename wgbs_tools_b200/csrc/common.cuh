// common.cuh -- context, error handling, stream-ordered scratch, host<->device staging, device-wide scan.
// Part of libwgbs_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/wgbs_b200.h"

#define WGBS_SM_COUNT_FALLBACK 148

extern thread_local std::string g_wgbs_err;
int wgbs_set_err(const char *fmt, ...);

#define CUDA_TRY(expr)                                                                                     \
    do {                                                                                                   \
        cudaError_t _e = (expr);                                                                           \
        if (_e != cudaSuccess) {                                                                           \
            return wgbs_set_err("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));     \
        }                                                                                                  \
    } while (0)
#define RC_TRY(expr)          \
    do {                      \
        int _rc = (expr);     \
        if (_rc < 0) return _rc; \
    } while (0)

struct wgbs_ctx {
    int device = 0;
    int sm_count = WGBS_SM_COUNT_FALLBACK;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    uint64_t launches = 0;
    // pinned staging buffers (grown on demand)
    void *pin[2] = {nullptr, nullptr};
    size_t pin_cap[2] = {0, 0};
    // second stream for wgbs_prefetch: host->device copies that overlap the kernels of the previous batch
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_copy = nullptr, ev_comp = nullptr;
    // small device scratch for flags / counters
    uint32_t *d_flags = nullptr;  // 64 words
    // large device scratch that outlives a call (grown on demand, plain cudaMalloc): see ctx_scratch
    void *scratch = nullptr; size_t scratch_cap = 0;
    // optional per-kernel timing (wgbs_prof_enable): one event pair per launch, aggregated by kernel name
    bool prof = false;
    struct ProfRec { const char *name; cudaEvent_t e0, e1; };
    std::vector<ProfRec> prof_recs;
};
void prof_begin(wgbs_ctx *ctx, const char *name);
void prof_end(wgbs_ctx *ctx);

// every kernel launch goes through this macro so ctx->launches is the number of OUR kernels launched
#define LAUNCH(ctx, kern, grid, block, smem, ...)                              \
    do {                                                                       \
        if ((ctx)->prof) prof_begin((ctx), #kern);                             \
        kern<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);         \
        if ((ctx)->prof) prof_end((ctx));                                      \
        (ctx)->launches++;                                                     \
    } while (0)
#define LAUNCH_CHECK() CUDA_TRY(cudaGetLastError())

int wgbs_ctx_activate(wgbs_ctx *ctx);
// at least nbytes of device memory owned by the context, valid until the next ctx_scratch call on it (calls on one context are
// serial and stream-ordered, so one call's kernels are done with it before the next call's kernels run)
int ctx_scratch(wgbs_ctx *ctx, size_t nbytes, void **p);
// stream-ordered allocation
int dmalloc(wgbs_ctx *ctx, void **p, size_t nbytes);
int dfree(wgbs_ctx *ctx, void *p);
template <typename T>
static inline int dalloc(wgbs_ctx *ctx, T **p, size_t n) { return dmalloc(ctx, (void **)p, (n ? n : 1) * sizeof(T)); }

bool is_device_ptr(const void *p);
// copy host->device / device->host / d2d on ctx stream; host side staged through pinned memory; returns after
// the data is safe to reuse on the host side.
int copy_any(wgbs_ctx *ctx, void *dst, const void *src, size_t nbytes);
// make `p` (host or device) available on the device. *owned is set when a temporary device copy was made.
int to_device(wgbs_ctx *ctx, const void *p, size_t nbytes, const void **dptr, bool *owned);

// RAII-less helper: list of temporaries freed at scope end
struct Temps {
    wgbs_ctx *ctx;
    std::vector<void *> v;
    explicit Temps(wgbs_ctx *c) : ctx(c) {}
    ~Temps() { for (void *p : v) dfree(ctx, p); }
    template <typename T>
    int alloc(T **p, size_t n) { int rc = dalloc(ctx, p, n); if (rc == 0) v.push_back((void *)*p); return rc; }
    void keep(void *p) { for (auto &q : v) if (q == p) { q = v.back(); v.pop_back(); return; } }
};

// device-wide exclusive scan of uint32 values into uint64 offsets (n+1 outputs: out[n] = total). in/out are device.
int scan_u32_u64(wgbs_ctx *ctx, const uint32_t *in, uint64_t *out, size_t n);
// same, uint32 outputs (caller guarantees total < 2^32); total (device uint64*) may be null
int scan_u32_u32(wgbs_ctx *ctx, const uint32_t *in, uint32_t *out, size_t n);

static inline unsigned grid_for(size_t n, unsigned block, unsigned per_thread = 1) {
    size_t g = (n + (size_t)block * per_thread - 1) / ((size_t)block * per_thread);
    if (g < 1) g = 1;
    if (g > 0x7fffffffULL) g = 0x7fffffffULL;
    return (unsigned)g;
}
