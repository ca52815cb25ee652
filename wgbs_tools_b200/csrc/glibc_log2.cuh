// glibc_log2.cuh -- bit-exact device ports of glibc 2.39's log2f(float) and log2(double) as the reference binary
// executes them on an FMA-capable x86-64 host (ifunc variants __log2f_fma / __log2_fma).
//
// Why: segmentor.cpp:130,133 feeds log2f(p) and log2(1.0 - p) into a float/double mix whose last bit decides marginal
// block borders (reference tests/integration/test_segment_integration.py:7-9).  CUDA's own log2f/log2 differ from
// glibc in the last ulp, so we evaluate glibc's algorithm (ARM optimized-routines: table + polynomial in double) with
// the same constants (glibc_log2_tables.cuh, read from libm's bytes) and EXACTLY the operation sequence of the
// compiled FMA variants (which operations are fused was read off `objdump -d libm.so.6`).
// All arithmetic uses explicit IEEE intrinsics so nvcc cannot re-associate or contract anything else.
#pragma once
#include <stdint.h>

#include "glibc_log2_tables.cuh"

// float log2f(float x), x > 0 finite (glibc sysdeps/ieee754/flt-32/e_log2f.c; __log2f_fma @ libm+0x7dd50)
__device__ __forceinline__ float glibc_log2f(float x) {
    uint32_t ix = __float_as_uint(x);
    if (ix == 0x3f800000u) return 0.0f;
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
        // subnormal positive: normalise (x <= 0, inf, nan never reach the segment kernel)
        ix = __float_as_uint(__fmul_rn(x, 8388608.0f));   // 0x1p23f
        ix -= 23u << 23;
    }
    const uint32_t tmp = ix - 0x3f330000u;
    const int i = (tmp >> 19) & 15;
    const uint32_t top = tmp & 0xff800000u;
    const uint32_t iz = ix - top;
    const int k = (int32_t)tmp >> 23;
    const double invc = G_LOG2F_TAB[2 * i], logc = G_LOG2F_TAB[2 * i + 1];
    const double z = (double)__uint_as_float(iz);
    const double r = __fma_rn(z, invc, -1.0);
    const double y0 = __dadd_rn(logc, (double)k);
    const double r2 = __dmul_rn(r, r);
    double y = __fma_rn(G_LOG2F_POLY[1], r, G_LOG2F_POLY[2]);
    y = __fma_rn(G_LOG2F_POLY[0], r2, y);
    const double p = __fma_rn(G_LOG2F_POLY[3], r, y0);
    y = __fma_rn(y, r2, p);
    return __double2float_rn(y);
}

// double log2(double x), x > 0 finite normal (glibc sysdeps/ieee754/dbl-64/e_log2.c; __log2_fma @ libm+0x79f90)
__device__ __forceinline__ double glibc_log2(double x) {
    uint64_t ix = (uint64_t)__double_as_longlong(x);
    const double hi0 = G_LOG2_INVLN2HI, lo0 = G_LOG2_INVLN2LO;
    const uint64_t LO = 0x3feea4af00000000ull;                 // asuint64(1.0 - 0x1.5b51p-5)
    if (ix - LO <= 0x000210a9ffffffffull) {                    // x close to 1
        if (ix == 0x3ff0000000000000ull) return 0.0;
        const double *B = G_LOG2_POLY1;
        const double r = __dsub_rn(x, 1.0);
        const double hi = __dmul_rn(r, hi0);
        const double res = __fma_rn(hi0, r, -hi);
        double lo = __fma_rn(r, lo0, res);
        const double r2 = __dmul_rn(r, r), r4 = __dmul_rn(r2, r2);
        const double q = __fma_rn(r, B[1], B[0]);
        const double y = __fma_rn(q, r2, hi);
        lo = __dadd_rn(__fma_rn(q, r2, __dsub_rn(hi, y)), lo);
        const double a = __fma_rn(__fma_rn(r, B[5], B[4]), r2, __fma_rn(r, B[3], B[2]));
        const double b = __fma_rn(__fma_rn(r, B[9], B[8]), r2, __fma_rn(r, B[7], B[6]));
        const double c = __fma_rn(b, r4, a);
        lo = __fma_rn(c, r4, lo);
        return __dadd_rn(y, lo);
    }
    const uint32_t top = (uint32_t)(ix >> 48);
    if (top - 0x0010u >= 0x7ff0u - 0x0010u) {
        ix = (uint64_t)__double_as_longlong(__dmul_rn(x, 4503599627370496.0));   // 0x1p52: subnormal
        ix -= 52ull << 52;
    }
    const uint64_t tmp = ix - 0x3fe6000000000000ull;           // OFF
    const int i = (int)((tmp >> 46) & 63);
    const int k = (int)((int64_t)tmp >> 52);
    const uint64_t iz = ix - (tmp & (0xfffull << 52));
    const double invc = G_LOG2_TAB[2 * i], logc = G_LOG2_TAB[2 * i + 1];
    const double z = __longlong_as_double((long long)iz);
    const double kd = (double)k;
    const double *A = G_LOG2_POLY;
    const double r = __fma_rn(z, invc, -1.0);
    const double t1 = __dmul_rn(hi0, r);
    const double t2 = __fma_rn(r, lo0, __fma_rn(hi0, r, -t1));
    const double t3 = __dadd_rn(kd, logc);
    const double hi = __dadd_rn(t1, t3);
    const double lo = __dadd_rn(__dadd_rn(__dsub_rn(t3, hi), t1), t2);
    const double r2 = __dmul_rn(r, r), r4 = __dmul_rn(r2, r2);
    const double q01 = __fma_rn(r, A[1], A[0]), q23 = __fma_rn(r, A[3], A[2]), q45 = __fma_rn(r, A[5], A[4]);
    const double p = __fma_rn(q45, r4, __fma_rn(q23, r2, q01));
    return __dadd_rn(__fma_rn(r2, p, lo), hi);
}
