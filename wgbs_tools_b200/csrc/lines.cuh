// lines.cuh -- text -> line table
#pragma once
#include "common.cuh"

int find_lines(wgbs_ctx *ctx, const char *dtext, size_t nbytes, Temps &T, uint32_t **nlpos, uint32_t *n_nl, uint32_t *n_lines);

// 16-byte chunk helpers shared by the tokenizers
__device__ __forceinline__ uint4 load16_guard(const char *__restrict__ text, size_t pos, size_t n) {
    if (pos + 16 <= n && ((uintptr_t)(text + pos) & 15) == 0) return *reinterpret_cast<const uint4 *>(text + pos);
    uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
    for (int b = 0; b < 16; b++)
        if (pos + b < n) w[b >> 2] |= (uint32_t)(uint8_t)text[pos + b] << ((b & 3) * 8);
    return make_uint4(w[0], w[1], w[2], w[3]);
}
// bit b set iff byte b of the 16-byte chunk equals c.  Per word: SIMD byte compare (0xff per equal byte), keep bit 0 of
// every byte, and gather the four bits into the top nibble with one multiply (2^24 + 2^17 + 2^10 + 2^3).
__device__ __forceinline__ uint32_t eq_mask16(uint4 v, uint32_t c) {
    const uint32_t rep = c * 0x01010101u;
    const uint32_t m0 = ((__vcmpeq4(v.x, rep) & 0x01010101u) * 0x01020408u) >> 24;
    const uint32_t m1 = ((__vcmpeq4(v.y, rep) & 0x01010101u) * 0x01020408u) >> 24;
    const uint32_t m2 = ((__vcmpeq4(v.z, rep) & 0x01010101u) * 0x01020408u) >> 24;
    const uint32_t m3 = ((__vcmpeq4(v.w, rep) & 0x01010101u) * 0x01020408u) >> 24;
    return m0 | (m1 << 4) | (m2 << 8) | (m3 << 12);
}
