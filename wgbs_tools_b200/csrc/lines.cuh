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
// bit b set iff byte b of the 16-byte chunk equals c
__device__ __forceinline__ uint32_t eq_mask16(uint4 v, uint32_t c) {
    const uint32_t rep = c * 0x01010101u;
    uint32_t m = 0;
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        uint32_t x = w[i] ^ rep;                                   // zero byte where equal
        uint32_t z = ~(((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x | 0x7f7f7f7fu);  // 0x80 in each zero byte
        m |= (((z >> 7) & 1u) | ((z >> 14) & 2u) | ((z >> 21) & 4u) | ((z >> 28) & 8u)) << (4 * i);
    }
    return m;
}
