// sort.cuh -- internal interface of sort.cu
#pragma once
#include "common.cuh"

int fill_iota(wgbs_ctx *ctx, uint32_t *p, size_t n);
// stable LSD radix sort of (key,val) u32 pairs; on return *keys/*vals point at the sorted buffers (ping-pong with alt)
int radix_sort_pairs(wgbs_ctx *ctx, uint32_t **keys, uint32_t **vals, uint32_t **keys_alt, uint32_t **vals_alt, size_t n);
