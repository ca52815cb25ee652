// bgzf.cuh -- BGZF block table shared by bamdev.cu (framing, C ABI) and inflate2.cu (the decoder kernels)
#pragma once
#include "common.cuh"

// one BGZF block: coff = first byte of its deflate payload in the compressed buffer, uoff = where its output goes
struct BgzfBlock { uint64_t coff, uoff; uint32_t clen, usize, crc, tok /* first token slot (inflate2.cu) */; };

// Two-phase decoder (inflate2.cu).  plan: fills blocks[i].tok on the HOST table (before it is uploaded) and returns the number of
// token slots the launch needs.  launch: inflates blocks[0..nb) (device table) from d_comp into out; d_comp must be readable for
// 64 bytes past its end; d_err: device word preset to ~0, a failing block stores (index << 8 | -code) with atomicMin.
uint64_t bgzf_inflate2_plan(BgzfBlock *h_blocks, uint32_t nb);
int bgzf_inflate2_launch(wgbs_ctx *ctx, const uint8_t *d_comp, const BgzfBlock *d_blocks, uint32_t nb, uint64_t token_slots, uint8_t *out,
                         unsigned long long *d_err);
// the round-1 decoder on every block (WGBS_INFLATE=2: the yardstick)
int bgzf_inflate_warp_launch(wgbs_ctx *ctx, const uint8_t *d_comp, const BgzfBlock *d_blocks, uint32_t nb, uint8_t *out, unsigned long long *d_err);
