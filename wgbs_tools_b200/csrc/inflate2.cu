// inflate2.cu -- the kernels of the two-phase BGZF decoder (inflate2_core.cuh, inflate3_core.cuh):
//
//   bgzf_team_decode_k   one WARP per BGZF block, 32 blocks per CTA and SM (7.2 KB of shared memory each): the leader lane parses the
//                        deflate block header, the team builds the two-level Huffman tables, then the lanes walk spans of the bit stream
//                        and verify each other; literals go to their final place, matches to a token list.
//   bgzf_resolve_k       one WARP per block: token replay (LZ77 copies, lane = output byte), then ISIZE and CRC32 of the block.
//   bgzf_warp_inflate_k  the round-1 decoder (one warp per block, every lane the same walk): blocks whose tables exceed the arena, and the
//                        yardstick (WGBS_INFLATE=2).
//
// Stands in for the zlib inflate inside `samtools view` (reference src/python/bam2pat.py:165).
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "bgzf.cuh"
#include "inflate2_core.cuh"
#include "inflate3_core.cuh"

namespace {

// slicing-by-4 tables of the CRC-32 (dflate2::crc_slice_entry), built by the compiler and read through L1: bgzf_resolve_k holds no shared
// memory at all, so its CTAs fit on an SM beside the team decoder's (230 KB) when several batches are in flight
struct Crc4Table { uint32_t v[1024]; };
constexpr uint32_t crc_byte_entry_cx(uint32_t i) { uint32_t c = i; for (int k = 0; k < 8; k++) c = (c & 1) ? (c >> 1) ^ 0xEDB88320u : c >> 1; return c; }
constexpr Crc4Table make_crc4() {
    Crc4Table t{};
    for (uint32_t i = 0; i < 256; i++) t.v[i] = crc_byte_entry_cx(i);
    for (uint32_t k = 1; k < 4; k++) for (uint32_t i = 0; i < 256; i++) t.v[k * 256 + i] = t.v[t.v[(k - 1) * 256 + i] & 0xff] ^ (t.v[(k - 1) * 256 + i] >> 8);
    return t;
}
__device__ const Crc4Table g_crc4 = make_crc4();

constexpr int RES_WARPS = 8;            // warps (= blocks) per CTA of bgzf_resolve_k
constexpr int OLD_WARPS = 4;            // warps (= blocks) per CTA of bgzf_warp_inflate_k
constexpr uint32_t CHUNK_BLOCKS = 8192; // blocks per launch pair: bounds the token scratch (~175 KB per block) to 1.4 GB

// Team decoder (inflate3_core.cuh): S lanes walk ONE block together (spans of its bit stream, verified against each other), TEAM_SLOTS
// blocks per CTA and SM -- the block's tables and the leader's header decoder in 7.2 KB of shared memory each.
constexpr int TEAM_SLOTS = 32;
static_assert(TEAM_SLOTS * sizeof(dflate3::TeamMem) <= 227 * 1024, "the teams of one CTA share one SM's shared memory");
template <int S>
__global__ void __launch_bounds__(TEAM_SLOTS * S, 1) bgzf_team_decode_k(const uint8_t *__restrict__ comp, const BgzfBlock *__restrict__ blocks, uint32_t nblocks,
                                                                        uint8_t *__restrict__ out, dflate2::Token *__restrict__ tok, uint32_t *__restrict__ ntok,
                                                                        int32_t *__restrict__ status) {
    extern __shared__ __align__(16) unsigned char smem[];
    const uint32_t team = threadIdx.x / S;
    dflate3::TeamMem *T = reinterpret_cast<dflate3::TeamMem *>(smem) + team;
    for (uint32_t base = blockIdx.x * TEAM_SLOTS; base < nblocks; base += gridDim.x * TEAM_SLOTS) {
        const uint32_t b = base + team;
        if (b >= nblocks) continue;                                     // whole teams skip together (collectives name the team's lanes only)
        const BgzfBlock B = blocks[b];
        uint32_t nt = 0;
        const int rc = dflate3::team_inflate(dflate::SubWarp<S>(), T, comp + B.coff, B.clen, out + B.uoff, B.usize, tok + B.tok, &nt);
        if (threadIdx.x % S == 0) { ntok[b] = nt; status[b] = rc; }
    }
}

// (4 CTAs of 8 warps per SM = 64 registers per thread: the ~4 200 blocks of a 1M-read BAM must be ONE wave -- at 78 registers they were two)
__global__ void __launch_bounds__(RES_WARPS * 32, 4) bgzf_resolve_k(const uint8_t *__restrict__ comp, const BgzfBlock *__restrict__ blocks, uint32_t nblocks, uint32_t block0,
                                                                  uint8_t *out, const dflate2::Token *__restrict__ tok, const uint32_t *__restrict__ ntok,
                                                                  const int32_t *__restrict__ status, unsigned long long *__restrict__ err) {
    const uint32_t *__restrict__ T = g_crc4.v;
    const uint32_t b = blockIdx.x * RES_WARPS + (threadIdx.x >> 5);
    if (b >= nblocks) return;                       // whole warps leave together (no block-wide barrier below)
    const BgzfBlock B = blocks[b];
    int rc = status[b];
    if (rc == dflate2::E_FALLBACK) return;          // bgzf_warp_inflate_k decodes this block
    if (rc == dflate2::OK) {
        rc = dflate2::resolve_bytes(dflate::WarpLanes(), tok + B.tok, ntok[b], out + B.uoff, B.usize, comp + B.coff);
        __syncwarp();
        if (rc == dflate2::OK && dflate2::crc32_block4(dflate::WarpLanes(), out + B.uoff, B.usize, T) != B.crc) rc = dflate2::E_CRC;
    }
    if (rc != dflate2::OK && (threadIdx.x & 31) == 0)
        atomicMin(err, ((unsigned long long)(block0 + b) << 8) | (unsigned long long)(uint8_t)(-rc));
}

// One warp per BGZF block, every lane running the same Huffman walk (dflate::Inflater2, inflate_core.cuh): the round-1 decoder.
// status != nullptr: only the blocks bgzf_team_decode_k handed back (E_FALLBACK: their Huffman tables exceed the arena);
// status == nullptr: every block (WGBS_INFLATE=2: the yardstick the two-phase decoder is measured against).
__global__ void __launch_bounds__(OLD_WARPS * 32, 8) bgzf_warp_inflate_k(const uint8_t *__restrict__ comp, const BgzfBlock *__restrict__ blocks, uint32_t nblocks, uint32_t block0,
                                                                          uint8_t *out, const int32_t *__restrict__ status, unsigned long long *__restrict__ err) {
    __shared__ dflate::Scratch S[OLD_WARPS];
    __shared__ dflate::Ring RG[OLD_WARPS];
    __shared__ uint32_t crc_table[256];
    const uint32_t w = threadIdx.x >> 5, b = blockIdx.x * OLD_WARPS + w;
    const bool mine = b < nblocks && (!status || status[b] == dflate2::E_FALLBACK);
    if (!__syncthreads_or(mine)) return;            // the usual case of the fallback launch: nothing to do
    for (uint32_t i = threadIdx.x; i < 256; i += OLD_WARPS * 32) crc_table[i] = dflate::crc_table_entry(i);
    __syncthreads();
    if (!mine) return;                              // whole warps leave together (no block-wide barrier below)
    const BgzfBlock B = blocks[b];
    dflate::Inflater2<dflate::WarpLanes> I;
    I.S = &S[w]; I.R = &RG[w]; I.dst = out + B.uoff; I.dst_len = B.usize;
    int rc = I.run(comp + B.coff, B.clen);
    if (rc == dflate::OK) {
        __syncwarp();
        if (dflate::crc32_block(dflate::WarpLanes(), out + B.uoff, B.usize, crc_table) != B.crc) rc = dflate::E_CRC;
    }
    if (rc != dflate::OK && (threadIdx.x & 31) == 0) atomicMin(err, ((unsigned long long)(block0 + b) << 8) | (unsigned long long)(uint8_t)(-rc));
}

}  // namespace

uint64_t bgzf_inflate2_plan(BgzfBlock *h_blocks, uint32_t nb) {
    // token slots are numbered inside a chunk of blocks: the scratch is reused chunk after chunk (launches on one stream run in order)
    uint64_t max_tok = 0;
    for (uint32_t c0 = 0; c0 < nb; c0 += CHUNK_BLOCKS) {
        uint64_t o = 0;
        for (uint32_t i = c0; i < nb && i < c0 + CHUNK_BLOCKS; i++) { h_blocks[i].tok = (uint32_t)o; o += dflate2::token_cap(h_blocks[i].usize); }
        if (o > max_tok) max_tok = o;
    }
    return max_tok;
}

int bgzf_inflate2_launch(wgbs_ctx *ctx, const uint8_t *d_comp, const BgzfBlock *d_blocks, uint32_t nb, uint64_t token_slots, uint8_t *out,
                         unsigned long long *d_err) {
    if (!nb) return 0;
    Temps T(ctx);
    dflate2::Token *tok; uint32_t *ntok; int32_t *status;
    // the token lists (~2.7 bytes per inflated byte in the worst case every block is sized for) live in the context's own scratch:
    // hundreds of MB taken from and returned to the stream-ordered pool on every call made the pool remap memory (6 ms per call)
    void *sc = nullptr;
    RC_TRY(ctx_scratch(ctx, (size_t)token_slots * sizeof(dflate2::Token), &sc));
    tok = (dflate2::Token *)sc;
    RC_TRY(T.alloc(&ntok, nb)); RC_TRY(T.alloc(&status, nb));
    static const size_t team_bytes = TEAM_SLOTS * sizeof(dflate3::TeamMem);
    CUDA_TRY(cudaFuncSetAttribute(bgzf_team_decode_k<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)team_bytes));
    for (uint32_t c0 = 0; c0 < nb; c0 += CHUNK_BLOCKS) {
        const uint32_t n = nb - c0 < CHUNK_BLOCKS ? nb - c0 : CHUNK_BLOCKS;
        const unsigned grid = (unsigned)std::min<uint32_t>((n + TEAM_SLOTS - 1) / TEAM_SLOTS, (uint32_t)ctx->sm_count);
        LAUNCH(ctx, bgzf_team_decode_k<32>, grid, TEAM_SLOTS * 32, team_bytes, d_comp, d_blocks + c0, n, out, tok, ntok + c0, status + c0);
        LAUNCH(ctx, bgzf_resolve_k, grid_for(n, RES_WARPS), RES_WARPS * 32, 0, d_comp, d_blocks + c0, n, c0, out, tok, ntok + c0, status + c0, d_err);
        LAUNCH(ctx, bgzf_warp_inflate_k, grid_for(n, OLD_WARPS), OLD_WARPS * 32, 0, d_comp, d_blocks + c0, n, c0, out, status + c0, d_err);
    }
    LAUNCH_CHECK();
    return 0;
}

int bgzf_inflate_warp_launch(wgbs_ctx *ctx, const uint8_t *d_comp, const BgzfBlock *d_blocks, uint32_t nb, uint8_t *out, unsigned long long *d_err) {
    if (!nb) return 0;
    LAUNCH(ctx, bgzf_warp_inflate_k, grid_for(nb, OLD_WARPS), OLD_WARPS * 32, 0, d_comp, d_blocks, nb, 0u, out, (const int32_t *)nullptr, d_err);
    LAUNCH_CHECK();
    return 0;
}
