// bamdev.cu -- BAM ingest ON THE DEVICE: the compressed .bam bytes are what crosses PCIe (~55 B per
// 150 bp read instead of ~360 B of SAM text); BGZF blocks are inflated by one warp each, BAM records are located in the
// inflated stream, filtered like `samtools view` and written out as the SAM text the pileup front end (wgbs_pileup_sam)
// consumes -- all in HBM.  Stands in for `samtools view BAM region -q Q -F X [-f Y] ...` of reference
// src/python/bam2pat.py:126-165.  The per-block / per-record logic lives in inflate_core.cuh / bam_core.cuh (host + device,
// pinned on the CPU by tests/test_bamdev_core.py); this file is the kernels around it and the C ABI.
//
//   bgzf_inflate_k   warp per BGZF block: lane 0 decodes Huffman symbols, the warp writes literals / copies matches
//   bam_guess_k      warp per 16 KiB segment of the inflated stream: first plausible record start (a guess)
//   bam_walk_k       thread per segment: follow the block_size chain from the segment's entry -> record count, exit
//   bam_check_k      entry[s+1] must equal exit[s]; repaired and re-walked until nothing changes (exact, whatever the guesses)
//   bam_fill_k       record offsets; bam_runs_k: per-reference record ranges + sortedness + record sanity
//   bam_measure_k    thread per record: filters + SAM line length;  bam_format_k: warp per record: the line
#include <chrono>
#include <string>
#include <vector>

#include "bam_core.cuh"
#include "bgzf.cuh"
#include "common.cuh"
#include "reads.cuh"

using namespace bamcore;

struct wgbs_dbam {
    uint8_t *data = nullptr;       // inflated BAM stream (device), padded by 16 bytes
    uint64_t n = 0;
    uint64_t *rec_off = nullptr;   // offset of every record's block_size field (device)
    uint64_t nrec = 0;
    std::string header_text;
    std::vector<std::string> ref_names;
    std::vector<int32_t> ref_lens;
    std::vector<uint64_t> ref_first, ref_last;   // record index range per refID (+ one slot for unplaced records)
    uint32_t *d_name_off = nullptr; char *d_names = nullptr; int32_t *d_ref_lens = nullptr;
    uint64_t comp_bytes = 0, n_blocks = 0;
};

namespace {

constexpr uint64_t SEG = 16384;     // bytes of inflated stream per boundary-search segment
constexpr int GUESS_DEPTH = 4;      // consecutive plausible records required by a guess
constexpr int FMT_G = 8;            // lanes per record in bam_format_k
__global__ void __launch_bounds__(128) bam_guess_k(const uint8_t *__restrict__ data, uint64_t n, uint64_t p0, uint64_t nseg, int32_t n_ref,
                                                    uint64_t *__restrict__ entry) {
    const uint64_t s = (uint64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (s >= nseg) return;
    dflate::WarpLanes lanes;
    const uint64_t e = s == 0 ? p0 : guess_entry(lanes, data, n, p0 + s * SEG, n_ref, GUESS_DEPTH);
    if ((threadIdx.x & 31) == 0) entry[s] = e;
}

__global__ void __launch_bounds__(128) bam_walk_k(const uint8_t *__restrict__ data, uint64_t n, uint64_t p0, uint64_t nseg,
                                                   const uint64_t *__restrict__ entry, uint8_t *__restrict__ dirty, uint64_t *__restrict__ exit_,
                                                   uint32_t *__restrict__ cnt, uint64_t *__restrict__ bad) {
    const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseg || !dirty[s]) return;
    uint32_t c = 0; uint64_t b = ~0ull;
    exit_[s] = walk_chain(data, n, entry[s], p0 + (s + 1) * SEG, &c, nullptr, &b);
    cnt[s] = c; bad[s] = b; dirty[s] = 0;
}

// thread s decides the entry of segment s+1 from the OLD entries (bam_core.cuh repaired_entry) and writes it to the new array
__global__ void __launch_bounds__(256) bam_check_k(uint64_t nseg, const uint64_t *__restrict__ entry, uint64_t *__restrict__ entry_new,
                                                    const uint64_t *__restrict__ exit_, const uint64_t *__restrict__ bad, uint8_t *__restrict__ dirty,
                                                    uint32_t *__restrict__ changed) {
    const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s == 0) entry_new[0] = entry[0];
    if (s + 1 >= nseg) return;
    bool ch = false;
    entry_new[s + 1] = repaired_entry(entry, exit_, bad, s, &ch);
    if (ch) { dirty[s + 1] = 1; atomicAdd(changed, 1u); }
}

// first corrupt record in stream order (segments are in stream order; a bad segment stops the chain)
__global__ void __launch_bounds__(256) bam_first_bad_k(uint64_t nseg, const uint64_t *__restrict__ bad, unsigned long long *__restrict__ out) {
    const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s < nseg && bad[s] != ~0ull) atomicMin(out, (unsigned long long)bad[s]);
}

__global__ void __launch_bounds__(128) bam_fill_k(const uint8_t *__restrict__ data, uint64_t n, uint64_t p0, uint64_t nseg,
                                                   const uint64_t *__restrict__ entry, const uint64_t *__restrict__ base, uint64_t *__restrict__ rec_off) {
    const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseg) return;
    uint32_t c; uint64_t b;
    walk_chain(data, n, entry[s], p0 + (s + 1) * SEG, &c, rec_off + base[s], &b);
}

// per-reference runs of a coordinate-sorted file: first / last record index and the number of separate runs per slot
// (slot n_ref: unplaced or out-of-range refID).  Also the record sanity check the per-record kernels rely on.
__global__ void __launch_bounds__(256) bam_runs_k(const uint8_t *__restrict__ data, const uint64_t *__restrict__ rec_off, uint64_t nrec, int32_t n_ref,
                                                   unsigned long long *__restrict__ first, unsigned long long *__restrict__ last, uint32_t *__restrict__ runs,
                                                   unsigned long long *__restrict__ bad_rec) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrec) return;
    Rec R; R.load(data + rec_off[i]);
    if (!R.consistent()) atomicMin(bad_rec, (unsigned long long)rec_off[i]);
    const int32_t me = R.refid;
    const uint32_t slot = (me >= 0 && me < n_ref) ? (uint32_t)me : (uint32_t)n_ref;
    const int32_t prev = i ? ldi32(data + rec_off[i - 1] + 4) : 0, next = i + 1 < nrec ? ldi32(data + rec_off[i + 1] + 4) : 0;
    if (i == 0 || prev != me) { atomicAdd(&runs[slot], 1u); first[slot] = i; }
    if (i + 1 == nrec || next != me) last[slot] = i + 1;
}

struct DevRefs { int32_t n; const uint32_t *name_off; const char *names; const int32_t *lens; };

__global__ void __launch_bounds__(256) bam_measure_k(const uint8_t *__restrict__ data, const uint64_t *__restrict__ rec_off, uint64_t nr, DevRefs F, ViewParams V,
                                                      uint32_t *__restrict__ len, uint32_t *__restrict__ pass, unsigned long long *__restrict__ too_long) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nr) return;
    Rec R; R.load(data + rec_off[i]);
    uint32_t l = 0, ok = 0;
    if (passes(R, V)) {
        const Refs RF{F.n, F.name_off, F.names, F.lens};
        CountSink cs; format_record(R, RF, cs);
        if (cs.n > 0xffffffffull) atomicMin(too_long, (unsigned long long)i); else { l = (uint32_t)cs.n; ok = 1; }
    }
    len[i] = l;
    if (pass) pass[i] = ok;
}

// head -N: only the first max_records passing records keep their line
__global__ void __launch_bounds__(256) bam_head_k(uint64_t nr, const uint32_t *__restrict__ rank, uint32_t max_records, uint32_t *__restrict__ len) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nr && rank[i] >= max_records) len[i] = 0;
}

__global__ void __launch_bounds__(256) bam_format_k(const uint8_t *__restrict__ data, const uint64_t *__restrict__ rec_off, uint64_t nr, DevRefs F,
                                                     const uint32_t *__restrict__ len, const uint64_t *__restrict__ off, char *__restrict__ text) {
    // FMT_G lanes per record: the scalar fields are written by the group's first lane, SEQ and QUAL by all of them; four
    // records per warp share one instruction stream (a whole warp per record spent 32 lanes on scalar work: 1.96 ms / 1M records)
    const uint64_t i = ((uint64_t)blockIdx.x * 256 + threadIdx.x) / FMT_G;
    if (i >= nr || len[i] == 0) return;
    Rec R; R.load(data + rec_off[i]);
    const Refs RF{F.n, F.name_off, F.names, F.lens};
    WriteSink<dflate::LaneGroup<FMT_G>> ws; ws.o = text + off[i];
    format_record(R, RF, ws);
}

// ---- BAM records -> pileup descriptors, no SAM text (BAM flavour of ReadBatch, reads.cuh) ---------------------------------------
__global__ void __launch_bounds__(256) bam_pass_k(const uint8_t *__restrict__ data, const uint64_t *__restrict__ rec_off, uint64_t nr, ViewParams V,
                                                   uint32_t *__restrict__ pass) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nr) return;
    Rec R; R.load(data + rec_off[i]);
    pass[i] = passes(R, V) ? 1u : 0u;
}

// first / last passing record of a view (only asked for when the whole record range is too wide for 32-bit offsets)
__global__ void __launch_bounds__(256) pass_range_k(const uint32_t *__restrict__ pass, uint64_t nr, uint32_t *__restrict__ first_last) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nr && pass[i]) { atomicMin(&first_last[0], (uint32_t)i); atomicMax(&first_last[1], (uint32_t)i); }
}

// has_mm[0] = 1 iff the FIRST passing record carries an MM:Z: / Mm:Z: tag (patter.cpp:337-338 looks at the first line only)
__global__ void __launch_bounds__(256) bam_records_k(const uint8_t *__restrict__ data, const uint64_t *__restrict__ rec_off, uint64_t nr,
                                                      const uint32_t *__restrict__ pass, const uint32_t *__restrict__ rank, uint32_t max_records,
                                                      uint64_t base, ReadBatch rb, uint64_t *__restrict__ rec_abs, uint32_t *__restrict__ has_mm) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nr || !pass[i]) return;
    const uint32_t k = rank[i];
    if (max_records && k >= max_records) return;
    const uint64_t o = rec_off[i];
    Rec R; R.load(data + o);
    const uint8_t *nm = R.name();
    uint32_t ql = 0; while (ql < R.l_name && nm[ql]) ql++;                 // l_read_name counts the NUL
    uint64_t h = 0;
    for (uint32_t p = 0; p < ql; p++) h += name_byte_mix(nm[p], p);
    h = fmix64(h ^ (uint64_t)ql);
    rb.line_off[k] = (uint32_t)(o + 36 - base); rb.line_len[k] = 0; rb.qn_len[k] = ql;
    rb.flag[k] = (int32_t)R.flag; rb.pos[k] = (int32_t)((int64_t)R.pos + 1); rb.pos_hi[k] = 0;
    rb.cig_off[k] = (uint32_t)(R.cigar() - data - base); rb.cig_len[k] = R.n_cig;
    if (R.l_seq > 0) { rb.seq_off[k] = (uint32_t)(R.seq() - data - base); rb.seq_len[k] = (uint32_t)R.l_seq; }
    else { rb.seq_off[k] = SEQ_STAR; rb.seq_len[k] = 1; }
    rb.hash_lo[k] = (uint32_t)h; rb.hash_hi[k] = (uint32_t)(h >> 32);
    rb.status[k] = REC_OK;
    rec_abs[k] = o;
    if (k == 0) {
        uint32_t zl = 0;
        has_mm[0] = (find_z_tag(R.tags(), R.end(), 'M', 'M', &zl) || find_z_tag(R.tags(), R.end(), 'M', 'm', &zl)) ? 1u : 0u;
    }
}

// MM/ML mode of the direct route: where the MM:Z string and the ML:B:C values of every passing record stand in the inflated stream
// (what the SAM tokenizer records as spans of its text, sam.cu tag_kind); ml_len = number of values + 1, 0 = no ML tag (np.cu)
__global__ void __launch_bounds__(256) bam_np_tags_k(const uint8_t *__restrict__ data, const uint64_t *__restrict__ rec_abs, uint32_t n, uint64_t base, ReadBatch rb) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    Rec R; R.load(data + rec_abs[k]);
    const uint8_t *mm, *ml; uint32_t mm_len, ml_cnt;
    find_np_tags(R.tags(), R.end(), &mm, &mm_len, &ml, &ml_cnt);
    rb.mm_off[k] = mm ? (uint32_t)(mm - data - base) : 0u; rb.mm_len[k] = mm ? mm_len : 0u;
    rb.ml_off[k] = ml ? (uint32_t)(ml - data - base) : 0u; rb.ml_len[k] = ml ? ml_cnt + 1 : 0u;
}

__global__ void __launch_bounds__(128) side_len_k(BamSide S, const uint32_t *__restrict__ ids, uint32_t m, uint32_t *__restrict__ len) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    Rec R; R.load(S.data + S.rec[ids[j]]);
    const Refs RF{S.n_ref, S.name_off, S.names, S.ref_lens};
    CountSink cs; format_record(R, RF, cs);
    len[j] = cs.n > 0xffffffffull ? 0xffffffffu : (uint32_t)cs.n;
}
__global__ void __launch_bounds__(128) side_write_k(BamSide S, const uint32_t *__restrict__ ids, uint32_t m, const uint32_t *__restrict__ len,
                                                     const uint64_t *__restrict__ off, char *__restrict__ text, uint32_t *__restrict__ line_off,
                                                     uint32_t *__restrict__ line_len) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    Rec R; R.load(S.data + S.rec[ids[j]]);
    const Refs RF{S.n_ref, S.name_off, S.names, S.ref_lens};
    WriteSink<dflate::OneLane> ws; ws.o = text + off[j];
    format_record(R, RF, ws);
    line_off[ids[j]] = (uint32_t)off[j]; line_len[ids[j]] = len[j] - 1;           // without the newline, like the tokenizer's lines
}

int fail_free(wgbs_ctx *ctx, wgbs_dbam *B, int rc) {
    if (B) { dfree(ctx, B->data); dfree(ctx, B->rec_off); dfree(ctx, B->d_name_off); dfree(ctx, B->d_names); dfree(ctx, B->d_ref_lens); delete B; }
    return rc;
}

const char *inflate_msg(int code) {
    switch (code) {
        case dflate::E_INPUT: return "deflate stream ends early";
        case dflate::E_BTYPE: return "invalid deflate block type";
        case dflate::E_STORED: return "invalid stored block length";
        case dflate::E_CODES: return "invalid code lengths";
        case dflate::E_SYMBOL: return "invalid code";
        case dflate::E_DIST: return "distance too far back";
        case dflate::E_OUTPUT: return "more data than ISIZE";
        case dflate::E_SHORT: return "less data than ISIZE";
        case dflate::E_CRC: return "CRC32 mismatch";
        default: return "inflate error";
    }
}

// hop over the block headers of a BGZF file in host memory: the block table (what `bgzip -i` records in a .gzi, plus CRC32 / ISIZE)
int bgzf_scan_blocks(const void *bgzf, size_t nbytes, const char *who, std::vector<BgzfBlock> &blocks, uint64_t *inflated) {
    const uint8_t *f = (const uint8_t *)bgzf;
    uint64_t off = 0, uoff = 0;
    while (off + 28 <= nbytes) {                       // one hop per block over the headers
        uint32_t xlen = 0; const uint32_t bs = dflate::bgzf_block_size(f + off, nbytes - off, &xlen);
        if (!bs) return wgbs_set_err("%s: not a BGZF file (bad block header at %llu)", who, (unsigned long long)off);
        if (off + bs > nbytes || bs < 12 + xlen + 8) return wgbs_set_err("%s: corrupt BGZF block at %llu", who, (unsigned long long)off);
        BgzfBlock b; b.coff = off + 12 + xlen; b.clen = bs - 12 - xlen - 8; b.usize = ld32(f + off + bs - 4); b.crc = ld32(f + off + bs - 8); b.uoff = uoff; b.tok = 0;
        if (b.usize > 65536) return wgbs_set_err("%s: corrupt BGZF block at %llu (ISIZE %u)", who, (unsigned long long)off, b.usize);
        blocks.push_back(b); off += bs; uoff += b.usize;
    }
    if (off != nbytes) return wgbs_set_err("%s: %llu trailing bytes after the last BGZF block", who, (unsigned long long)(nbytes - off));
    if (blocks.size() >= 0xffffffffull) return wgbs_set_err("%s: too many BGZF blocks", who);
    *inflated = uoff;
    return 0;
}

// Inflate a whole BGZF file into a fresh device buffer (padded by 16 zero bytes).  bgzf: host bytes; or, with a prebuilt block
// table (pre != nullptr: wgbs_bgzf_index), host or DEVICE bytes.  Synchronises the stream: errors (framing, deflate, ISIZE, CRC32)
// are reported here.
int bgzf_inflate_device(wgbs_ctx *ctx, const void *bgzf, size_t nbytes, const char *who, uint8_t **d_out, uint64_t *n_out, uint64_t *n_blocks,
                        const std::vector<BgzfBlock> *pre = nullptr) {
    const uint8_t *f = (const uint8_t *)bgzf;
    std::vector<BgzfBlock> scanned; uint64_t uoff = 0;
    if (pre) { if (!pre->empty()) uoff = pre->back().uoff + pre->back().usize; }
    else RC_TRY(bgzf_scan_blocks(bgzf, nbytes, who, scanned, &uoff));
    std::vector<BgzfBlock> blocks = pre ? *pre : std::move(scanned);
    const bool src_on_device = is_device_ptr(bgzf);
    Temps T(ctx);
    uint8_t *d_comp, *data = nullptr; BgzfBlock *d_blocks; unsigned long long *d_err;
    int rc;
    const uint64_t token_slots = bgzf_inflate2_plan(blocks.data(), (uint32_t)blocks.size());
    d_comp = nullptr;
    if (src_on_device) d_comp = const_cast<uint8_t *>(f);             // resident already (readable for 64 bytes past its end: wgbs_dbam_open_indexed)
    if ((!src_on_device && (rc = T.alloc(&d_comp, nbytes + 64)) < 0) || (rc = T.alloc(&d_blocks, blocks.size())) < 0 || (rc = T.alloc(&d_err, 1)) < 0 || (rc = dalloc(ctx, &data, uoff + 16)) < 0) {
        // (cudaMemGetInfo costs a fraction of a millisecond: asked only to word the error)
        cudaGetLastError();
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        return wgbs_set_err("%s: the inflated stream (%.1f GB + %.1f GB compressed) does not fit in device memory (%.1f GB free); process the file in parts",
                            who, uoff / 1e9, nbytes / 1e9, free_b / 1e9);
    }
    if ((!src_on_device && (rc = copy_any(ctx, d_comp, f, nbytes)) < 0) || (rc = copy_any(ctx, d_blocks, blocks.data(), blocks.size() * sizeof(BgzfBlock))) < 0) { dfree(ctx, data); return rc; }
    unsigned long long herr = 0;
    cudaError_t e = cudaMemsetAsync(d_err, 0xff, 8, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(data + uoff, 0, 16, ctx->stream);
    // read per call (a getenv is nothing next to an inflate): tests and bench legs switch decoders inside one process
    const char *ev = getenv("WGBS_INFLATE");
    const bool old_decoder = ev && ev[0] == '2';
    if (e == cudaSuccess && !blocks.empty()) {
        const uint32_t nb = (uint32_t)blocks.size();
        if (old_decoder) { if ((rc = bgzf_inflate_warp_launch(ctx, d_comp, d_blocks, nb, data, d_err)) < 0) { dfree(ctx, data); return rc; } }
        else if ((rc = bgzf_inflate2_launch(ctx, d_comp, d_blocks, nb, token_slots, data, d_err)) < 0) { dfree(ctx, data); return rc; }
        if (e == cudaSuccess) e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(&herr, d_err, 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { dfree(ctx, data); return wgbs_set_err("%s: %s", who, cudaGetErrorString(e)); }
    if (herr != ~0ull) { dfree(ctx, data); return wgbs_set_err("%s: inflate failed in BGZF block %llu (%s)", who, herr >> 8, inflate_msg(-(int)(herr & 0xff))); }
    *d_out = data; *n_out = uoff; if (n_blocks) *n_blocks = blocks.size();
    return 0;
}

}  // namespace

// bgzf: the bytes of a whole .bam file in HOST memory (pinned memory makes the upload one DMA).
// part: a window of a file (wgbs_dbam_open_part).  The stream then has no BAM header unless part->has_header (the first window of a
// file), records are indexed from part->first_record on, and the last record may be cut off by the end of the window
// (*part->tail = its offset; everything behind it is left unindexed).
struct PartArgs { int has_header; int n_ref; const char *const *ref_names; const int32_t *ref_lens; uint64_t first_record; uint64_t *tail; };
struct wgbs_bgzf_index { std::vector<BgzfBlock> blocks; uint64_t file_bytes = 0, inflated = 0; };
static int dbam_open_impl(wgbs_ctx *ctx, const void *bgzf, size_t nbytes, const PartArgs *part, wgbs_dbam **out, const wgbs_bgzf_index *bix = nullptr) {
    RC_TRY(wgbs_ctx_activate(ctx));
    if (!bgzf || !out) return wgbs_set_err("wgbs_dbam_open: null argument");
    *out = nullptr;
    if (!bix && is_device_ptr(bgzf)) return wgbs_set_err("wgbs_dbam_open: the compressed bytes must be in host memory (the block table is read on the host); device-resident bytes need wgbs_dbam_open_indexed");
    if (bix && bix->file_bytes != nbytes) return wgbs_set_err("wgbs_dbam_open_indexed: the index was built for a file of %llu bytes, not %zu", (unsigned long long)bix->file_bytes, nbytes);
    wgbs_dbam *B = new wgbs_dbam();
    Temps T(ctx);
    int rc;
    unsigned long long *d_err;
    static const bool dbg = getenv("WGBS_BAM_DEBUG") != nullptr;          // host-side phase times on stderr
    auto T0 = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (!dbg) return;
        cudaStreamSynchronize(ctx->stream);
        auto t = std::chrono::steady_clock::now();
        fprintf(stderr, "[dbam] %-12s %.3f ms\n", what, std::chrono::duration<double, std::milli>(t - T0).count()); T0 = t;
    };
    if ((rc = T.alloc(&d_err, 4)) < 0) return fail_free(ctx, B, rc);
    // 1 + 2. BGZF block table (host) and inflate (one warp per block)
    { uint64_t nb = 0; if ((rc = bgzf_inflate_device(ctx, bgzf, nbytes, "wgbs_dbam_open", &B->data, &B->n, &nb, bix ? &bix->blocks : nullptr)) < 0) return fail_free(ctx, B, rc); B->n_blocks = nb; }
    B->comp_bytes = nbytes;
    lap("scan+inflate");
    const uint64_t uoff = B->n;
    unsigned long long herr[4];
    // 3. header: "BAM\1" l_text text n_ref { l_name name l_ref }   (SAM spec 4.2)
    std::vector<uint8_t> head(std::min<uint64_t>(uoff, 1u << 16));
    if (!head.empty()) CUDA_TRY(cudaMemcpyAsync(head.data(), B->data, head.size(), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    auto need = [&](uint64_t upto) -> int {       // make head[0..upto) available
        if (upto > uoff) return wgbs_set_err("wgbs_dbam_open: truncated BAM header");
        if (upto <= head.size()) return 0;
        const size_t old = head.size(); head.resize(std::min<uint64_t>(uoff, std::max<uint64_t>(upto, 2 * old)));
        CUDA_TRY(cudaMemcpyAsync(head.data() + old, B->data + old, head.size() - old, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        return 0;
    };
    uint64_t p = 0; uint32_t n_ref = 0;
    std::vector<uint32_t> name_off{0}; std::string names;
    if (!part || part->has_header) {
        if ((rc = need(12)) < 0 || memcmp(head.data(), "BAM\1", 4)) return fail_free(ctx, B, rc < 0 ? rc : wgbs_set_err("wgbs_dbam_open: not a BAM file"));
        const uint32_t l_text = ld32(head.data() + 4);
        if ((rc = need(8ull + l_text + 4)) < 0) return fail_free(ctx, B, rc);
        B->header_text.assign((const char *)head.data() + 8, strnlen((const char *)head.data() + 8, l_text));
        p = 8ull + l_text; n_ref = ld32(head.data() + p); p += 4;
        if (n_ref > 0x7ffffffeu) return fail_free(ctx, B, wgbs_set_err("wgbs_dbam_open: corrupt BAM header"));
        for (uint32_t i = 0; i < n_ref; i++) {
            if ((rc = need(p + 4)) < 0) return fail_free(ctx, B, rc);
            const uint32_t l = ld32(head.data() + p); p += 4;
            if ((rc = need(p + l + 4)) < 0) return fail_free(ctx, B, rc);
            B->ref_names.emplace_back((const char *)head.data() + p, l ? strnlen((const char *)head.data() + p, l - 1) : 0); p += l;
            B->ref_lens.push_back(ldi32(head.data() + p)); p += 4;
            names += B->ref_names.back(); name_off.push_back((uint32_t)names.size());
        }
    } else {
        n_ref = (uint32_t)part->n_ref;
        for (uint32_t i = 0; i < n_ref; i++) {
            B->ref_names.emplace_back(part->ref_names[i]); B->ref_lens.push_back(part->ref_lens ? part->ref_lens[i] : 0);
            names += B->ref_names.back(); name_off.push_back((uint32_t)names.size());
        }
        p = part->first_record;
        if (p > uoff) return fail_free(ctx, B, wgbs_set_err("wgbs_dbam_open_part: first record offset %llu beyond the part (%llu inflated bytes)", (unsigned long long)p, (unsigned long long)uoff));
    }
    if ((rc = dalloc(ctx, &B->d_name_off, name_off.size())) < 0 || (rc = dalloc(ctx, &B->d_names, names.size())) < 0 || (rc = dalloc(ctx, &B->d_ref_lens, (size_t)n_ref)) < 0 ||
        (rc = copy_any(ctx, B->d_name_off, name_off.data(), name_off.size() * 4)) < 0 || (rc = copy_any(ctx, B->d_names, names.data(), names.size())) < 0 ||
        (rc = copy_any(ctx, B->d_ref_lens, B->ref_lens.data(), (size_t)n_ref * 4)) < 0) return fail_free(ctx, B, rc);
    lap("header");
    // 4. record table: guess the first record of every segment, walk, repair until every entry equals its predecessor's exit
    const uint64_t p0 = p, n = uoff; uint64_t nseg = n > p0 ? (n - p0 + SEG - 1) / SEG : 0;
    if (part && part->tail) *part->tail = n > p0 ? n : p0;
    B->ref_first.assign(n_ref + 1, 0); B->ref_last.assign(n_ref + 1, 0);
    if (nseg) {
        uint64_t *entry, *entry2, *exit_, *bad, *base; uint32_t *cnt, *changed; uint8_t *dirty; unsigned long long *d_first, *d_last; uint32_t *d_runs;
        if ((rc = T.alloc(&entry, nseg)) < 0 || (rc = T.alloc(&entry2, nseg)) < 0 || (rc = T.alloc(&exit_, nseg)) < 0 || (rc = T.alloc(&bad, nseg)) < 0 || (rc = T.alloc(&base, nseg + 1)) < 0 ||
            (rc = T.alloc(&cnt, nseg)) < 0 || (rc = T.alloc(&changed, 1)) < 0 || (rc = T.alloc(&dirty, nseg)) < 0) return fail_free(ctx, B, rc);
        CUDA_TRY(cudaMemsetAsync(dirty, 1, nseg, ctx->stream));
        LAUNCH(ctx, bam_guess_k, grid_for(nseg, 4), 128, 0, B->data, n, p0, nseg, (int32_t)n_ref, entry);
        for (uint64_t round = 0;; round++) {
            CUDA_TRY(cudaMemsetAsync(changed, 0, 4, ctx->stream));
            LAUNCH(ctx, bam_walk_k, grid_for(nseg, 128), 128, 0, B->data, n, p0, nseg, entry, dirty, exit_, cnt, bad);
            LAUNCH(ctx, bam_check_k, grid_for(nseg, 256), 256, 0, nseg, entry, entry2, exit_, bad, dirty, changed);
            std::swap(entry, entry2);
            uint32_t h = 0;
            CUDA_TRY(cudaMemcpyAsync(&h, changed, 4, cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_TRY(cudaStreamSynchronize(ctx->stream));
            if (!h) break;
            if (round > nseg) return fail_free(ctx, B, wgbs_set_err("wgbs_dbam_open: record table did not converge"));
        }
        lap("guess+walk");
        CUDA_TRY(cudaMemsetAsync(d_err, 0xff, 4 * 8, ctx->stream));
        LAUNCH(ctx, bam_first_bad_k, grid_for(nseg, 256), 256, 0, nseg, bad, d_err);
        if (part) {
            // a window of a file may end inside a record: the first record that does not fit is where the chain of this part ends
            // (its offset = the tail the caller restarts from); segments behind it hold no record of this part.  Anything else that
            // cannot be a record is still an error.
            CUDA_TRY(cudaMemcpyAsync(herr, d_err, 8, cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_TRY(cudaStreamSynchronize(ctx->stream));
            if (herr[0] != ~0ull) {
                const uint64_t o = herr[0]; uint32_t bs = 0;
                if (o + 4 <= n) { uint8_t b4[4]; CUDA_TRY(cudaMemcpyAsync(b4, B->data + o, 4, cudaMemcpyDeviceToHost, ctx->stream)); CUDA_TRY(cudaStreamSynchronize(ctx->stream)); bs = ld32(b4); }
                const bool cut = o + 4 > n || (bs >= 32 && o + 4 + (uint64_t)bs > n);
                if (!cut) return fail_free(ctx, B, wgbs_set_err("wgbs_dbam_open_part: corrupt BAM record at uncompressed offset %llu", herr[0]));
                if (part->tail) *part->tail = o;
                nseg = o > p0 ? (o - p0) / SEG + 1 : 1;                       // the segment the cut-off record starts in is the last one
                CUDA_TRY(cudaMemsetAsync(d_err, 0xff, 4 * 8, ctx->stream));
            }
        }
        if ((rc = scan_u32_u64(ctx, cnt, base, nseg)) < 0) return fail_free(ctx, B, rc);
        uint64_t nrec = 0;
        CUDA_TRY(cudaMemcpyAsync(&nrec, base + nseg, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(herr, d_err, sizeof herr, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        if (herr[0] != ~0ull) return fail_free(ctx, B, wgbs_set_err("wgbs_dbam_open: corrupt BAM record at uncompressed offset %llu", herr[0]));
        B->nrec = nrec;
        if ((rc = dalloc(ctx, &B->rec_off, (size_t)nrec)) < 0) return fail_free(ctx, B, rc);
        LAUNCH(ctx, bam_fill_k, grid_for(nseg, 128), 128, 0, B->data, n, p0, nseg, entry, base, B->rec_off);
        if (nrec) {
            if ((rc = T.alloc(&d_first, (size_t)n_ref + 1)) < 0 || (rc = T.alloc(&d_last, (size_t)n_ref + 1)) < 0 || (rc = T.alloc(&d_runs, (size_t)n_ref + 1)) < 0) return fail_free(ctx, B, rc);
            CUDA_TRY(cudaMemsetAsync(d_first, 0, ((size_t)n_ref + 1) * 8, ctx->stream)); CUDA_TRY(cudaMemsetAsync(d_last, 0, ((size_t)n_ref + 1) * 8, ctx->stream));
            CUDA_TRY(cudaMemsetAsync(d_runs, 0, ((size_t)n_ref + 1) * 4, ctx->stream));
            LAUNCH(ctx, bam_runs_k, grid_for(nrec, 256), 256, 0, B->data, B->rec_off, nrec, (int32_t)n_ref, d_first, d_last, d_runs, d_err + 1);
            std::vector<uint32_t> runs(n_ref + 1);
            static_assert(sizeof(unsigned long long) == sizeof(uint64_t), "");
            CUDA_TRY(cudaMemcpyAsync(B->ref_first.data(), d_first, ((size_t)n_ref + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_TRY(cudaMemcpyAsync(B->ref_last.data(), d_last, ((size_t)n_ref + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_TRY(cudaMemcpyAsync(runs.data(), d_runs, ((size_t)n_ref + 1) * 4, cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_TRY(cudaMemcpyAsync(herr, d_err, sizeof herr, cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_TRY(cudaStreamSynchronize(ctx->stream));
            if (herr[1] != ~0ull) return fail_free(ctx, B, wgbs_set_err("wgbs_dbam_open: corrupt BAM record at uncompressed offset %llu", herr[1]));
            for (uint32_t i = 0; i <= n_ref; i++)
                if (runs[i] > 1) return fail_free(ctx, B, wgbs_set_err("the BAM is not sorted by coordinate (reference %d appears in two separate runs)", i < n_ref ? (int)i : -1));
        }
    }
    lap("table+runs");
    LAUNCH_CHECK();
    *out = B;
    return 0;
}

extern "C" int wgbs_dbam_open(wgbs_ctx *ctx, const void *bgzf, size_t nbytes, wgbs_dbam **out) { return dbam_open_impl(ctx, bgzf, nbytes, nullptr, out); }

// The block table of a BGZF file, built once per file from its bytes in host memory (the offsets a .gzi holds, plus each block's
// CRC32 and ISIZE): with it the compressed bytes may already be resident on the device when the file is opened.
extern "C" int wgbs_bgzf_index_build(const void *bgzf, size_t nbytes, wgbs_bgzf_index **out) {
    if (!bgzf || !out) return wgbs_set_err("wgbs_bgzf_index_build: null argument");
    if (is_device_ptr(bgzf)) return wgbs_set_err("wgbs_bgzf_index_build: the bytes must be in host memory");
    wgbs_bgzf_index *ix = new wgbs_bgzf_index();
    const int rc = bgzf_scan_blocks(bgzf, nbytes, "wgbs_bgzf_index_build", ix->blocks, &ix->inflated);
    if (rc < 0) { delete ix; return rc; }
    ix->file_bytes = nbytes;
    *out = ix;
    return 0;
}
extern "C" void wgbs_bgzf_index_free(wgbs_bgzf_index *ix) { delete ix; }
extern "C" uint64_t wgbs_bgzf_index_blocks(const wgbs_bgzf_index *ix) { return ix ? ix->blocks.size() : 0; }
// wgbs_dbam_open with the block table at hand: bgzf may be HOST or DEVICE memory (a device buffer must be readable for 64 bytes past
// nbytes: the decoder fetches whole 16-byte chunks)
extern "C" int wgbs_dbam_open_indexed(wgbs_ctx *ctx, const void *bgzf, size_t nbytes, const wgbs_bgzf_index *ix, wgbs_dbam **out) {
    if (!ix) return wgbs_set_err("wgbs_dbam_open_indexed: null index");
    return dbam_open_impl(ctx, bgzf, nbytes, nullptr, out, ix);
}

// A window of a .bam on the device: see wgbs_bam_open_part (include/wgbs_b200.h)
extern "C" int wgbs_dbam_open_part(wgbs_ctx *ctx, const void *bgzf, size_t nbytes, int n_ref, const char *const *ref_names, const int32_t *ref_lens,
                                   int has_header, uint64_t first_record, wgbs_dbam **out, uint64_t *tail) {
    if (!tail || (!has_header && n_ref > 0 && !ref_names)) return wgbs_set_err("wgbs_dbam_open_part: null argument");
    PartArgs pa{has_header, n_ref, ref_names, ref_lens, first_record, tail};
    return dbam_open_impl(ctx, bgzf, nbytes, &pa, out);
}

__global__ void __launch_bounds__(256) bam_first_key_k(const uint8_t *__restrict__ data, const uint64_t *__restrict__ rec_off, uint64_t nr, ViewParams V,
                                                        int64_t key, unsigned long long *__restrict__ first) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nr) return;
    Rec R; R.load(data + rec_off[i]);
    if (template_key((int32_t)R.flag, R.refid, R.pos, R.nref, R.npos) >= key && passes(R, V)) atomicMin(first, (unsigned long long)rec_off[i]);
}

// inflated offset of the first record of reference refid that passes the filters (key window ignored) and whose template key is
// >= key; *found = 0 when there is none (wgbs_bam_first_key on the device)
extern "C" int wgbs_dbam_first_key(wgbs_ctx *ctx, const wgbs_dbam *B, const wgbs_view_opts *vo, int refid, int64_t key, uint64_t *offset, int *found) {
    RC_TRY(wgbs_ctx_activate(ctx));
    if (!B || !vo || !offset || !found) return wgbs_set_err("wgbs_dbam_first_key: null argument");
    *found = 0; *offset = 0;
    if (refid < 0 || refid >= (int)B->ref_names.size()) return 0;
    if (vo->n_flag_eq < 0 || vo->n_flag_eq > 4) return wgbs_set_err("wgbs_dbam_first_key: n_flag_eq must be 0..4");
    if (vo->n_iv && (!vo->iv_beg || !vo->iv_end || is_device_ptr(vo->iv_beg) || is_device_ptr(vo->iv_end))) return wgbs_set_err("wgbs_dbam_first_key: interval lists must be host arrays");
    const uint64_t r0 = B->ref_first[refid], r1 = B->ref_last[refid], nr = r1 - r0;
    if (!nr) return 0;
    Temps T(ctx);
    ViewParams V; memset(&V, 0, sizeof V);
    V.refid = refid; V.min_mapq = vo->min_mapq; V.exclude_flags = vo->exclude_flags; V.include_flags = vo->include_flags; V.beg = vo->beg; V.end = vo->end;
    V.n_flag_eq = vo->n_flag_eq; for (int k = 0; k < 4; k++) V.flag_eq[k] = vo->flag_eq[k];
    V.n_iv = vo->n_iv; V.iv_exclude = vo->iv_exclude;
    if (vo->n_iv) {
        int64_t *a, *b;
        RC_TRY(T.alloc(&a, vo->n_iv)); RC_TRY(T.alloc(&b, vo->n_iv));
        RC_TRY(copy_any(ctx, a, vo->iv_beg, vo->n_iv * 8)); RC_TRY(copy_any(ctx, b, vo->iv_end, vo->n_iv * 8));
        V.iv_beg = a; V.iv_end = b;
    }
    if (vo->read_group) {
        V.have_rg = 1; V.rg_len = (uint32_t)strlen(vo->read_group);
        char *g; RC_TRY(T.alloc(&g, (size_t)V.rg_len + 1)); RC_TRY(copy_any(ctx, g, vo->read_group, (size_t)V.rg_len + 1));
        V.rg = g;
    }
    unsigned long long *d_first, h = ~0ull;
    RC_TRY(T.alloc(&d_first, 1));
    CUDA_TRY(cudaMemsetAsync(d_first, 0xff, 8, ctx->stream));
    LAUNCH(ctx, bam_first_key_k, grid_for(nr, 256), 256, 0, B->data, B->rec_off + r0, nr, V, (int64_t)key, d_first);
    LAUNCH_CHECK();
    CUDA_TRY(cudaMemcpyAsync(&h, d_first, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (h != ~0ull) { *offset = h; *found = 1; }
    return 0;
}

// (refid, 0-based POS) of the last complete record; *refid = -2 when there is none
extern "C" int wgbs_dbam_last_record(wgbs_ctx *ctx, const wgbs_dbam *B, int *refid, int64_t *pos) {
    RC_TRY(wgbs_ctx_activate(ctx));
    if (!B || !refid || !pos) return wgbs_set_err("wgbs_dbam_last_record: null argument");
    *refid = -2; *pos = -1;
    if (!B->nrec) return 0;
    uint64_t o = 0; uint8_t h[12];
    CUDA_TRY(cudaMemcpyAsync(&o, B->rec_off + (B->nrec - 1), 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(h, B->data + o, 12, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    *refid = ldi32(h + 4); *pos = ldi32(h + 8);
    return 0;
}

// General BGZF inflate on the device: bgzf = the bytes of a BGZF file (.bam, .pat.gz, CpG.bed.gz ...) in HOST memory;
// *dev_out = DEVICE buffer with the inflated bytes (release with wgbs_dev_free).
extern "C" int wgbs_bgzf_inflate(wgbs_ctx *ctx, const void *bgzf, size_t nbytes, void **dev_out, size_t *out_bytes) {
    RC_TRY(wgbs_ctx_activate(ctx));
    if ((!bgzf && nbytes) || !dev_out || !out_bytes) return wgbs_set_err("wgbs_bgzf_inflate: null argument");
    if (is_device_ptr(bgzf)) return wgbs_set_err("wgbs_bgzf_inflate: the compressed bytes must be in host memory (the block table is read on the host)");
    uint8_t *d = nullptr; uint64_t n = 0;
    RC_TRY(bgzf_inflate_device(ctx, bgzf, nbytes, "wgbs_bgzf_inflate", &d, &n, nullptr));
    *dev_out = d; *out_bytes = (size_t)n;
    return 0;
}

extern "C" int wgbs_dbam_open_file(wgbs_ctx *ctx, const char *path, wgbs_dbam **out) {
    RC_TRY(wgbs_ctx_activate(ctx));
    if (!path || !out) return wgbs_set_err("wgbs_dbam_open_file: null argument");
    FILE *f = fopen(path, "rb");
    if (!f) return wgbs_set_err("wgbs_dbam_open_file: cannot open %s", path);
    fseek(f, 0, SEEK_END); const long fsz = ftell(f); fseek(f, 0, SEEK_SET);
    void *pin = nullptr;
    if (cudaMallocHost(&pin, (size_t)fsz + 16) != cudaSuccess) { fclose(f); cudaGetLastError(); return wgbs_set_err("wgbs_dbam_open_file: cannot pin %ld bytes", fsz); }
    const size_t got = fsz ? fread(pin, 1, (size_t)fsz, f) : 0;
    fclose(f);
    int rc = got == (size_t)fsz ? wgbs_dbam_open(ctx, pin, (size_t)fsz, out) : wgbs_set_err("wgbs_dbam_open_file: short read on %s", path);
    cudaFreeHost(pin);
    if (rc < 0) { std::string m = g_wgbs_err; return wgbs_set_err("%s: %s", path, m.c_str()); }
    return rc;
}

extern "C" void wgbs_dbam_close(wgbs_ctx *ctx, wgbs_dbam *B) {
    if (!ctx || !B) return;
    cudaSetDevice(ctx->device);
    fail_free(ctx, B, 0);
}
extern "C" int wgbs_dbam_nref(const wgbs_dbam *B) { return B ? (int)B->ref_names.size() : -1; }
extern "C" const char *wgbs_dbam_ref_name(const wgbs_dbam *B, int i) { return (B && i >= 0 && i < (int)B->ref_names.size()) ? B->ref_names[i].c_str() : nullptr; }
extern "C" const char *wgbs_dbam_header(const wgbs_dbam *B) { return B ? B->header_text.c_str() : nullptr; }
extern "C" uint64_t wgbs_dbam_nrecords(const wgbs_dbam *B, int refid) {
    if (!B) return 0;
    if (refid < 0) return B->nrec;
    return refid < (int)B->ref_names.size() ? B->ref_last[refid] - B->ref_first[refid] : 0;
}
extern "C" uint64_t wgbs_dbam_inflated_bytes(const wgbs_dbam *B) { return B ? B->n : 0; }

// SAM text (DEVICE memory, release with wgbs_dev_free) of the records that pass the filters: wgbs_bam_view_ex on the device.
extern "C" int wgbs_dbam_view(wgbs_ctx *ctx, const wgbs_dbam *B, const wgbs_view_opts *vo, char **dev_text, size_t *nbytes, uint64_t *nrecords) {
    RC_TRY(wgbs_ctx_activate(ctx));
    if (!B || !vo || !dev_text || !nbytes) return wgbs_set_err("wgbs_dbam_view: null argument");
    *dev_text = nullptr; *nbytes = 0;
    if (vo->n_flag_eq < 0 || vo->n_flag_eq > 4) return wgbs_set_err("wgbs_dbam_view: n_flag_eq must be 0..4");
    if (vo->n_iv && (!vo->iv_beg || !vo->iv_end)) return wgbs_set_err("wgbs_dbam_view: interval list is null");
    if (vo->n_iv && (is_device_ptr(vo->iv_beg) || is_device_ptr(vo->iv_end))) return wgbs_set_err("wgbs_dbam_view: interval lists must be host arrays");
    for (size_t k = 0; k + 1 < vo->n_iv; k++)
        if (vo->iv_beg[k + 1] < vo->iv_end[k] || vo->iv_end[k] < vo->iv_beg[k]) return wgbs_set_err("wgbs_dbam_view: intervals must be sorted and non-overlapping");
    if (vo->max_records > 0xffffffffull) return wgbs_set_err("wgbs_dbam_view: max_records too large");
    uint64_t r0 = 0, r1 = B->nrec;
    if (vo->refid >= 0) { if (vo->refid >= (int)B->ref_names.size()) return wgbs_set_err("wgbs_dbam_view: no such reference"); r0 = B->ref_first[vo->refid]; r1 = B->ref_last[vo->refid]; }
    const uint64_t nr = r1 - r0;
    if (nr >= 0xffffffffull) return wgbs_set_err("wgbs_dbam_view: too many records in one call");
    Temps T(ctx);
    ViewParams V; memset(&V, 0, sizeof V);
    V.refid = vo->refid; V.min_mapq = vo->min_mapq; V.exclude_flags = vo->exclude_flags; V.include_flags = vo->include_flags; V.beg = vo->beg; V.end = vo->end;
    V.n_flag_eq = vo->n_flag_eq; for (int k = 0; k < 4; k++) V.flag_eq[k] = vo->flag_eq[k];
    V.n_iv = vo->n_iv; V.iv_exclude = vo->iv_exclude; V.key_beg = vo->key_beg; V.key_end = vo->key_end;
    if (vo->n_iv) {
        int64_t *a, *b;
        RC_TRY(T.alloc(&a, vo->n_iv)); RC_TRY(T.alloc(&b, vo->n_iv));
        RC_TRY(copy_any(ctx, a, vo->iv_beg, vo->n_iv * 8)); RC_TRY(copy_any(ctx, b, vo->iv_end, vo->n_iv * 8));
        V.iv_beg = a; V.iv_end = b;
    }
    if (vo->read_group) {
        V.have_rg = 1; V.rg_len = (uint32_t)strlen(vo->read_group);
        char *g; RC_TRY(T.alloc(&g, (size_t)V.rg_len + 1)); RC_TRY(copy_any(ctx, g, vo->read_group, (size_t)V.rg_len + 1));
        V.rg = g;
    }
    const DevRefs F{(int32_t)B->ref_names.size(), B->d_name_off, B->d_names, B->d_ref_lens};
    uint32_t *len, *pass; uint64_t *off; unsigned long long *too_long;
    RC_TRY(T.alloc(&len, nr)); RC_TRY(T.alloc(&pass, nr + 1)); RC_TRY(T.alloc(&off, nr + 1)); RC_TRY(T.alloc(&too_long, 1));
    CUDA_TRY(cudaMemsetAsync(too_long, 0xff, 8, ctx->stream));
    uint64_t npass = 0, tot = 0;
    if (nr) {
        LAUNCH(ctx, bam_measure_k, grid_for(nr, 256), 256, 0, B->data, B->rec_off + r0, nr, F, V, len, pass, too_long);
        uint32_t *rank; RC_TRY(T.alloc(&rank, nr + 1));
        RC_TRY(scan_u32_u32(ctx, pass, rank, nr));
        if (vo->max_records) LAUNCH(ctx, bam_head_k, grid_for(nr, 256), 256, 0, nr, rank, (uint32_t)vo->max_records, len);
        RC_TRY(scan_u32_u64(ctx, len, off, nr));
        uint32_t np32 = 0; unsigned long long tl = 0;
        CUDA_TRY(cudaMemcpyAsync(&np32, rank + nr, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(&tot, off + nr, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(&tl, too_long, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        if (tl != ~0ull) return wgbs_set_err("wgbs_dbam_view: record %llu formats to more than 4 GiB of text", tl + r0);
        npass = vo->max_records ? std::min<uint64_t>(np32, vo->max_records) : np32;
    }
    char *text = nullptr;
    RC_TRY(dalloc(ctx, &text, (size_t)tot + 16));
    if (tot) LAUNCH(ctx, bam_format_k, grid_for(nr, 256 / FMT_G), 256, 0, B->data, B->rec_off + r0, nr, F, len, off, text);
    LAUNCH_CHECK();
    *dev_text = text; *nbytes = (size_t)tot; if (nrecords) *nrecords = npass;
    return 0;
}

int bam_side_lines(wgbs_ctx *ctx, const BamSide &S, const uint32_t *ids, uint32_t m, uint32_t n, Temps &T, char **text, uint32_t **line_off,
                   uint32_t **line_len) {
    uint32_t *len; uint64_t *off;
    RC_TRY(T.alloc(&len, (size_t)m + 1)); RC_TRY(T.alloc(&off, (size_t)m + 1)); RC_TRY(T.alloc(line_off, n)); RC_TRY(T.alloc(line_len, n));
    LAUNCH(ctx, side_len_k, grid_for(m, 128), 128, 0, S, ids, m, len);
    RC_TRY(scan_u32_u64(ctx, len, off, m));
    uint64_t tot = 0;
    CUDA_TRY(cudaMemcpyAsync(&tot, off + m, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (tot >= 0xffffffffull) return wgbs_set_err("bam_side_lines: %llu bytes of SAM text for the multi-record QNAME groups", (unsigned long long)tot);
    RC_TRY(T.alloc(text, (size_t)tot + 16));
    LAUNCH(ctx, side_write_k, grid_for(m, 128), 128, 0, S, ids, m, len, off, *text, *line_off, *line_len);
    LAUNCH_CHECK();
    return 0;
}

int pileup_records(wgbs_ctx *ctx, const wgbs_index *ix, const ReadBatch &rb, const wgbs_pileup_opts *opts, Temps &T, wgbs_pats **out,
                   uint64_t *stats_out, int32_t *mbias_out);      // pileup.cu

// The direct route: filters -> descriptors of the passing records (no SAM text is formatted or tokenized); the pileup kernels
// read binary CIGARs, 4-bit bases and (MM/ML mode) the MM string / ML array in place.  Returns 1 when the batch needs the text route
// (a window of the stream wider than 32-bit offsets), 0 when done.
static int pileup_dbam_direct(wgbs_ctx *ctx, const wgbs_index *ix, const wgbs_dbam *B, const wgbs_view_opts *vo, const wgbs_pileup_opts *opts,
                              wgbs_pats **out, uint64_t *stats, int32_t *mbias) {
    if (vo->n_flag_eq < 0 || vo->n_flag_eq > 4) return wgbs_set_err("wgbs_pileup_dbam: n_flag_eq must be 0..4");
    if (vo->n_iv && (!vo->iv_beg || !vo->iv_end)) return wgbs_set_err("wgbs_pileup_dbam: interval list is null");
    if (vo->n_iv && (is_device_ptr(vo->iv_beg) || is_device_ptr(vo->iv_end))) return wgbs_set_err("wgbs_pileup_dbam: interval lists must be host arrays");
    for (size_t k = 0; k + 1 < vo->n_iv; k++)
        if (vo->iv_beg[k + 1] < vo->iv_end[k] || vo->iv_end[k] < vo->iv_beg[k]) return wgbs_set_err("wgbs_pileup_dbam: intervals must be sorted and non-overlapping");
    if (vo->max_records > 0xffffffffull) return wgbs_set_err("wgbs_pileup_dbam: max_records too large");
    uint64_t r0 = 0, r1 = B->nrec;
    if (vo->refid >= 0) { if (vo->refid >= (int)B->ref_names.size()) return wgbs_set_err("wgbs_pileup_dbam: no such reference"); r0 = B->ref_first[vo->refid]; r1 = B->ref_last[vo->refid]; }
    const uint64_t nr = r1 - r0;
    if (nr >= 0xffffffffull) return wgbs_set_err("wgbs_pileup_dbam: too many records in one call");
    Temps T(ctx);
    ViewParams V; memset(&V, 0, sizeof V);
    V.refid = vo->refid; V.min_mapq = vo->min_mapq; V.exclude_flags = vo->exclude_flags; V.include_flags = vo->include_flags; V.beg = vo->beg; V.end = vo->end;
    V.n_flag_eq = vo->n_flag_eq; for (int k = 0; k < 4; k++) V.flag_eq[k] = vo->flag_eq[k];
    V.n_iv = vo->n_iv; V.iv_exclude = vo->iv_exclude; V.key_beg = vo->key_beg; V.key_end = vo->key_end;
    if (vo->n_iv) {
        int64_t *a, *b;
        RC_TRY(T.alloc(&a, vo->n_iv)); RC_TRY(T.alloc(&b, vo->n_iv));
        RC_TRY(copy_any(ctx, a, vo->iv_beg, vo->n_iv * 8)); RC_TRY(copy_any(ctx, b, vo->iv_end, vo->n_iv * 8));
        V.iv_beg = a; V.iv_end = b;
    }
    if (vo->read_group) {
        V.have_rg = 1; V.rg_len = (uint32_t)strlen(vo->read_group);
        char *g; RC_TRY(T.alloc(&g, (size_t)V.rg_len + 1)); RC_TRY(copy_any(ctx, g, vo->read_group, (size_t)V.rg_len + 1));
        V.rg = g;
    }
    uint32_t *pass, *rank, *has_mm = ctx->d_flags + 24;
    RC_TRY(T.alloc(&pass, nr + 1)); RC_TRY(T.alloc(&rank, nr + 1));
    CUDA_TRY(cudaMemsetAsync(has_mm, 0, 4, ctx->stream));
    uint32_t np32 = 0; uint64_t span[2] = {0, 0};
    if (nr) {
        LAUNCH(ctx, bam_pass_k, grid_for(nr, 256), 256, 0, B->data, B->rec_off + r0, nr, V, pass);
        RC_TRY(scan_u32_u32(ctx, pass, rank, nr));
        CUDA_TRY(cudaMemcpyAsync(&np32, rank + nr, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(&span[0], B->rec_off + r0, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        span[1] = r1 < B->nrec ? 0 : B->n;                       // end of the window: start of the record after the range, or the stream's end
        if (r1 < B->nrec) { CUDA_TRY(cudaMemcpyAsync(&span[1], B->rec_off + r1, 8, cudaMemcpyDeviceToHost, ctx->stream)); CUDA_TRY(cudaStreamSynchronize(ctx->stream)); }
    }
    const uint32_t n = vo->max_records ? (uint32_t)std::min<uint64_t>(np32, vo->max_records) : np32;
    if (span[1] - span[0] >= 0xfffffff0ull && np32) {
        // a large chromosome seen through a template window (bam2pat piles it up in windows): the passing records lie close
        // together, so the window of the stream the 32-bit descriptors index is narrowed to them
        uint32_t *fl = ctx->d_flags + 26; const uint32_t init[2] = {0xffffffffu, 0u}; uint32_t hfl[2] = {0, 0};
        CUDA_TRY(cudaMemcpyAsync(fl, init, 8, cudaMemcpyHostToDevice, ctx->stream));
        LAUNCH(ctx, pass_range_k, grid_for(nr, 256), 256, 0, pass, nr, fl);
        CUDA_TRY(cudaMemcpyAsync(hfl, fl, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(&span[0], B->rec_off + r0 + hfl[0], 8, cudaMemcpyDeviceToHost, ctx->stream));
        if (r0 + hfl[1] + 1 < B->nrec) CUDA_TRY(cudaMemcpyAsync(&span[1], B->rec_off + r0 + hfl[1] + 1, 8, cudaMemcpyDeviceToHost, ctx->stream));
        else span[1] = B->n;
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    }
    const uint64_t base = span[0];
    if (span[1] - base >= 0xfffffff0ull) return 1;                // still too wide: the text route formats only the passing records
    ReadBatch rb;
    rb.text = (const char *)B->data + base; rb.nbytes = (uint32_t)(span[1] - base); rb.n = n; rb.bam = 1;
    RC_TRY(T.alloc(&rb.line_off, n)); RC_TRY(T.alloc(&rb.line_len, n)); RC_TRY(T.alloc(&rb.qn_len, n));
    RC_TRY(T.alloc(&rb.flag, n)); RC_TRY(T.alloc(&rb.pos, n)); RC_TRY(T.alloc(&rb.pos_hi, n));
    RC_TRY(T.alloc(&rb.cig_off, n)); RC_TRY(T.alloc(&rb.cig_len, n)); RC_TRY(T.alloc(&rb.seq_off, n)); RC_TRY(T.alloc(&rb.seq_len, n));
    RC_TRY(T.alloc(&rb.hash_lo, n)); RC_TRY(T.alloc(&rb.hash_hi, n)); RC_TRY(T.alloc(&rb.status, n));
    uint64_t *rec_abs; RC_TRY(T.alloc(&rec_abs, n));
    uint32_t mm = 0;
    if (n) {
        LAUNCH(ctx, bam_records_k, grid_for(nr, 256), 256, 0, B->data, B->rec_off + r0, nr, pass, rank, (uint32_t)vo->max_records, base, rb, rec_abs, has_mm);
        CUDA_TRY(cudaMemcpyAsync(&mm, has_mm, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        LAUNCH_CHECK();
    }
    if (mm || opts->nanopore) {                                  // MM/ML mode (forced, or the first record has an MM tag: patter.cpp:337-338)
        RC_TRY(T.alloc(&rb.mm_off, n)); RC_TRY(T.alloc(&rb.mm_len, n)); RC_TRY(T.alloc(&rb.ml_off, n)); RC_TRY(T.alloc(&rb.ml_len, n));
        if (n) { LAUNCH(ctx, bam_np_tags_k, grid_for(n, 256), 256, 0, B->data, rec_abs, n, base, rb); LAUNCH_CHECK(); }
    }
    BamSide side{B->data, rec_abs, (int32_t)B->ref_names.size(), B->d_name_off, B->d_names, B->d_ref_lens};
    rb.side = &side;
    RC_TRY(pileup_records(ctx, ix, rb, opts, T, out, stats, mbias));
    return 0;
}

// `samtools view ... | [match_maker |] patter ...` without leaving the device.  Direct route (WGBS_DBAM_DIRECT=1, or the
// default below): BAM records feed the pileup kernels as they are.  Text route: wgbs_dbam_view + wgbs_pileup_sam_mbias.
#ifndef WGBS_DBAM_DIRECT_DEFAULT
#define WGBS_DBAM_DIRECT_DEFAULT 1      // measured (round 1 driver run): 7.08 ms vs 8.30 ms per 1M-read step, identical outputs
#endif
extern "C" int wgbs_pileup_dbam(wgbs_ctx *ctx, const wgbs_index *ix, const wgbs_dbam *B, const wgbs_view_opts *vo, const wgbs_pileup_opts *opts,
                                wgbs_pats **out, uint64_t *stats, int32_t *mbias) {
    RC_TRY(wgbs_ctx_activate(ctx));
    if (!ix || !B || !vo || !opts || !out) return wgbs_set_err("wgbs_pileup_dbam: null argument");
    *out = nullptr;
    const char *e = getenv("WGBS_DBAM_DIRECT");                  // read per call: tests switch routes inside one process
    const int direct = e ? (e[0] == '1') : (WGBS_DBAM_DIRECT_DEFAULT != 0);
    if (direct) {
        const int rc = pileup_dbam_direct(ctx, ix, B, vo, opts, out, stats, mbias);
        if (rc <= 0) return rc;                                  // done, or failed; 1: the batch needs the text route
    }
    char *text = nullptr; size_t nb = 0;
    RC_TRY(wgbs_dbam_view(ctx, B, vo, &text, &nb, nullptr));
    if (nb >= 0xffffffffull) { dfree(ctx, text); return wgbs_set_err("wgbs_pileup_dbam: %zu bytes of SAM text in one call (limit 4 GiB): restrict the region", nb); }
    const int rc = wgbs_pileup_sam_mbias(ctx, ix, text, nb, opts, out, stats, mbias);
    dfree(ctx, text);
    return rc;
}
