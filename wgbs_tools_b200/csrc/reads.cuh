// reads.cuh -- device-side description of a batch of alignment records and of the templates built from them.
#pragma once
#include <stdint.h>

#include "common.cuh"
#include "pats.cuh"

// status of a record
enum : uint8_t { REC_OK = 0, REC_INVALID = 1 /* < 11 fields */, REC_BLANK = 2 /* empty input line: counted, never processed */,
                 REC_BADINT = 3 /* FLAG / POS not numeric: std::stoi throws when the read is processed */ };

// One entry per SAM line.  All *_off are byte offsets into the SAM text buffer (which stays resident: the pileup
// reads CIGAR and SEQ bytes straight from it, nothing is re-packed).
// Two flavours of batch share the descriptors:
//   text (bam == 0): `text` is SAM text; cig_off/cig_len and seq_off/seq_len delimit the CIGAR and SEQ strings.
//   BAM  (bam == 1): `text` is a window of the inflated BAM stream (bamdev.cu); line_off = offset of read_name, cig_off = offset
//        of the binary CIGAR with cig_len = n_cigar_op, seq_off = offset of the 4-bit SEQ with seq_len = l_seq -- except SEQ "*"
//        (l_seq 0), which the text path sees as ONE character: seq_len 1, seq_off SEQ_STAR.  Nothing is unpacked: the pileup
//        kernels read ops and bases through rec_* accessors (pileup_dev.cuh).  MM/ML batches always take the text flavour.
constexpr uint32_t SEQ_STAR = 0xffffffffu;
struct BamSide;                  // what build_mates needs to order 3+-record QNAME groups of a BAM batch by their SAM lines
struct ReadBatch {
    const char *text = nullptr;  // device
    uint32_t nbytes = 0;
    uint32_t n = 0;              // records (lines)
    int bam = 0;
    const BamSide *side = nullptr;   // host struct, BAM flavour only
    uint32_t *line_off = nullptr, *line_len = nullptr;
    uint32_t *qn_len = nullptr;
    int32_t *flag = nullptr, *pos = nullptr, *pos_hi = nullptr;   // POS is parsed as 64 bits (stoul): low word (as int) / high word
    uint32_t *cig_off = nullptr, *cig_len = nullptr, *seq_off = nullptr, *seq_len = nullptr;
    uint32_t *hash_lo = nullptr, *hash_hi = nullptr;   // 64-bit QNAME hash (pairing keys; names are still byte-verified)
    uint32_t *mm_off = nullptr, *mm_len = nullptr, *ml_off = nullptr, *ml_len = nullptr;  // MM:Z: / ML:B:C payloads (len 0: absent)
    uint8_t *status = nullptr;
};

struct ReadBatchView {
    const char *text; uint32_t nbytes, n; int bam;
    const uint32_t *line_off, *line_len, *qn_len;
    const int32_t *flag, *pos, *pos_hi;
    const uint32_t *cig_off, *cig_len, *seq_off, *seq_len, *hash_lo, *hash_hi, *mm_off, *mm_len, *ml_off, *ml_len;
    const uint8_t *status;
};
static inline ReadBatchView view_of(const ReadBatch &b) {
    return ReadBatchView{b.text, b.nbytes, b.n, b.bam, b.line_off, b.line_len, b.qn_len, b.flag, b.pos, b.pos_hi, b.cig_off, b.cig_len, b.seq_off,
                         b.seq_len, b.hash_lo, b.hash_hi, b.mm_off, b.mm_len, b.ml_off, b.ml_len, b.status};
}

// CpG dictionary of one chromosome / region: sorted 1-based loci; CpG index of loci[k] is first_idx + k
// (replaces patter's unordered_map + bool conv[chromlen], reference pipeline_wgbs/patter.cpp:14-42)
struct wgbs_index {
    uint32_t *loci = nullptr;  // device
    uint32_t n = 0;
    uint32_t first_idx = 1;
};

// counters (reference patter.h:28-34 reads_stats + line counter)
enum { ST_LINES = 0, ST_PAIRS, ST_EMPTY, ST_SHORT, ST_INVALID, ST_TEMPLATES, ST_N = 8 };

// records of a BAM batch, for the rare whole-line ordering (pair.cu): filled by bamdev.cu
struct BamSide {
    const uint8_t *data;            // inflated stream (device)
    const uint64_t *rec;            // per batch record: absolute offset of its block_size field (device)
    int32_t n_ref; const uint32_t *name_off; const char *names; const int32_t *ref_lens;   // device
};
// SAM lines of the batch records ids[0..m) (device ids) into one text buffer; line_off / line_len are indexed by RECORD id (n entries,
// only the m listed ones are filled).  bamdev.cu
int bam_side_lines(wgbs_ctx *ctx, const BamSide &S, const uint32_t *ids, uint32_t m, uint32_t n, Temps &T, char **text, uint32_t **line_off,
                   uint32_t **line_len);

// QNAME hash shared by the SAM tokenizer (sam.cu) and the BAM record kernel (bamdev.cu): per-(byte, position) mixing summed over
// the name, so it does not depend on how the name is aligned or chunked
__device__ __forceinline__ uint64_t fmix64(uint64_t k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
    return k;
}
__device__ __forceinline__ uint64_t name_byte_mix(uint32_t byte, uint32_t pos) {
    uint64_t x = ((uint64_t)(byte | (pos << 8)) + 1) * 0x9e3779b97f4a7c15ULL;
    x ^= x >> 29; x *= 0xbf58476d1ce4e5b9ULL; x ^= x >> 32;
    return x;
}

// sam.cu
// want_tags: 1 record MM/ML tag spans, 0 do not, -1 decide from the first line (rb->mm_off != nullptr tells the outcome)
int sam_tokenize(wgbs_ctx *ctx, const char *dtext, size_t nbytes, int want_tags, Temps &T, ReadBatch *out);
// pair.cu: mate[r] = record id of the mate of r (0xffffffff: none)
int build_mates(wgbs_ctx *ctx, const ReadBatch &rb, bool paired, Temps &T, uint32_t **mate_out,
                unsigned long long *d_stats);
