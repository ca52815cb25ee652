/*
 * wgbs_b200.h -- C ABI of the B200-native wgbstools hot path.
 *
 * The reference (nloyfer/wgbs_tools v0.3.0) has no in-process FFI: its hot path sits behind PROCESS boundaries
 * (argv + text on stdin/stdout of five small executables).  Every entry point below replaces
 * one of those executables (or the coreutils / numpy step next to it) and documents the reference interface it
 * stands in for.  INTEGRATION.md shows the ctypes binding a reference maintainer would add to
 * src/python/{bam2pat,pat2beta,homog,segment}.py.
 *
 * Conventions
 *   - plain C, no torch types.  All functions return 0 on success, <0 on error; wgbs_last_error() (thread-local)
 *     has the message.  Bad READS are never errors: they are skipped and counted (reference patter.cpp:229-244).
 *   - one wgbs_ctx per GPU (device ordinal + stream + stream-ordered scratch).  Calls on one ctx are serialised by the
 *     caller; different ctxs are independent (mirrors the reference's one-process-per-chromosome model).
 *   - every data pointer may be a HOST or a DEVICE pointer unless stated otherwise; the library detects which
 *     (cudaPointerGetAttributes) and stages host buffers through pinned memory.  The caller owns all buffers.
 *   - functions are asynchronous on ctx's stream when all their buffers are device buffers; any host output is
 *     complete when the function returns.
 *   - there is NO CPU fallback: every entry point fails (rc<0) when no CUDA device is usable.
 */
#ifndef WGBS_B200_H
#define WGBS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WGBS_B200_ABI_VERSION 5

typedef struct wgbs_ctx wgbs_ctx;
typedef struct wgbs_pats wgbs_pats;   /* device-resident pat records: (idx, len, count, 2-bit symbol pool) */
typedef struct wgbs_index wgbs_index; /* device-resident CpG dictionary of one chromosome / region */

/* ---------------------------------------------------------------------------------------------------------------
 * context
 * ------------------------------------------------------------------------------------------------------------- */
int wgbs_abi_version(void);
const char *wgbs_last_error(void);
/* stream: a cudaStream_t to run on (e.g. torch.cuda.current_stream().cuda_stream), or NULL to create one. */
wgbs_ctx *wgbs_create(int device, void *stream);
void wgbs_destroy(wgbs_ctx *);
int wgbs_sync(wgbs_ctx *);
/* number of kernels this ctx has launched so far (bench.py's "gpu_launches") */
uint64_t wgbs_launch_count(const wgbs_ctx *);
/* per-kernel device timing (CUDA events on ctx's stream around every launch).  report: "kernel \t launches \t total_ms"
 * lines for the launches since wgbs_prof_enable(ctx, 1); returns the number of bytes written. */
int wgbs_prof_enable(wgbs_ctx *, int on);
int wgbs_prof_report(wgbs_ctx *, char *buf, size_t cap);
/* device memory helpers so that callers without torch can keep inputs resident in HBM */
int wgbs_dev_alloc(wgbs_ctx *, size_t nbytes, void **dptr);
int wgbs_dev_free(wgbs_ctx *, void *dptr);
int wgbs_memcpy(wgbs_ctx *, void *dst, const void *src, size_t nbytes); /* any host/device combination */
/* streaming inputs: begin the host->device copy of the NEXT batch (pinned host memory) on the context's copy stream while
 * the current batch is being processed; wgbs_prefetch_wait orders the following calls on this context after that copy.
 * (The reference streams chromosome after chromosome through its pipes, bam2pat.py:319-346; this is the same overlap.) */
int wgbs_prefetch(wgbs_ctx *, void *dev_dst, const void *host_src, size_t nbytes);
int wgbs_prefetch_wait(wgbs_ctx *);

/* ---------------------------------------------------------------------------------------------------------------
 * pat records  (on-disk format: reference docs/pat_format.md:3-47  "chr \t idx \t pattern \t count")
 * Symbol codes (2 bit): '.'=0 'C'=1 'H'=2 'T'=3 -- the C-locale collation order the reference's `sort -k3,3` uses.
 * Pool layout: record r owns words pool[off[r] .. off[r]+ceil(len[r]/16)), 16 symbols per uint32, first symbol in
 * the two MOST significant bits, unused tail bits zero.  Records need not be contiguous in the pool (off has n entries).
 * ------------------------------------------------------------------------------------------------------------- */
/* Parse pat TEXT (what `gunzip -c X.pat.gz` prints; stdin of reference stdin2beta.cpp:95-123 / homog.cpp:262-313)
 * on the GPU.  Lines with < 4 columns or a non-numeric idx/count make the call fail, like the reference
 * ("failed calculating beta": stdin2beta.cpp:104-106,118-122).  Empty lines are skipped. */
int wgbs_pats_from_text(wgbs_ctx *, const char *text, size_t nbytes, wgbs_pats **out);
int wgbs_pats_count(const wgbs_pats *, uint64_t *n_records, uint64_t *n_pool_words);
/* copy the SoA out (any pointer may be NULL to skip that array) */
int wgbs_pats_download(wgbs_ctx *, const wgbs_pats *, uint32_t *idx, uint32_t *len, uint32_t *count, uint32_t *off,
                       uint32_t *pool);
void wgbs_pats_free(wgbs_ctx *, wgbs_pats *);

/* ---------------------------------------------------------------------------------------------------------------
 * pat -> beta   (replaces `stdin2beta START END`, reference src/pat2beta/stdin2beta.cpp:59-93, and
 *                utils_wgbs.trim_to_uint8, reference src/python/utils_wgbs.py:277-290)
 * ------------------------------------------------------------------------------------------------------------- */
/* meth_cov: DEVICE int32[end-start, 2] rows (meth, cover) -- the numbers stdin2beta prints.  zero_first!=0 clears
 * it first; otherwise the records are ADDED (multi-batch / multi-GPU partial sums). start is 1-based, end exclusive. */
int wgbs_pat2beta(wgbs_ctx *, const wgbs_pats *, uint32_t start, uint32_t end, int32_t *meth_cov, int zero_first);
/* rows with cover > max (255 | 65535): meth = trunc(float64(meth)/float64(cover)*max), cover = max; cast to
 * uint8 / uint16 (nbits 8 | 16).  out: [n,2] of that type.  meth_cov / out: host or device. */
int wgbs_trim(wgbs_ctx *, const int32_t *meth_cov, size_t n, int nbits, void *out);
/* One call, host buffers: pat text -> .beta bytes (what pat2beta.py:32-37 writes).  meth_cov_out (int32[n,2], host)
 * may be NULL. */
int wgbs_pat2beta_text(wgbs_ctx *, const char *text, size_t nbytes, uint32_t start, uint32_t end, int nbits,
                       void *beta_out, int32_t *meth_cov_out);

/* ---------------------------------------------------------------------------------------------------------------
 * homog   (replaces `homog -b BLOCKS -r r0,..,rk -l MINLEN [--inclusive]`, reference src/homog/homog.cpp:154-260)
 * blocks: startCpG/endCpG (end exclusive) in the order the reference processes them (file order, which must be
 * sorted by startCpG; the host side applies --sort_blocks / --chrom).  range: nbins+1 float32 edges as parsed by
 * `istream >> float`.  out: int32[nblocks, nbins], one row per block in block order (what homog prints).
 * Records must be sorted by idx (a pat file is, by definition).
 * ------------------------------------------------------------------------------------------------------------- */
int wgbs_homog(wgbs_ctx *, const wgbs_pats *, const int32_t *bstart, const int32_t *bend, size_t nblocks,
               const float *range, int nbins, int min_cpgs, int inclusive, int32_t *out);

/* ---------------------------------------------------------------------------------------------------------------
 * bam -> pat pileup
 *   replaces   [match_maker |] patter CPG_DICT REGION [--min_cpg N] [--clip N] [--nanopore --np_thresh F
 *              --cpc_call C|H|. --combine_mods]            (reference src/pipeline_wgbs/{match_maker,main,patter,ont}.cpp)
 *   and        sort -k2,2n -k3,3 | uniq -c | awk '{print $2,$3,$4,$1}'   (reference src/python/bam2pat.py:99-106)
 * ------------------------------------------------------------------------------------------------------------- */
/* CpG dictionary of one chromosome / region: `tabix CpG.bed.gz REGION | cut -f2-3` (reference patter.cpp:14-42).
 * loci: sorted 1-based positions of the C of every CpG in the region; the CpG index of loci[k] is first_idx + k. */
int wgbs_index_load(wgbs_ctx *, const uint32_t *loci, size_t n, uint32_t first_idx, wgbs_index **out);
void wgbs_index_free(wgbs_ctx *, wgbs_index *);

typedef struct wgbs_pileup_opts {
    int32_t min_cpg;      /* --min_cpg  (patter main.cpp:16-22, default 1): drop templates whose pattern is shorter */
    int32_t clip;         /* --clip     (main.cpp:23-29, default 0): ignore calls in the first/last `clip` positions */
    int32_t paired;       /* 1 / 0, or -1 = decide from FLAG&1 of the first line like patter.cpp:324-333 */
    int32_t nanopore;     /* --nanopore (MM/ML mode); also switched on when the first line carries an MM tag */
    int32_t combine_mods; /* --combine_mods */
    float np_thresh;      /* --np_thresh (float32, default 0.67) */
    char cpc_call;        /* --cpc_call 'C' | 'H' | '.' (0 = 'C') */
    int32_t keep_names;   /* --long (main.cpp:37): keep every template's read name for wgbs_collapse_long / wgbs_pats_format_long */
} wgbs_pileup_opts;

/* sam: SAM text without header, one chromosome, coordinate sorted -- exactly what `samtools view BAM chr` feeds the
 * reference pipeline (host or device pointer, < 4 GiB per call).  Mates are found by QNAME (match_maker's job).
 * out: one record per template with >= min_cpg symbols, count = 1, in input order of the template's first record.
 * stats[8]: lines, pairs, empty, short ("too few CpGs"), invalid, paired(0/1), nanopore(0/1), templates out
 *           -- the counters of patter's summary line (patter.cpp:298-316). */
int wgbs_pileup_sam(wgbs_ctx *, const wgbs_index *, const char *sam, size_t nbytes, const wgbs_pileup_opts *,
                    wgbs_pats **out, uint64_t *stats);
/* Same, plus the M-bias tables of `patter --mbias` (reference patter.cpp:50-72,116-165): mbias = int32[2][2][1000][2]
 * indexed [OT=0|OB=1][mate 1|2][read position][meth=0|unmeth=1] (host or device; NULL = off), zeroed by the call.  The
 * reference's `<path>.OT.txt` / `.OB.txt` rows are `r1m r1u r2m r2u` per position.  Ignored in MM/ML mode, like the reference. */
int wgbs_pileup_sam_mbias(wgbs_ctx *, const wgbs_index *, const char *sam, size_t nbytes, const wgbs_pileup_opts *,
                          wgbs_pats **out, uint64_t *stats, int32_t *mbias);
/* sort by (idx, pattern) in the C locale and merge identical records, summing counts (in place) */
int wgbs_collapse(wgbs_ctx *, wgbs_pats *);
/* --long variant (reference bam2pat.py:102-103: `sort -k2,2n -k3,3 | awk '{print $1,$2,$3,1,$4}'`, no uniq): records are only
 * ordered -- by (idx, pattern, read name), the read name being what `sort`'s last-resort whole-line comparison sees -- and
 * wgbs_pats_format_long writes "chrom \t idx \t pattern \t 1 \t qname \n".  Needs opts.keep_names. */
int wgbs_collapse_long(wgbs_ctx *, wgbs_pats *);
int wgbs_pats_format_long(wgbs_ctx *, const wgbs_pats *, const char *chrom, char *out, size_t cap, size_t *nbytes);
/* the collapse with its variants spelled out.  DOTTED: patterns may begin/end with '.' (cview output without --strip):
 * "C" sorts before "C." as `sort -k3,3` has it.  ADJACENT: no sort, only neighbouring identical records merge -- what
 * reference src/collapse_pat.pl does to the unsorted stream of a whole-file `wgbstools view` (cview.py:29-51). */
#define WGBS_COLLAPSE_SORTED 0
#define WGBS_COLLAPSE_LONG 1
#define WGBS_COLLAPSE_DOTTED 2
#define WGBS_COLLAPSE_ADJACENT 3
int wgbs_collapse_ex(wgbs_ctx *, wgbs_pats *, int mode);
/* "chrom \t idx \t pattern \t count \n" per record, record order.  out NULL: only *nbytes. out: host or device. */
int wgbs_pats_format(wgbs_ctx *, const wgbs_pats *, const char *chrom, char *out, size_t cap, size_t *nbytes);
/* utility: stable radix sort of (key, value) uint32 pairs, device pointers */
int wgbs_sort_pairs_u32(wgbs_ctx *, uint32_t *keys, uint32_t *vals, size_t n);

/* ---------------------------------------------------------------------------------------------------------------
 * segment   (replaces `segmentor B1.beta .. BK.beta -s START0 -n NSITES -max_cpg M -max_bp B -ps P < loci`,
 *            reference src/segment_betas/{main,segmentor}.cpp; one call solves MANY chunks / stitching patches)
 * betas: K pointers (host array) to uint8[nsites,2] beta arrays covering one site range (host or device memory;
 *        device-resident arrays are used in place).  dists: uint32[nsites] bp locus of each site (the `tabix rev.CpG.bed.gz`
 *        column the reference pipes to stdin, segment.py:53), non-decreasing inside a chunk.
 * chunks: [start, start+n) site ranges relative to the arrays; each is an independent DP (segment.py:129-134).
 * borders: int32, (n_c + 1) slots per chunk in chunk order; chunk c's borders (ascending, relative to its start,
 *        first 0, last n_c -- exactly the line segmentor prints) occupy the first nborders[c] slots of its region.
 * Errors: meth > cover anywhere (reference: `throw 0`), max_cpg >= 16384.
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct wgbs_chunk { uint32_t start, n; } wgbs_chunk;
int wgbs_segment(wgbs_ctx *, const uint8_t *const *betas, int K, const uint32_t *dists, size_t nsites,
                 const wgbs_chunk *chunks, int nchunks, int max_cpg, uint32_t max_bp, float pseudo,
                 int32_t *borders, int32_t *nborders);
/* ---------------------------------------------------------------------------------------------------------------
 * Consumers next to the hot path
 * ------------------------------------------------------------------------------------------------------------- */
/* cview (reference src/cview/cview.cpp:87-167; argv `--sites "s\te"` / `--blocks_path F [--strict] [--strip] [--no_gaps]
 * [--min_cpgs N]`): the records of `in` (in file order = sorted by start) that overlap the blocks [bstart, bend) (CpG
 * indices, end exclusive, sorted by start the way cview's `sort -k1,1n` leaves them; host or device int32), unchanged
 * or -- with strict -- cut into one piece per overlapped block (blocks must then be disjoint: the reference aborts on
 * overlapping ones).  strip removes leading/trailing '.', no_gaps drops pieces containing '.', min_cpgs drops pieces
 * shorter than that.  pre_lo/pre_hi (npre closed ranges of START indices, sorted, disjoint; npre 0 = none) restate the
 * `tabix pat chr:lo-hi` pre-selection in front of cview (cview.py:40, :88-96).  *out: new records (count copied), in the
 * reference's output order; follow with wgbs_collapse_ex + wgbs_pats_format for the `| sort | collapse_pat.pl` tail. */
int wgbs_cview(wgbs_ctx *, const wgbs_pats *in, const int32_t *bstart, const int32_t *bend, size_t nblocks,
               const int32_t *pre_lo, const int32_t *pre_hi, size_t npre, int strict, int strip, int no_gaps, int min_cpgs,
               wgbs_pats **out);
/* beta_to_blocks (reference src/python/beta_to_blocks.py:101-126 + utils_wgbs.py:277-290): for every block the sums of
 * (meth, cover) over rows startCpG-1 .. endCpG-2 of a beta file (uint8 pairs: in_bits 8; .lbeta uint16 pairs: 16),
 * clamped to the file like a numpy slice; an empty / NA block (bend <= bstart) gives (0, 0).  out_sums: int64[nblocks,2]
 * (optional); out_trimmed: the `.bin` / `.lbeta` rows (trim_to_uint8 of the sums; out_bits 8 or 16; optional). Host or device. */
int wgbs_beta_to_blocks(wgbs_ctx *, const void *beta, int in_bits, size_t nsites, const int32_t *bstart, const int32_t *bend,
                        size_t nblocks, int out_bits, void *out_trimmed, int64_t *out_sums);

/* numerics self-test: log2f(p[i]) and log2(1.0 - (double)p[i]) exactly as glibc 2.39 (FMA build) computes them */
int wgbs_glibc_log2_probe(wgbs_ctx *, const float *p, size_t n, float *out_log2f, double *out_log2_1mp);

/* ---------------------------------------------------------------------------------------------------------------
 * BAM ingest (host side; replaces the `samtools view BAM chr -q Q -F X [-f Y]` stage of reference bam2pat.py:165 where
 * samtools is not available).  BGZF blocks are inflated on `threads` host threads (0 = all cores); the file must be
 * coordinate sorted.  wgbs_bam_view returns, malloc'ed (release with wgbs_host_free), the SAM text samtools view would
 * print for reference `refid` (-1 = all records), optionally restricted to the 1-based closed interval beg..end
 * (end <= 0: whole reference).  No GPU is needed for these calls.
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct wgbs_bam wgbs_bam;
int wgbs_bam_open(const char *path, int threads, wgbs_bam **out);
void wgbs_bam_close(wgbs_bam *);
int wgbs_bam_nref(const wgbs_bam *);
const char *wgbs_bam_ref_name(const wgbs_bam *, int i);
const char *wgbs_bam_header(const wgbs_bam *);
uint64_t wgbs_bam_nrecords(const wgbs_bam *, int refid);
int wgbs_bam_view(const wgbs_bam *, int refid, int min_mapq, int exclude_flags, int include_flags, int64_t beg, int64_t end,
                  char **text, size_t *nbytes, uint64_t *nrecords);

/* The full `samtools view` stage of reference bam2pat.py:126-159, one pass over the decoded records:
 *   samtools view BAM region -q Q -F X [-f Y] [| awk '($2 == A || $2 == B)'] [-r RG] [-M -L whitelist.bed]
 *   [... -b | bedtools intersect -sorted -v -abam stdin -b blacklist.bed | samtools view]        [| head -N]
 * Intervals are 0-based half-open [iv_beg, iv_end) on reference `refid`, sorted by start and non-overlapping (merge
 * them first); a record is tested by its reference span POS..POS+span (span from the CIGAR, at least 1). */
typedef struct wgbs_view_opts {
    int refid;                  /* -1: every record */
    int min_mapq, exclude_flags, include_flags;
    int64_t beg, end;           /* 1-based closed region; end <= 0: whole reference */
    int n_flag_eq;              /* 0..4: keep only records whose FLAG equals one of flag_eq[] (--top_strand / --bottom_strand) */
    int flag_eq[4];
    const char *read_group;     /* keep only records whose RG:Z: tag equals this (samtools view -r); NULL: no filter */
    const int64_t *iv_beg, *iv_end;
    size_t n_iv;                /* 0: no interval filter */
    int iv_exclude;             /* 0: keep records overlapping an interval (-L); 1: keep records overlapping none (bedtools -v) */
    uint64_t max_records;       /* 0: all; else only the first max_records passing records (is_pair_end / detect_nanopore peeks) */
    /* Template window (key_end <= 0: none): keep the records whose TEMPLATE KEY lies in the 0-based half-open window
     * [key_beg, key_end).  key = POS, or max(POS, PNEXT) for a paired record whose mate is mapped to the same reference -- so
     * both mates of a pair carry the same key and a chromosome can be piled up in windows (a 30x chromosome is far more than
     * the 4 GiB of SAM text / BAM stream one call takes) without ever separating mates: the windows partition the records.
     * This is match_maker's own rule for what may still find its mate (PNEXT against the last position seen,
     * reference match_maker.cpp:77-88), applied at window granularity. */
    int64_t key_beg, key_end;
} wgbs_view_opts;
int wgbs_bam_view_ex(const wgbs_bam *, const wgbs_view_opts *, char **text, size_t *nbytes, uint64_t *nrecords);
void wgbs_host_free(void *);

/* A .bam that does not fit in memory as a whole is read as a sequence of PARTS (bam2pat --bam_decode stream; samtools streams
 * the file the same way, reference bam2pat.py:165).  bgzf = the bytes of consecutive WHOLE BGZF blocks of the file.
 * has_header != 0: the part begins with the BAM header (the first part of a file); else the reference list is passed in
 * (n_ref names / lengths, from the first part) and records are indexed from inflated offset first_record on.  The last record
 * may be cut off by the end of the part: *tail = inflated offset of the first byte that is not part of a complete record (the
 * next part starts at the block holding it, or earlier: wgbs_bam_first_key).  A part is viewed like a file; combined with the
 * template window of wgbs_view_opts the parts of a file give exactly the templates of the whole file, each once. */
int wgbs_bam_open_part(const void *bgzf, size_t nbytes, int n_ref, const char *const *ref_names, const int32_t *ref_lens,
                       int has_header, uint64_t first_record, int threads, wgbs_bam **out, uint64_t *tail);
/* (refid, 0-based POS) of the last complete record; *refid = -2 when there is none */
int wgbs_bam_last_record(const wgbs_bam *, int *refid, int64_t *pos);
/* inflated offset of the first record of reference refid that passes the filters (key window ignored) and whose template key is
 * >= key; *found = 0 when there is none */
int wgbs_bam_first_key(const wgbs_bam *, const wgbs_view_opts *, int refid, int64_t key, uint64_t *offset, int *found);
uint64_t wgbs_bam_inflated_bytes(const wgbs_bam *);
/* bgzf: a few consecutive whole BGZF blocks from anywhere in a .bam.  *offset = inflated offset (within these blocks) of the first
 * record that starts in them, with its refid / 0-based POS; *found = 0 when no record starts there (probe more blocks).  Lets a
 * multi-GPU bam2pat give every rank its own block range of a streamed file without a .bai (binary search per chromosome). */
int wgbs_bam_probe(const void *bgzf, size_t nbytes, int n_ref, uint64_t *offset, int *refid, int64_t *pos, int *found);


/* ---------------------------------------------------------------------------------------------------------------
 * BAM ingest ON THE DEVICE (the `samtools view` stage of reference bam2pat.py:126-165 "without a SAM-text
 * detour" over PCIe).  The COMPRESSED bytes of a coordinate-sorted .bam file are uploaded (about 55 B per 150 bp read
 * instead of ~360 B of SAM text); one warp per BGZF block inflates it in HBM, the BAM records are located in the inflated
 * stream (guess + verified chain walk: exact for any record / block alignment), and views with the same filters as
 * wgbs_bam_view_ex are formatted to SAM text in HBM, ready for wgbs_pileup_sam.  Same results, byte for byte, as the host
 * reader above.  The whole inflated stream stays resident: a file whose inflated size exceeds free device memory is refused.
 * ------------------------------------------------------------------------------------------------------------- */
/* BGZF inflate on the device (htslib's bgzf reader: deflate + ISIZE + CRC32 checks): bgzf = the bytes of any BGZF file
 * (.bam, .pat.gz, CpG.bed.gz) in HOST memory; *dev_out = DEVICE buffer with the inflated bytes (wgbs_dev_free).  E.g. a
 * .pat.gz goes from disk to wgbs_pats_from_text with only its compressed bytes crossing PCIe (`gunzip -c` of reference
 * pat2beta.py:17-22 / homog.py:66). */
int wgbs_bgzf_inflate(wgbs_ctx *, const void *bgzf, size_t nbytes, void **dev_out, size_t *out_bytes);
typedef struct wgbs_dbam wgbs_dbam;
/* bgzf: the bytes of a whole .bam file in HOST memory (pinned memory makes the upload a single DMA) */
int wgbs_dbam_open(wgbs_ctx *, const void *bgzf, size_t nbytes, wgbs_dbam **out);
int wgbs_dbam_open_file(wgbs_ctx *, const char *path, wgbs_dbam **out);
/* The block table of a BGZF file (what `bgzip -i` keeps in a .gzi, plus each block's CRC32 / ISIZE), built once per file from its
 * bytes in HOST memory.  With it wgbs_dbam_open_indexed takes the compressed bytes from host OR DEVICE memory (a device buffer
 * must be readable for 64 bytes past nbytes): a file that is already resident in HBM is opened without touching the host copy. */
typedef struct wgbs_bgzf_index wgbs_bgzf_index;
int wgbs_bgzf_index_build(const void *bgzf, size_t nbytes, wgbs_bgzf_index **out);
void wgbs_bgzf_index_free(wgbs_bgzf_index *);
uint64_t wgbs_bgzf_index_blocks(const wgbs_bgzf_index *);
int wgbs_dbam_open_indexed(wgbs_ctx *, const void *bgzf, size_t nbytes, const wgbs_bgzf_index *, wgbs_dbam **out);
void wgbs_dbam_close(wgbs_ctx *, wgbs_dbam *);
int wgbs_dbam_nref(const wgbs_dbam *);
const char *wgbs_dbam_ref_name(const wgbs_dbam *, int i);
const char *wgbs_dbam_header(const wgbs_dbam *);
uint64_t wgbs_dbam_nrecords(const wgbs_dbam *, int refid);
uint64_t wgbs_dbam_inflated_bytes(const wgbs_dbam *);
/* parts of a file larger than device memory: wgbs_bam_open_part / wgbs_bam_last_record / wgbs_bam_first_key on the device (the
 * bytes of the part are uploaded, inflated and indexed in HBM; same arguments, same results) */
int wgbs_dbam_open_part(wgbs_ctx *, const void *bgzf, size_t nbytes, int n_ref, const char *const *ref_names, const int32_t *ref_lens,
                        int has_header, uint64_t first_record, wgbs_dbam **out, uint64_t *tail);
int wgbs_dbam_last_record(wgbs_ctx *, const wgbs_dbam *, int *refid, int64_t *pos);
int wgbs_dbam_first_key(wgbs_ctx *, const wgbs_dbam *, const wgbs_view_opts *, int refid, int64_t key, uint64_t *offset, int *found);
/* *dev_text: DEVICE buffer holding the SAM text of the passing records (release with wgbs_dev_free) */
int wgbs_dbam_view(wgbs_ctx *, const wgbs_dbam *, const wgbs_view_opts *, char **dev_text, size_t *nbytes, uint64_t *nrecords);
/* wgbs_dbam_view + wgbs_pileup_sam_mbias without leaving the device (mbias may be NULL) */
int wgbs_pileup_dbam(wgbs_ctx *, const wgbs_index *, const wgbs_dbam *, const wgbs_view_opts *, const wgbs_pileup_opts *,
                     wgbs_pats **out, uint64_t *stats, int32_t *mbias);

#ifdef __cplusplus
}
#endif
#endif /* WGBS_B200_H */
