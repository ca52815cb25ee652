// bam.cu -- host-side BAM ingest: BGZF inflate on a thread pool + BAM records -> the SAM text that
// `samtools view BAM chr -q Q -F X [-f Y]` would print (reference src/python/bam2pat.py:165), which is what the pileup
// front end (wgbs_pileup_sam) consumes.  No CUDA in this file; it lives in the library so that the drop-in covers the
// reference's `samtools view |` stage where samtools is not installed.  Format: SAM/BAM specification v1, sections 4.1-4.2.
#include <zlib.h>
#include <chrono>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "bam_core.cuh"
#include "common.cuh"

// the inflated stream: allocated WITHOUT zero-filling (a std::vector would write every byte once on one thread before the
// inflate threads write it again: 0.3 s per GB, more than the parallel inflate itself)
struct RawBytes {
    uint8_t *p = nullptr; size_t n = 0;
    RawBytes() = default;
    RawBytes(const RawBytes &) = delete;
    RawBytes &operator=(const RawBytes &) = delete;
    ~RawBytes() { free(p); }
    bool resize(size_t k) { free(p); p = (uint8_t *)malloc(k ? k : 1); n = p ? k : 0; return p != nullptr; }
    uint8_t *data() { return p; }
    const uint8_t *data() const { return p; }
    size_t size() const { return n; }
};

struct wgbs_bam {
    RawBytes data;                             // uncompressed BAM stream
    std::string header_text;
    std::vector<std::string> ref_names;
    std::vector<int32_t> ref_lens;
    std::vector<uint64_t> rec_off;             // offset of every record's block_size field
    std::vector<uint64_t> ref_first, ref_last; // record index range [first, last) per refID (coordinate-sorted input)
    int threads = 1;
};

namespace {

inline uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
inline int32_t rdi32(const uint8_t *p) { int32_t v; memcpy(&v, p, 4); return v; }
inline uint16_t rd16(const uint8_t *p) { uint16_t v; memcpy(&v, p, 2); return v; }

struct Block { uint64_t coff; uint32_t csize, usize; uint64_t uoff; };

int inflate_block(const uint8_t *src, uint32_t csize, uint8_t *dst, uint32_t usize) {
    if (usize == 0) return rd32(src + csize - 8) == 0 ? 0 : -1;
    z_stream zs; memset(&zs, 0, sizeof zs);
    if (inflateInit2(&zs, -15) != Z_OK) return -1;
    const uint16_t xlen = rd16(src + 10);
    zs.next_in = const_cast<Bytef *>(src + 12 + xlen); zs.avail_in = csize - 12 - xlen - 8;
    zs.next_out = dst; zs.avail_out = usize;
    int rc = inflate(&zs, Z_FINISH);
    inflateEnd(&zs);
    if (!(rc == Z_STREAM_END && zs.avail_out == 0)) return -1;
    return (uint32_t)crc32(crc32(0L, Z_NULL, 0), dst, usize) == rd32(src + csize - 8) ? 0 : -1;   // gzip trailer: CRC32, ISIZE
}

// ---- formatting: raw pointer writes into a per-thread buffer that is grown per record, no per-char container calls ----
struct OutBuf {
    char *p = nullptr; size_t n = 0, cap = 0;
    ~OutBuf() { free(p); }
    bool reserve(size_t extra) {
        if (n + extra <= cap) return true;
        size_t nc = std::max(cap * 2, n + extra + (1u << 20));
        char *q = (char *)realloc(p, nc);
        if (!q) return false;
        p = q; cap = nc; return true;
    }
};
inline char *put_u64(char *o, unsigned long long v) {
    char t[24]; int k = 0;
    do { t[k++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (k) *o++ = t[--k];
    return o;
}
inline char *put_i64(char *o, long long v) {
    if (v < 0) { *o++ = '-'; return put_u64(o, 0ull - (unsigned long long)v); }
    return put_u64(o, (unsigned long long)v);
}
inline char *put_str(char *o, const char *s, size_t n) { memcpy(o, s, n); return o + n; }
struct SeqTable { char t[256][2]; SeqTable() { const char *a = "=ACMGRSVTWYHKDBN"; for (int i = 0; i < 256; i++) { t[i][0] = a[i >> 4]; t[i][1] = a[i & 15]; } } };
static const SeqTable g_seq;

// one record -> one SAM line (samtools view formatting).  Worst-case growth: SEQ 2x, int8 B-array items "-128," 5x.
bool format_record(const wgbs_bam *B, const uint8_t *r, OutBuf &ob) {
    const uint32_t bs = rd32(r);
    if (!ob.reserve((size_t)bs * 6 + 256)) return false;
    char *o = ob.p + ob.n;
    const uint8_t *p = r + 4, *end = r + 4 + bs;
    const int32_t refid = rdi32(p), pos = rdi32(p + 4);
    const uint8_t l_name = p[8], mapq = p[9];
    const uint16_t n_cig = rd16(p + 12), flag = rd16(p + 14);
    const int32_t l_seq = rdi32(p + 16), nref = rdi32(p + 20), npos = rdi32(p + 24), tlen = rdi32(p + 28);
    const char *name = (const char *)(p + 32);
    const uint8_t *cig = p + 32 + l_name, *seq = cig + 4 * (size_t)n_cig, *qual = seq + (l_seq + 1) / 2, *tags = qual + l_seq;
    const int nrefs = (int)B->ref_names.size();
    o = put_str(o, name, l_name ? strnlen(name, l_name) : 0);       // l_read_name counts the trailing NUL
    *o++ = '\t'; o = put_u64(o, flag); *o++ = '\t';
    if (refid >= 0 && refid < nrefs) o = put_str(o, B->ref_names[refid].data(), B->ref_names[refid].size()); else *o++ = '*';
    *o++ = '\t'; o = put_i64(o, (long long)pos + 1); *o++ = '\t'; o = put_u64(o, mapq); *o++ = '\t';
    if (n_cig == 0) *o++ = '*';
    else for (uint16_t k = 0; k < n_cig; k++) { const uint32_t c = rd32(cig + 4 * k); o = put_u64(o, c >> 4); *o++ = "MIDNSHP=XB??????"[c & 15]; }
    *o++ = '\t';
    if (nref < 0) *o++ = '*'; else if (nref == refid) *o++ = '='; else if (nref < nrefs) o = put_str(o, B->ref_names[nref].data(), B->ref_names[nref].size()); else *o++ = '*';
    *o++ = '\t'; o = put_i64(o, (long long)npos + 1); *o++ = '\t'; o = put_i64(o, tlen); *o++ = '\t';
    if (l_seq <= 0) *o++ = '*';
    else {
        const int32_t full = l_seq >> 1;
        for (int32_t k = 0; k < full; k++) { o[0] = g_seq.t[seq[k]][0]; o[1] = g_seq.t[seq[k]][1]; o += 2; }
        if (l_seq & 1) *o++ = g_seq.t[seq[full]][0];
    }
    *o++ = '\t';
    if (l_seq <= 0 || qual[0] == 0xff) *o++ = '*';
    else { for (int32_t k = 0; k < l_seq; k++) o[k] = (char)(qual[k] + 33); o += l_seq; }
    // optional fields
    const uint8_t *t = tags;
    while (t + 3 <= end) {
        *o++ = '\t'; *o++ = (char)t[0]; *o++ = (char)t[1]; *o++ = ':';
        const char ty = (char)t[2]; t += 3;
        switch (ty) {
            case 'A': *o++ = 'A'; *o++ = ':'; *o++ = (char)*t; t += 1; break;
            case 'c': *o++ = 'i'; *o++ = ':'; o = put_i64(o, (int8_t)*t); t += 1; break;
            case 'C': *o++ = 'i'; *o++ = ':'; o = put_u64(o, *t); t += 1; break;
            case 's': *o++ = 'i'; *o++ = ':'; o = put_i64(o, (int16_t)rd16(t)); t += 2; break;
            case 'S': *o++ = 'i'; *o++ = ':'; o = put_u64(o, rd16(t)); t += 2; break;
            case 'i': *o++ = 'i'; *o++ = ':'; o = put_i64(o, rdi32(t)); t += 4; break;
            case 'I': *o++ = 'i'; *o++ = ':'; o = put_u64(o, rd32(t)); t += 4; break;
            case 'f': { float f; memcpy(&f, t, 4); *o++ = 'f'; *o++ = ':'; o += snprintf(o, 32, "%g", f); t += 4; break; }
            case 'Z': case 'H': { *o++ = ty; *o++ = ':'; const char *z = (const char *)t; size_t l = strnlen(z, end - t); o = put_str(o, z, l); t += l + 1; break; }
            case 'B': {
                const char sub = (char)t[0]; const uint32_t cnt = rd32(t + 1); t += 5;
                *o++ = 'B'; *o++ = ':'; *o++ = sub;
                for (uint32_t k = 0; k < cnt && t < end; k++) {
                    *o++ = ',';
                    switch (sub) {
                        case 'c': o = put_i64(o, (int8_t)*t); t += 1; break;
                        case 'C': o = put_u64(o, *t); t += 1; break;
                        case 's': o = put_i64(o, (int16_t)rd16(t)); t += 2; break;
                        case 'S': o = put_u64(o, rd16(t)); t += 2; break;
                        case 'i': o = put_i64(o, rdi32(t)); t += 4; break;
                        case 'I': o = put_u64(o, rd32(t)); t += 4; break;
                        case 'f': { float f; memcpy(&f, t, 4); o += snprintf(o, 32, "%g", f); t += 4; break; }
                        default: t = end; break;
                    }
                }
                break;
            }
            default: t = end; break;   // unknown type: stop (malformed)
        }
    }
    *o++ = '\n';
    ob.n = (size_t)(o - ob.p);
    return true;
}

}  // namespace

namespace {
// comp[0..fsz): consecutive whole BGZF blocks.  part == false: a whole file (BAM header first, every record complete).
// part == true: a window of a file -- the reference list comes from the caller, records are indexed from inflated offset
// `first_record` on, and the last record may be cut off by the end of the window: *tail = inflated offset of the first byte
// that is not part of a complete record.
int open_stream(const char *who, const uint8_t *comp, uint64_t fsz, int threads, bool part, int n_ref_in, const char *const *ref_names,
                const int32_t *ref_lens, uint64_t first_record, wgbs_bam **out, uint64_t *tail) {
    const bool dbg = getenv("WGBS_BAM_DEBUG") != nullptr; auto T0 = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) { if (dbg) { auto t = std::chrono::steady_clock::now(); fprintf(stderr, "[bam] %-10s %.3f s\n", what, std::chrono::duration<double>(t - T0).count()); T0 = t; } };
    // 1. BGZF block table
    std::vector<Block> blocks; uint64_t off = 0, uoff = 0;
    while (off + 28 <= fsz) {
        const uint8_t *h = comp + off;
        if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) return wgbs_set_err("%s: not a BGZF file (bad block header at %llu)", who, (unsigned long long)off);
        const uint16_t xlen = rd16(h + 10);
        uint32_t bsize = 0; bool found = false;
        for (uint32_t x = 0; x + 4 <= xlen;) {
            const uint8_t *sf = h + 12 + x; const uint16_t sl = rd16(sf + 2);
            if (sf[0] == 'B' && sf[1] == 'C' && sl == 2) { bsize = rd16(sf + 4) + 1u; found = true; break; }
            x += 4 + sl;
        }
        if (!found || off + bsize > fsz) return wgbs_set_err("%s: corrupt BGZF block at %llu", who, (unsigned long long)off);
        Block b; b.coff = off; b.csize = bsize; b.usize = rd32(h + bsize - 4); b.uoff = uoff;
        blocks.push_back(b); off += bsize; uoff += b.usize;
    }
    if (part && off != fsz) return wgbs_set_err("%s: a part must consist of whole BGZF blocks (%llu trailing bytes)", who, (unsigned long long)(fsz - off));
    wgbs_bam *B = new wgbs_bam();
    B->threads = threads > 0 ? threads : (int)std::max(1u, std::thread::hardware_concurrency());
    lap("blocks");
    if (!B->data.resize(uoff)) { delete B; return wgbs_set_err("%s: cannot allocate %llu bytes for the inflated stream", who, (unsigned long long)uoff); }
    lap("alloc");
    // 2. inflate in parallel
    std::atomic<size_t> next(0); std::atomic<int> bad(0);
    auto work = [&]() {
        for (size_t i; (i = next.fetch_add(1)) < blocks.size();)
            if (inflate_block(comp + blocks[i].coff, blocks[i].csize, B->data.data() + blocks[i].uoff, blocks[i].usize)) bad.store(1);
    };
    {
        std::vector<std::thread> th; int nt = std::min<int>(B->threads, (int)std::max<size_t>(1, blocks.size() / 4));
        for (int t = 1; t < nt; t++) th.emplace_back(work);
        work();
        for (auto &t : th) t.join();
    }
    lap("inflate");
    if (bad.load()) { delete B; return wgbs_set_err("%s: inflate failed (corrupt BGZF block or CRC32 mismatch)", who); }
    // 3. header
    const uint8_t *d = B->data.data(); const uint64_t n = B->data.size();
    uint64_t p = 0; uint32_t n_ref = 0;
    if (!part) {
        if (n < 12 || memcmp(d, "BAM\1", 4)) { delete B; return wgbs_set_err("%s: not a BAM file", who); }
        const uint32_t l_text = rd32(d + 4);
        if (8ull + l_text + 4 > n) { delete B; return wgbs_set_err("%s: truncated BAM header", who); }
        B->header_text.assign((const char *)d + 8, strnlen((const char *)d + 8, l_text));
        p = 8ull + l_text; n_ref = rd32(d + p); p += 4;
        for (uint32_t i = 0; i < n_ref; i++) {
            if (p + 4 > n) { delete B; return wgbs_set_err("%s: truncated reference list", who); }
            const uint32_t l = rd32(d + p); p += 4;
            if (p + l + 4 > n) { delete B; return wgbs_set_err("%s: truncated reference list", who); }
            B->ref_names.emplace_back((const char *)d + p, l ? l - 1 : 0); p += l;
            B->ref_lens.push_back(rdi32(d + p)); p += 4;
        }
    } else {
        n_ref = (uint32_t)n_ref_in;
        for (uint32_t i = 0; i < n_ref; i++) { B->ref_names.emplace_back(ref_names[i]); B->ref_lens.push_back(ref_lens ? ref_lens[i] : 0); }
        p = first_record;
        if (p > n) { delete B; return wgbs_set_err("%s: first record offset %llu beyond the part (%llu inflated bytes)", who, (unsigned long long)p, (unsigned long long)n); }
    }
    // 4. record table (records are length-prefixed: a sequential walk) + per-reference ranges
    B->ref_first.assign(n_ref + 1, 0); B->ref_last.assign(n_ref + 1, 0);
    int32_t cur = -2; uint64_t idx = 0;
    while (p + 4 <= n) {
        const uint32_t bs = rd32(d + p);
        if (part && bs >= 32 && p + 4 + bs > n) break;                     // cut off by the end of the window: the next part starts here
        if (bs < 32 || p + 4 + bs > n) { delete B; return wgbs_set_err("%s: corrupt BAM record at uncompressed offset %llu", who, (unsigned long long)p); }
        const int32_t refid = rdi32(d + p + 4);
        const uint32_t slot = (refid >= 0 && (uint32_t)refid < n_ref) ? (uint32_t)refid : n_ref;
        if (refid != cur) {
            if (B->ref_last[slot] != 0) { delete B; return wgbs_set_err("%s is not sorted by coordinate (reference %d appears in two separate runs)", who, refid); }
            B->ref_first[slot] = idx; cur = refid;
        }
        B->ref_last[slot] = idx + 1;
        B->rec_off.push_back(p); p += 4 + bs; idx++;
    }
    lap("walk");
    if (tail) *tail = p;
    *out = B;
    return 0;
}
}  // namespace

extern "C" int wgbs_bam_open(const char *path, int threads, wgbs_bam **out) {
    if (!path || !out) return wgbs_set_err("wgbs_bam_open: null argument");
    *out = nullptr;
    FILE *f = fopen(path, "rb");
    if (!f) return wgbs_set_err("wgbs_bam_open: cannot open %s", path);
    fseek(f, 0, SEEK_END); const long fsz = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<uint8_t> comp((size_t)fsz);
    if (fsz && fread(comp.data(), 1, (size_t)fsz, f) != (size_t)fsz) { fclose(f); return wgbs_set_err("wgbs_bam_open: short read on %s", path); }
    fclose(f);
    return open_stream(path, comp.data(), (uint64_t)fsz, threads, false, 0, nullptr, nullptr, 0, out, nullptr);
}

// A window of a .bam that does not fit in memory as a whole (bam2pat streams such files): see wgbs_b200.h
extern "C" int wgbs_bam_open_part(const void *bgzf, size_t nbytes, int n_ref, const char *const *ref_names, const int32_t *ref_lens,
                                  int has_header, uint64_t first_record, int threads, wgbs_bam **out, uint64_t *tail) {
    if (!bgzf || !out || !tail || (!has_header && n_ref > 0 && !ref_names)) return wgbs_set_err("wgbs_bam_open_part: null argument");
    *out = nullptr;
    if (has_header) {
        // the first window of a file: header, then records; its last record may be cut off like any other window's
        wgbs_bam *H = nullptr;
        // parse the header with the whole-file walker disabled: open as a part at the offset right behind the reference list
        // (found by a header-only pass over the first inflated bytes)
        std::vector<const char *> names; std::vector<int32_t> lens; std::vector<std::string> keep; uint64_t first = 0; std::string text;
        {
            // inflate just enough leading blocks to hold the header
            const uint8_t *c = (const uint8_t *)bgzf; uint64_t off = 0; std::vector<uint8_t> head;
            auto need = [&](uint64_t upto) -> bool {
                while (head.size() < upto && off + 28 <= nbytes) {
                    const uint8_t *h = c + off;
                    if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) return false;
                    const uint16_t xlen = rd16(h + 10); uint32_t bsize = 0;
                    for (uint32_t x = 0; x + 4 <= xlen;) { const uint8_t *sf = h + 12 + x; const uint16_t sl = rd16(sf + 2); if (sf[0] == 'B' && sf[1] == 'C' && sl == 2) { bsize = rd16(sf + 4) + 1u; break; } x += 4 + sl; }
                    if (!bsize || off + bsize > nbytes) return false;
                    const uint32_t usize = rd32(h + bsize - 4); const size_t at = head.size();
                    head.resize(at + usize);
                    if (inflate_block(h, bsize, head.data() + at, usize)) return false;
                    off += bsize;
                }
                return head.size() >= upto;
            };
            if (!need(12) || memcmp(head.data(), "BAM\1", 4)) return wgbs_set_err("wgbs_bam_open_part: not a BAM file");
            const uint32_t l_text = rd32(head.data() + 4);
            if (!need(8ull + l_text + 4)) return wgbs_set_err("wgbs_bam_open_part: truncated BAM header");
            text.assign((const char *)head.data() + 8, strnlen((const char *)head.data() + 8, l_text));
            uint64_t p = 8ull + l_text; const uint32_t nr = rd32(head.data() + p); p += 4;
            for (uint32_t i = 0; i < nr; i++) {
                if (!need(p + 4)) return wgbs_set_err("wgbs_bam_open_part: truncated reference list");
                const uint32_t l = rd32(head.data() + p); p += 4;
                if (!need(p + l + 4)) return wgbs_set_err("wgbs_bam_open_part: truncated reference list");
                keep.emplace_back((const char *)head.data() + p, l ? strnlen((const char *)head.data() + p, l - 1) : 0); p += l;
                lens.push_back(rdi32(head.data() + p)); p += 4;
            }
            first = p;
        }
        for (auto &k : keep) names.push_back(k.c_str());
        int rc = open_stream("wgbs_bam_open_part", (const uint8_t *)bgzf, nbytes, threads, true, (int)names.size(), names.data(), lens.data(), first, &H, tail);
        if (rc < 0) return rc;
        H->header_text = text;
        *out = H;
        return 0;
    }
    return open_stream("wgbs_bam_open_part", (const uint8_t *)bgzf, nbytes, threads, true, n_ref, ref_names, ref_lens, first_record, out, tail);
}

// Where does a record start in the middle of a file?  bgzf: a few consecutive whole BGZF blocks from anywhere in a .bam.  Finds
// the first offset of their inflated bytes from which the length-prefixed record chain runs plausibly (bam_core.cuh
// plausible_at) all the way to the end of the probed bytes -- hundreds of records when the probe spans a few blocks, so a
// chance hit inside sequence or quality bytes cannot survive -- and returns it with that record's refid / POS.  *found = 0:
// no record starts inside the probed bytes (take more blocks).  bam2pat uses it to give every rank of a multi-GPU run its own
// block range of a streamed file (binary search for the first block of each chromosome) without a .bai index.
extern "C" int wgbs_bam_probe(const void *bgzf, size_t nbytes, int n_ref, uint64_t *offset, int *refid, int64_t *pos, int *found) {
    if (!bgzf || !offset || !refid || !pos || !found) return wgbs_set_err("wgbs_bam_probe: null argument");
    *found = 0; *offset = 0; *refid = -2; *pos = -1;
    const uint8_t *c = (const uint8_t *)bgzf; uint64_t off = 0; std::vector<uint8_t> d;
    while (off + 28 <= nbytes) {
        const uint8_t *h = c + off;
        if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) return wgbs_set_err("wgbs_bam_probe: not a BGZF block at %llu", (unsigned long long)off);
        const uint16_t xlen = rd16(h + 10); uint32_t bsize = 0;
        for (uint32_t x = 0; x + 4 <= xlen;) { const uint8_t *sf = h + 12 + x; const uint16_t sl = rd16(sf + 2); if (sf[0] == 'B' && sf[1] == 'C' && sl == 2) { bsize = rd16(sf + 4) + 1u; break; } x += 4 + sl; }
        if (!bsize || off + bsize > nbytes) return wgbs_set_err("wgbs_bam_probe: corrupt BGZF block at %llu", (unsigned long long)off);
        const uint32_t usize = rd32(h + bsize - 4); const size_t at = d.size();
        d.resize(at + usize);
        if (inflate_block(h, bsize, d.data() + at, usize)) return wgbs_set_err("wgbs_bam_probe: inflate failed");
        off += bsize;
    }
    const uint64_t n = d.size();
    for (uint64_t o = 0; o + 36 <= n; o++) {
        if (!bamcore::plausible_at(d.data(), n, o, n_ref)) continue;
        uint64_t q = o; int chain = 0; bool ok = true;
        while (q + 4 <= n) {                                             // the whole chain to the end of the probe must hold
            const uint32_t bs = rd32(d.data() + q);
            if (bs >= 32 && q + 4 + (uint64_t)bs > n) break;             // cut off by the end of the probe: fine
            if (!bamcore::plausible_at(d.data(), n, q, n_ref)) { ok = false; break; }
            q += 4 + (uint64_t)bs; chain++;
        }
        if (ok && chain >= 2) { *found = 1; *offset = o; *refid = rdi32(d.data() + o + 4); *pos = rdi32(d.data() + o + 8); return 0; }
    }
    return 0;
}

// (refid, 0-based POS) of the last complete record of a file / part; *refid = -2 when there is no record
extern "C" int wgbs_bam_last_record(const wgbs_bam *B, int *refid, int64_t *pos) {
    if (!B || !refid || !pos) return wgbs_set_err("wgbs_bam_last_record: null argument");
    *refid = -2; *pos = -1;
    if (!B->rec_off.empty()) { const uint8_t *r = B->data.data() + B->rec_off.back(); *refid = rdi32(r + 4); *pos = rdi32(r + 8); }
    return 0;
}
extern "C" uint64_t wgbs_bam_inflated_bytes(const wgbs_bam *B) { return B ? B->data.size() : 0; }

extern "C" void wgbs_bam_close(wgbs_bam *B) { delete B; }
extern "C" int wgbs_bam_nref(const wgbs_bam *B) { return B ? (int)B->ref_names.size() : -1; }
extern "C" const char *wgbs_bam_ref_name(const wgbs_bam *B, int i) { return (B && i >= 0 && i < (int)B->ref_names.size()) ? B->ref_names[i].c_str() : nullptr; }
extern "C" const char *wgbs_bam_header(const wgbs_bam *B) { return B ? B->header_text.c_str() : nullptr; }
extern "C" uint64_t wgbs_bam_nrecords(const wgbs_bam *B, int refid) {
    if (!B) return 0;
    if (refid < 0) return B->rec_off.size();
    return refid < (int)B->ref_names.size() ? B->ref_last[refid] - B->ref_first[refid] : 0;
}

namespace {
// value of a Z-typed tag (e.g. "RG"), or nullptr
const char *find_z_tag(const uint8_t *t, const uint8_t *end, char a, char b, size_t *len) {
    while (t + 3 <= end) {
        const char ty = (char)t[2]; const bool hit = (char)t[0] == a && (char)t[1] == b; t += 3;
        switch (ty) {
            case 'A': case 'c': case 'C': t += 1; break;
            case 's': case 'S': t += 2; break;
            case 'i': case 'I': case 'f': t += 4; break;
            case 'Z': case 'H': { const char *z = (const char *)t; size_t l = strnlen(z, end - t); if (hit && ty == 'Z') { *len = l; return z; } t += l + 1; break; }
            case 'B': {
                if (t + 5 > end) return nullptr;
                const char sub = (char)t[0]; const uint32_t cnt = rd32(t + 1); t += 5;
                const size_t w = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : (sub == 'i' || sub == 'I' || sub == 'f') ? 4 : 0;
                if (!w) return nullptr;
                t += (size_t)cnt * w; break;
            }
            default: return nullptr;
        }
    }
    return nullptr;
}
}  // namespace

// the `samtools view` filters of wgbs_view_opts on one record (r -> block_size)
static bool rec_passes(const uint8_t *r, const wgbs_view_opts *vo, size_t rg_len) {
    const int min_mapq = vo->min_mapq, exclude_flags = vo->exclude_flags, include_flags = vo->include_flags;
    const int64_t beg = vo->beg, end = vo->end;
    const bool need_span = end > 0 || vo->n_iv;
    const uint16_t flag = rd16(r + 4 + 14); const uint8_t mapq = r[4 + 9];
    if (mapq < min_mapq || (flag & exclude_flags) || (include_flags && (flag & include_flags) != include_flags)) return false;
    if (vo->n_flag_eq) {                                     // awk '($2 == A || $2 == B)' (bam2pat.py:135-144)
        bool ok = false;
        for (int k = 0; k < vo->n_flag_eq; k++) ok |= (int)flag == vo->flag_eq[k];
        if (!ok) return false;
    }
    if (vo->key_end > 0) {                                   // template window: max(POS, PNEXT) of a pair on one reference (wgbs_b200.h)
        const int32_t rid = rdi32(r + 4), pos0 = rdi32(r + 8), nref = rdi32(r + 24), npos = rdi32(r + 28);
        const int64_t key = ((flag & 1) && !(flag & 8) && nref == rid && npos > pos0) ? npos : pos0;
        if (key < vo->key_beg || key >= vo->key_end) return false;
    }
    const uint16_t n_cig = rd16(r + 4 + 12); const uint8_t l_name = r[4 + 8];
    if (need_span) {
        const int64_t pos0 = (int64_t)rdi32(r + 8);          // 0-based
        int64_t span = 0; const uint8_t *cig = r + 36 + l_name;
        for (uint16_t k = 0; k < n_cig; k++) { uint32_t c = rd32(cig + 4 * k); uint32_t op = c & 15; if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) span += c >> 4; }
        if (span < 1) span = 1;
        if (end > 0 && (pos0 + 1 > end || pos0 + span < beg)) return false;
        if (vo->n_iv) {                                      // [pos0, pos0+span) against sorted disjoint [iv_beg, iv_end)
            const int64_t *e = std::upper_bound(vo->iv_end, vo->iv_end + vo->n_iv, pos0);     // first interval ending after pos0
            const bool hit = e != vo->iv_end + vo->n_iv && vo->iv_beg[e - vo->iv_end] < pos0 + span;
            if (hit == (vo->iv_exclude != 0)) return false;
        }
    }
    if (rg_len || vo->read_group) {                          // samtools view -r RG
        const uint32_t bs = rd32(r); const int32_t l_seq = rdi32(r + 4 + 16);
        const uint8_t *tags = r + 36 + l_name + 4 * (size_t)n_cig + (size_t)((l_seq + 1) / 2) + (size_t)(l_seq > 0 ? l_seq : 0);
        size_t zl = 0; const char *z = find_z_tag(tags, r + 4 + bs, 'R', 'G', &zl);
        if (!z || zl != rg_len || memcmp(z, vo->read_group, zl)) return false;
    }
    return true;
}

// SAM text of the records that pass the filters of `o` (see wgbs_view_opts in the header).  *text is malloc'ed: release
// with wgbs_host_free.
extern "C" int wgbs_bam_view_ex(const wgbs_bam *B, const wgbs_view_opts *vo, char **text, size_t *nbytes, uint64_t *nrecords) {
    if (!B || !vo || !text || !nbytes) return wgbs_set_err("wgbs_bam_view: null argument");
    if (vo->n_flag_eq < 0 || vo->n_flag_eq > 4) return wgbs_set_err("wgbs_bam_view: n_flag_eq must be 0..4");
    if (vo->n_iv && (!vo->iv_beg || !vo->iv_end)) return wgbs_set_err("wgbs_bam_view: interval list is null");
    for (size_t k = 0; k + 1 < vo->n_iv; k++)
        if (vo->iv_beg[k + 1] < vo->iv_end[k] || vo->iv_end[k] < vo->iv_beg[k]) return wgbs_set_err("wgbs_bam_view: intervals must be sorted and non-overlapping");
    const int refid = vo->refid;
    const size_t rg_len = vo->read_group ? strlen(vo->read_group) : 0;
    uint64_t r0 = 0, r1 = B->rec_off.size();
    if (refid >= 0) { if (refid >= (int)B->ref_names.size()) return wgbs_set_err("wgbs_bam_view: no such reference"); r0 = B->ref_first[refid]; r1 = B->ref_last[refid]; }
    // head -N: a sequential walk (the caller wants the FIRST max_records passing records)
    const int nt = vo->max_records ? 1 : (int)std::max<uint64_t>(1, std::min<uint64_t>(B->threads, (r1 - r0) / 2048 + 1));
    std::vector<OutBuf> parts(nt); std::vector<uint64_t> cnt(nt, 0); std::atomic<int> oom(0);
    auto work = [&](int t) {
        const uint64_t a = r0 + (r1 - r0) * t / nt, b = r0 + (r1 - r0) * (t + 1) / nt;
        OutBuf &o = parts[t];
        if (!o.reserve((vo->max_records ? std::min<uint64_t>(b - a, vo->max_records) : (b - a)) * 384 + 4096)) { oom.store(1); return; }
        for (uint64_t i = a; i < b; i++) {
            const uint8_t *r = B->data.data() + B->rec_off[i];
            if (!rec_passes(r, vo, rg_len)) continue;
            if (!format_record(B, r, o)) { oom.store(1); return; }
            cnt[t]++;
            if (vo->max_records && cnt[t] >= vo->max_records) break;
        }
    };
    {
        std::vector<std::thread> th;
        for (int t = 1; t < nt; t++) th.emplace_back(work, t);
        work(0);
        for (auto &t : th) t.join();
    }
    if (oom.load()) return wgbs_set_err("wgbs_bam_view: out of memory");
    size_t tot = 0; uint64_t nr = 0; std::vector<size_t> at(nt);
    for (int t = 0; t < nt; t++) { at[t] = tot; tot += parts[t].n; nr += cnt[t]; }
    char *buf = (char *)malloc(tot ? tot : 1);
    if (!buf) return wgbs_set_err("wgbs_bam_view: out of memory");
    {
        auto cp = [&](int t) { if (parts[t].n) memcpy(buf + at[t], parts[t].p, parts[t].n); };
        std::vector<std::thread> th;
        for (int t = 1; t < nt; t++) th.emplace_back(cp, t);
        cp(0);
        for (auto &t : th) t.join();
    }
    *text = buf; *nbytes = tot; if (nrecords) *nrecords = nr;
    return 0;
}

// Inflated offset (in this file / part) of the first record of reference `refid` that passes the filters of `vo` (its key
// window ignored) and whose template key is >= key; *found = 0 when there is none.  bam2pat's streaming reader restarts the
// next window of the file there: everything a later template window needs lies at or behind that record.
extern "C" int wgbs_bam_first_key(const wgbs_bam *B, const wgbs_view_opts *vo, int refid, int64_t key, uint64_t *offset, int *found) {
    if (!B || !vo || !offset || !found) return wgbs_set_err("wgbs_bam_first_key: null argument");
    *found = 0; *offset = 0;
    if (refid < 0 || refid >= (int)B->ref_names.size()) return 0;
    wgbs_view_opts v = *vo; v.key_beg = 0; v.key_end = 0; v.refid = refid;
    const size_t rg_len = v.read_group ? strlen(v.read_group) : 0;
    for (uint64_t i = B->ref_first[refid]; i < B->ref_last[refid]; i++) {
        const uint8_t *r = B->data.data() + B->rec_off[i];
        const uint16_t flag = rd16(r + 4 + 14);
        const int32_t pos0 = rdi32(r + 8), nref = rdi32(r + 24), npos = rdi32(r + 28);
        const int64_t k = ((flag & 1) && !(flag & 8) && nref == refid && npos > pos0) ? npos : pos0;
        if (k >= key && rec_passes(r, &v, rg_len)) { *offset = B->rec_off[i]; *found = 1; return 0; }
    }
    return 0;
}

// SAM text of the records of reference `refid` (-1: every record) that pass `-q min_mapq -F exclude -f include`
// [and overlap the 1-based closed interval beg..end when end > 0].
extern "C" int wgbs_bam_view(const wgbs_bam *B, int refid, int min_mapq, int exclude_flags, int include_flags, int64_t beg, int64_t end,
                             char **text, size_t *nbytes, uint64_t *nrecords) {
    wgbs_view_opts vo; memset(&vo, 0, sizeof vo);
    vo.refid = refid; vo.min_mapq = min_mapq; vo.exclude_flags = exclude_flags; vo.include_flags = include_flags; vo.beg = beg; vo.end = end;
    return wgbs_bam_view_ex(B, &vo, text, nbytes, nrecords);
}

extern "C" void wgbs_host_free(void *p) { free(p); }
