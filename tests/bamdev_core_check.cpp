// TEST INFRASTRUCTURE: builds wgbs_tools_b200/csrc/{inflate_core,bam_core}.cuh -- the per-block / per-record logic the
// kernels of bamdev.cu run -- as plain host C++ (g++ -std=c++20), so tests/test_bamdev_core.py can pin it against zlib,
// printf and the original SAM text without a GPU.  Not part of libwgbs_b200.so.
//
//   bamdev_core_check inflate FILE.bgzf [emu_blocks]   every BGZF block through Inflater<OneLane> vs zlib; the first
//                                                      emu_blocks blocks also through a 32-lane lock-step emulation
//   bamdev_core_check view FILE.bam SEG DEPTH [REFID MAPQ EXCL INCL BEG END FLAGEQ RG IVFILE IVEXCL MAXREC [KEYBEG KEYEND]]
//                                                      inflate, find the records with the segment guess / walk / repair
//                                                      scheme (segment size SEG bytes), apply the `samtools view` filters
//                                                      (FLAGEQ: comma list or '-', RG: read group or '-', IVFILE: "beg end"
//                                                      lines or '-'), print the SAM text; stderr: stats
//   bamdev_core_check inflate3 FILE.bgzf               every BGZF block through the team decoder (inflate3_core.cuh: teams of 32, 16, 8
//                                                      lanes in lock step, and one lane) + token replay vs zlib; stderr: chain statistics
//   bamdev_core_check nptags FILE.bam                  the MM:Z string and the ML:B:C values of every record as the direct route finds them
//   bamdev_core_check tables N SEED                    two-level Huffman tables of the two-phase decoder vs a canonical-code walk
//   bamdev_core_check fmtg N SEED                      fmt_g vs snprintf("%g") on N random floats + edge cases
//   bamdev_core_check part FILE.bam SEG DEPTH B0 NB FIRST REFS   blocks [B0, B0+NB) as a part of a streamed file (dbam_open_impl, part mode)
#include <zlib.h>

#include <algorithm>
#include <barrier>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "../wgbs_tools_b200/csrc/bam_core.cuh"
#include "../wgbs_tools_b200/csrc/inflate2_core.cuh"
#include "../wgbs_tools_b200/csrc/inflate3_core.cuh"

using namespace dflate;

// ---- NL lanes in lock step: one std::thread per lane, every collective is a barrier ------------------------------------
// NL = 32: a warp (dflate::WarpLanes); NL = 4, 8, 16: a team of a warp's lanes (dflate::SubWarp<NL>)
template <int NL>
struct EmuSharedN {
    std::barrier<> bar{NL};
    uint32_t slot[NL];
};
template <int NL>
struct EmuLanesN {
    static constexpr int N = NL;
    int lane; EmuSharedN<NL> *sh;
    int id() const { return lane; }
    void sync() const { sh->bar.arrive_and_wait(); }
    uint32_t shfl(uint32_t v, int src) const { sh->slot[lane] = v; sh->bar.arrive_and_wait(); uint32_t r = sh->slot[src]; sh->bar.arrive_and_wait(); return r; }
    uint32_t ballot(bool p) const {
        sh->slot[lane] = p ? 1u : 0u; sh->bar.arrive_and_wait();
        uint32_t m = 0; for (int i = 0; i < NL; i++) m |= sh->slot[i] << i;
        sh->bar.arrive_and_wait(); return m;
    }
    uint32_t exscan(uint32_t v, uint32_t *total, uint32_t) const {
        sh->slot[lane] = v; sh->bar.arrive_and_wait();
        uint32_t pre = 0, tot = 0; for (int i = 0; i < NL; i++) { if (i < lane) pre += sh->slot[i]; tot += sh->slot[i]; }
        sh->bar.arrive_and_wait(); *total = tot; return pre;
    }
};
using EmuShared = EmuSharedN<32>;
using EmuLanes = EmuLanesN<32>;

static std::vector<uint8_t> slurp(const char *path) {
    FILE *f = fopen(path, "rb"); if (!f) { perror(path); exit(2); }
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<uint8_t> v((size_t)n + 16, 0);
    if (n && fread(v.data(), 1, (size_t)n, f) != (size_t)n) { perror("read"); exit(2); }
    fclose(f); v.resize((size_t)n);
    return v;
}
struct Blk { uint64_t coff; uint32_t csize, xlen, usize; uint64_t uoff; };
static std::vector<Blk> scan_blocks(const std::vector<uint8_t> &f, uint64_t *utotal) {
    std::vector<Blk> b; uint64_t off = 0, uoff = 0;
    while (off + 28 <= f.size()) {
        uint32_t xlen = 0; const uint32_t bs = bgzf_block_size(f.data() + off, f.size() - off, &xlen);
        if (!bs || off + bs > f.size()) { fprintf(stderr, "bad BGZF block at %llu\n", (unsigned long long)off); exit(3); }
        Blk k; k.coff = off; k.csize = bs; k.xlen = xlen; k.usize = bamcore::ld32(f.data() + off + bs - 4); k.uoff = uoff;
        b.push_back(k); off += bs; uoff += k.usize;
    }
    *utotal = uoff;
    return b;
}
static int zlib_inflate(const uint8_t *src, uint32_t n, uint8_t *dst, uint32_t usize) {
    z_stream zs; memset(&zs, 0, sizeof zs);
    if (inflateInit2(&zs, -15) != Z_OK) return -1;
    zs.next_in = const_cast<Bytef *>(src); zs.avail_in = n; zs.next_out = dst; zs.avail_out = usize;
    int rc = usize ? inflate(&zs, Z_FINISH) : Z_STREAM_END;
    inflateEnd(&zs);
    return (rc == Z_STREAM_END && zs.avail_out == 0) ? 0 : -1;
}
static uint32_t g_crc_table[256], g_crc4[1024];
static int one_lane(const uint8_t *src, uint32_t n, uint8_t *dst, uint32_t usize, uint32_t want_crc = 0, bool check_crc = false) {
    Scratch S; Inflater<OneLane> I; I.S = &S; I.dst = dst; I.dst_len = usize;
    int rc = I.run(src, n);
    if (rc == OK && check_crc && crc32_block(OneLane(), dst, usize, g_crc_table) != want_crc) rc = E_CRC;
    return rc;
}
static int one_lane2(const uint8_t *src, uint32_t n, uint8_t *dst, uint32_t usize, uint32_t want_crc) {
    Scratch S; Ring R; Inflater2<OneLane> I; I.S = &S; I.R = &R; I.dst = dst; I.dst_len = usize;
    int rc = I.run(src, n);
    if (rc == OK && crc32_block(OneLane(), dst, usize, g_crc_table) != want_crc) rc = E_CRC;
    return rc;
}
static int emu_warp(const uint8_t *src, uint32_t n, uint8_t *dst, uint32_t usize, uint32_t want_crc) {
    static EmuShared sh; Scratch S; int rcs[32];
    std::vector<std::thread> th;
    for (int l = 0; l < 32; l++) th.emplace_back([&, l]() {
        Inflater<EmuLanes> I; I.lanes = EmuLanes{l, &sh}; I.S = &S; I.dst = dst; I.dst_len = usize; rcs[l] = I.run(src, n);
        if (rcs[l] == OK) { I.lanes.sync(); if (crc32_block(I.lanes, dst, usize, g_crc_table) != want_crc) rcs[l] = E_CRC; }
    });
    for (auto &t : th) t.join();
    for (int l = 1; l < 32; l++) if (rcs[l] != rcs[0]) { fprintf(stderr, "lanes disagree on rc\n"); exit(4); }
    return rcs[0];
}
template <int NL>
static int emu_team2(const uint8_t *src, uint32_t n, uint8_t *dst, uint32_t usize, uint32_t want_crc) {
    static EmuSharedN<NL> sh; Scratch S; Ring R; int rcs[NL];
    std::vector<std::thread> th;
    for (int l = 0; l < NL; l++) th.emplace_back([&, l]() {
        Inflater2<EmuLanesN<NL>> I; I.lanes = EmuLanesN<NL>{l, &sh}; I.S = &S; I.R = &R; I.dst = dst; I.dst_len = usize; rcs[l] = I.run(src, n);
        if (rcs[l] == OK) { I.lanes.sync(); if (crc32_block(I.lanes, dst, usize, g_crc_table) != want_crc) rcs[l] = E_CRC; }
    });
    for (auto &t : th) t.join();
    for (int l = 1; l < NL; l++) if (rcs[l] != rcs[0]) { fprintf(stderr, "lanes disagree on rc (v2, %d lanes)\n", NL); exit(4); }
    return rcs[0];
}
static int emu_warp2(const uint8_t *src, uint32_t n, uint8_t *dst, uint32_t usize, uint32_t want_crc) { return emu_team2<32>(src, n, dst, usize, want_crc); }


static size_t g_fallbacks = 0;

static int cmd_inflate(const char *path, int emu_blocks) {
    auto f = slurp(path); uint64_t ut; auto blocks = scan_blocks(f, &ut);
    size_t nbad = 0, nemu = 0; uint64_t bytes = 0;
    for (size_t i = 0; i < blocks.size(); i++) {
        const Blk &b = blocks[i];
        const uint8_t *src = f.data() + b.coff + 12 + b.xlen; const uint32_t n = b.csize - 12 - b.xlen - 8;
        std::vector<uint8_t> a(b.usize + 1), c(b.usize + 1), e(b.usize + 1);
        const uint32_t want = bamcore::ld32(f.data() + b.coff + b.csize - 8);
        int rz = zlib_inflate(src, n, a.data(), b.usize);
        if (rz == 0 && (uint32_t)crc32(crc32(0L, Z_NULL, 0), a.data(), b.usize) != want) rz = -1;       // what htslib's bgzf reader checks
        const int r1 = one_lane(src, n, c.data(), b.usize, want, true);
        bool ok = (rz == 0) == (r1 == 0) && (rz != 0 || !memcmp(a.data(), c.data(), b.usize));
        { std::vector<uint8_t> c2(b.usize + 1); const int r3 = one_lane2(src, n, c2.data(), b.usize, want); ok = ok && (rz == 0) == (r3 == 0) && (rz != 0 || !memcmp(a.data(), c2.data(), b.usize)); }
        if ((int)i < emu_blocks) {
            const int r2 = emu_warp(src, n, e.data(), b.usize, want); ok = ok && r2 == r1 && (r1 != 0 || !memcmp(a.data(), e.data(), b.usize)); nemu++;
            std::vector<uint8_t> e2(b.usize + 1); const int r4 = emu_warp2(src, n, e2.data(), b.usize, want); ok = ok && (rz == 0) == (r4 == 0) && (rz != 0 || !memcmp(a.data(), e2.data(), b.usize));
            // the team decoders (bgzf_inflate_team_k<G>): the same Inflater2 with batches of 4 / 8 / 16 symbols
            for (int g : {4, 8, 16}) {
                std::vector<uint8_t> eg(b.usize + 1);
                const int rg = g == 4 ? emu_team2<4>(src, n, eg.data(), b.usize, want) : g == 8 ? emu_team2<8>(src, n, eg.data(), b.usize, want) : emu_team2<16>(src, n, eg.data(), b.usize, want);
                const bool okg = rg == r4 && (rz != 0 || !memcmp(a.data(), eg.data(), b.usize));
                if (!okg) fprintf(stderr, "block %zu: team of %d lanes: rc %d (warp %d)\n", i, g, rg, r4);
                ok = ok && okg;
            }
        }
        if (!ok) { nbad++; fprintf(stderr, "block %zu: zlib %d core %d\n", i, rz, r1); }
        bytes += b.usize;
    }
    printf("blocks %zu emu %zu bytes %llu mismatches %zu\n", blocks.size(), nemu, (unsigned long long)bytes, nbad);
    return nbad ? 1 : 0;
}

// ---- the team decoder (inflate3_core.cuh): NL lanes in lock step walk one block, then the token replay of inflate2_core.cuh ----------
template <int NL>
static int emu_team3(const uint8_t *src, uint32_t n, uint8_t *dst, uint32_t usize, uint32_t want_crc, dflate3::TeamStats *stats, bool emu_replay) {
    static EmuSharedN<NL> sh; static dflate3::TeamMem T;
    std::vector<dflate2::Token> tok(dflate2::token_cap(usize));
    int rcs[NL]; uint32_t nts[NL];
    std::vector<std::thread> th;
    for (int l = 0; l < NL; l++) th.emplace_back([&, l]() { rcs[l] = dflate3::team_inflate(EmuLanesN<NL>{l, &sh}, &T, src, n, dst, usize, tok.data(), &nts[l], stats); });
    for (auto &t : th) t.join();
    for (int l = 1; l < NL; l++) if (rcs[l] != rcs[0] || nts[l] != nts[0]) { fprintf(stderr, "lanes disagree (team decoder, %d lanes)\n", NL); exit(4); }
    int rc = rcs[0];
    if (rc == dflate2::E_FALLBACK) { g_fallbacks++; return one_lane2(src, n, dst, usize, want_crc); }
    if (rc == OK && !emu_replay) rc = dflate2::resolve_bytes(OneLane(), tok.data(), nts[0], dst, usize, src);
    else if (rc == OK) {                                             // the byte-per-lane replay under the same lock-step emulation (slow: a sample)
        const uint32_t nt = nts[0];
        std::vector<std::thread> t2;
        for (int l = 0; l < NL; l++) t2.emplace_back([&, l]() { rcs[l] = dflate2::resolve_bytes(EmuLanesN<NL>{l, &sh}, tok.data(), nt, dst, usize, src); });
        for (auto &t : t2) t.join();
        rc = rcs[0];
    }
    if (rc == OK && dflate2::crc32_block4(OneLane(), dst, usize, g_crc4) != want_crc) rc = E_CRC;
    return rc;
}
static int one_team3(const uint8_t *src, uint32_t n, uint8_t *dst, uint32_t usize, uint32_t want_crc) {
    static dflate3::TeamMem T;
    std::vector<dflate2::Token> tok(dflate2::token_cap(usize));
    uint32_t nt = 0;
    int rc = dflate3::team_inflate(OneLane(), &T, src, n, dst, usize, tok.data(), &nt);
    if (rc == dflate2::E_FALLBACK) { g_fallbacks++; return one_lane2(src, n, dst, usize, want_crc); }
    if (rc == OK) rc = dflate2::resolve_bytes(OneLane(), tok.data(), nt, dst, usize, src);
    if (rc == OK && dflate2::crc32_block4(OneLane(), dst, usize, g_crc4) != want_crc) rc = E_CRC;
    return rc;
}
static int cmd_inflate3(const char *path) {
    auto f = slurp(path); uint64_t ut; auto blocks = scan_blocks(f, &ut);
    size_t nbad = 0; uint64_t bytes = 0;
    dflate3::TeamStats st32{}, st16{}, st8{};
    for (size_t i = 0; i < blocks.size(); i++) {
        const Blk &b = blocks[i];
        const uint8_t *src = f.data() + b.coff + 12 + b.xlen; const uint32_t n = b.csize - 12 - b.xlen - 8;
        std::vector<uint8_t> a(b.usize + 1);
        const uint32_t want = bamcore::ld32(f.data() + b.coff + b.csize - 8);
        int rz = zlib_inflate(src, n, a.data(), b.usize);
        if (rz == 0 && (uint32_t)crc32(crc32(0L, Z_NULL, 0), a.data(), b.usize) != want) rz = -1;
        for (int g : {1, 8, 16, 32}) {
            std::vector<uint8_t> c(b.usize + 1);
            const int rc = g == 1 ? one_team3(src, n, c.data(), b.usize, want) : g == 8 ? emu_team3<8>(src, n, c.data(), b.usize, want, &st8, i % 16 == 5)
                         : g == 16 ? emu_team3<16>(src, n, c.data(), b.usize, want, &st16, false) : emu_team3<32>(src, n, c.data(), b.usize, want, &st32, i < 8 || i % 16 == 0);
            const bool ok = (rz == 0) == (rc == 0) && (rz != 0 || !memcmp(a.data(), c.data(), b.usize));
            if (!ok) { nbad++; fprintf(stderr, "block %zu: team decoder (%d lanes): rc %d (zlib %d)\n", i, g, rc, rz); }
        }
        bytes += b.usize;
    }
    for (auto [g, s] : {std::pair<int, dflate3::TeamStats *>{8, &st8}, {16, &st16}, {32, &st32}})
        fprintf(stderr, "team %d: deflate blocks %llu lanes started %llu dropped %llu\n", g, (unsigned long long)s->blocks, (unsigned long long)s->lanes_started, (unsigned long long)s->lanes_dropped);
    fprintf(stderr, "fallbacks %zu\n", g_fallbacks);
    printf("blocks %zu bytes %llu mismatches %zu\n", blocks.size(), (unsigned long long)bytes, nbad);
    return nbad ? 1 : 0;
}

// the MM / ML tags of every record as the direct route finds them (bam_core.cuh find_np_tags): "qname \t MM string | - \t ML values | -" per record
static int cmd_nptags(const char *path) {
    auto f = slurp(path); uint64_t n; auto blocks = scan_blocks(f, &n);
    std::vector<uint8_t> d(n + 16, 0);
    for (const Blk &b : blocks) if (one_lane(f.data() + b.coff + 12 + b.xlen, b.csize - 12 - b.xlen - 8, d.data() + b.uoff, b.usize)) { fprintf(stderr, "inflate failed\n"); return 3; }
    using namespace bamcore;
    if (n < 12 || memcmp(d.data(), "BAM\1", 4)) { fprintf(stderr, "not a BAM\n"); return 3; }
    uint64_t p = 8ull + ld32(d.data() + 4); const int32_t n_ref = ldi32(d.data() + p); p += 4;
    for (int32_t i = 0; i < n_ref; i++) { const uint32_t l = ld32(d.data() + p); p += 4 + l + 4; }
    while (p + 4 <= n) {
        const uint32_t bs = ld32(d.data() + p);
        if (bs < 32 || p + 4 + bs > n) { fprintf(stderr, "corrupt BAM record at %llu\n", (unsigned long long)p); return 3; }
        Rec R; R.load(d.data() + p);
        const uint8_t *mm, *ml; uint32_t mm_len, ml_cnt;
        find_np_tags(R.tags(), R.end(), &mm, &mm_len, &ml, &ml_cnt);
        fwrite(R.name(), 1, strlen((const char *)R.name()), stdout); putchar('\t');
        if (mm) fwrite(mm, 1, mm_len, stdout); else putchar('-');
        putchar('\t');
        if (ml) { for (uint32_t k = 0; k < ml_cnt; k++) printf(k ? ",%u" : "%u", ml[k]); if (!ml_cnt) putchar('.'); } else putchar('-');
        putchar('\n');
        p += 4 + bs;
    }
    return 0;
}

static int cmd_view(const char *path, uint64_t SEG, int depth, int argc, char **argv) {
    auto f = slurp(path); uint64_t n; auto blocks = scan_blocks(f, &n);
    std::vector<uint8_t> d(n + 16, 0);
    for (const Blk &b : blocks) if (one_lane(f.data() + b.coff + 12 + b.xlen, b.csize - 12 - b.xlen - 8, d.data() + b.uoff, b.usize)) { fprintf(stderr, "inflate failed\n"); return 3; }
    using namespace bamcore;
    if (n < 12 || memcmp(d.data(), "BAM\1", 4)) { fprintf(stderr, "not a BAM\n"); return 3; }
    uint64_t p = 8ull + ld32(d.data() + 4); const int32_t n_ref = ldi32(d.data() + p); p += 4;
    std::vector<uint32_t> name_off{0}; std::string names; std::vector<int32_t> lens;
    for (int32_t i = 0; i < n_ref; i++) { const uint32_t l = ld32(d.data() + p); p += 4; names.append((const char *)d.data() + p, l ? l - 1 : 0); name_off.push_back((uint32_t)names.size()); p += l; lens.push_back(ldi32(d.data() + p)); p += 4; }
    Refs F{n_ref, name_off.data(), names.data(), lens.data()};
    // segments over [p, n): entry[s] = first record start >= p + s*SEG
    const uint64_t p0 = p, nseg = n > p0 ? (n - p0 + SEG - 1) / SEG : 0;
    std::vector<uint64_t> entry(nseg + 1), exit_(nseg), bad(nseg, ~0ull); std::vector<uint32_t> cnt(nseg);
    OneLane one;
    size_t wrong_guesses = 0, rounds = 0;
    // depth < 0: real guesses (depth 4), but every (-depth)-th one is replaced by a wrong one (a few bytes into a record or,
    // every other time, far ahead): isolated wrong guesses must be repaired in one round without disturbing their neighbours
    for (uint64_t s = 0; s < nseg; s++) {
        entry[s] = s == 0 ? p0 : guess_entry(one, d.data(), n, p0 + s * SEG, n_ref, depth < 0 ? 4 : depth);
        if (depth < 0 && s && s % (uint64_t)(-depth) == 0) entry[s] = (s / (uint64_t)(-depth)) & 1 ? std::min<uint64_t>(n, entry[s] + 7) : std::min<uint64_t>(n, entry[s] + 5 * SEG + 3);
    }
    std::vector<char> dirty(nseg, 1);
    for (bool changed = true; changed;) {
        changed = false; rounds++;
        for (uint64_t s = 0; s < nseg; s++) if (dirty[s]) { bad[s] = ~0ull; exit_[s] = walk_chain(d.data(), n, entry[s], p0 + (s + 1) * SEG, &cnt[s], nullptr, &bad[s]); dirty[s] = 0; }
        std::vector<uint64_t> next(entry);                                   // the kernel reads the old entries and writes new ones
        for (uint64_t s = 0; s + 1 < nseg; s++) { bool ch = false; next[s + 1] = repaired_entry(entry.data(), exit_.data(), bad.data(), s, &ch); if (ch) { dirty[s + 1] = 1; changed = true; wrong_guesses++; } }
        entry.swap(next);
    }
    for (uint64_t s = 0; s < nseg; s++) if (bad[s] != ~0ull) { fprintf(stderr, "corrupt BAM record at %llu\n", (unsigned long long)bad[s]); return 3; }
    std::vector<uint64_t> rec;
    for (uint64_t s = 0; s < nseg; s++) { std::vector<uint64_t> o(cnt[s]); uint32_t c; uint64_t b = ~0ull; walk_chain(d.data(), n, entry[s], p0 + (s + 1) * SEG, &c, o.data(), &b); rec.insert(rec.end(), o.begin(), o.end()); }
    ViewParams V; memset(&V, 0, sizeof V); V.refid = -1;
    std::vector<int64_t> ivb, ive; std::string rg; uint64_t max_rec = 0, npass = 0;
    if (argc >= 11) {
        V.refid = atoi(argv[0]); V.min_mapq = atoi(argv[1]); V.exclude_flags = atoi(argv[2]); V.include_flags = atoi(argv[3]);
        V.beg = atoll(argv[4]); V.end = atoll(argv[5]);
        if (strcmp(argv[6], "-")) { char *t = strdup(argv[6]); for (char *q = strtok(t, ","); q; q = strtok(nullptr, ",")) V.flag_eq[V.n_flag_eq++] = atoi(q); free(t); }
        if (strcmp(argv[7], "-")) { rg = argv[7]; V.rg = rg.c_str(); V.rg_len = (uint32_t)rg.size(); V.have_rg = 1; }
        if (strcmp(argv[8], "-")) { FILE *fi = fopen(argv[8], "r"); long long a, b; while (fi && fscanf(fi, "%lld %lld", &a, &b) == 2) { ivb.push_back(a); ive.push_back(b); } if (fi) fclose(fi); }
        V.iv_beg = ivb.data(); V.iv_end = ive.data(); V.n_iv = ivb.size(); V.iv_exclude = atoi(argv[9]);
        max_rec = strtoull(argv[10], nullptr, 10);
        if (argc >= 13) { V.key_beg = atoll(argv[11]); V.key_end = atoll(argv[12]); }     // template window (wgbs_view_opts.key_beg / key_end)
    }
    if (argc >= 15 && !strcmp(argv[13], "firstkey")) {
        // bam_first_key_k (bamdev.cu): the first record of reference V.refid that passes the filters (key window ignored) with
        // template key >= KEY; prints its offset, or -1
        const long long key = atoll(argv[14]); ViewParams W = V; W.key_beg = 0; W.key_end = 0;
        long long best = -1;
        for (uint64_t o : rec) { Rec R; R.load(d.data() + o); if (template_key((int32_t)R.flag, R.refid, R.pos, R.nref, R.npos) >= key && passes(R, W)) { best = (long long)o; break; } }
        printf("%lld\n", best);
        return 0;
    }
    std::string out;
    for (uint64_t o : rec) {
        Rec R; R.load(d.data() + o);
        if (!R.consistent()) { fprintf(stderr, "inconsistent record at %llu\n", (unsigned long long)o); return 3; }
        if (!passes(R, V)) continue;
        if (max_rec && npass >= max_rec) break;
        npass++;
        CountSink cs; format_record(R, F, cs);
        const size_t at = out.size(); out.resize(at + cs.n);
        WriteSink<OneLane> ws; ws.o = &out[at]; format_record(R, F, ws);
        if (ws.n != cs.n) { fprintf(stderr, "count/write disagree\n"); return 4; }
    }
    fwrite(out.data(), 1, out.size(), stdout);
    fprintf(stderr, "records %zu segments %llu wrong_guesses %zu rounds %zu\n", rec.size(), (unsigned long long)nseg, wrong_guesses, rounds);
    return 0;
}

// A PART of a file (wgbs_dbam_open_part, bamdev.cu dbam_open_impl with part != nullptr): blocks [b0, b0 + nb) inflated, records
// indexed from inflated offset `first` on with the segment guess / walk / repair scheme, the record cut off by the end of the part
// ends the chain (tail), segments behind it hold no record of this part.  Prints "nrec tail" on stderr and the SAM text of the
// records on stdout; refs: "name,name,..." (the reference list of the file).
static int cmd_part(const char *path, uint64_t SEG, int depth, size_t b0, size_t nb, uint64_t first, const char *refs_csv) {
    auto f = slurp(path); uint64_t ut; auto all = scan_blocks(f, &ut);
    if (b0 + nb > all.size()) nb = all.size() - b0;
    uint64_t n = 0; for (size_t i = b0; i < b0 + nb; i++) n += all[i].usize;
    std::vector<uint8_t> d(n + 16, 0); uint64_t at = 0;
    for (size_t i = b0; i < b0 + nb; i++) { const Blk &b = all[i]; if (one_lane(f.data() + b.coff + 12 + b.xlen, b.csize - 12 - b.xlen - 8, d.data() + at, b.usize)) { fprintf(stderr, "inflate failed\n"); return 3; } at += b.usize; }
    using namespace bamcore;
    std::vector<uint32_t> name_off{0}; std::string names; std::vector<int32_t> lens;
    { std::string r(refs_csv); size_t p = 0; while (p <= r.size()) { size_t q = r.find(',', p); if (q == std::string::npos) q = r.size(); names.append(r, p, q - p); name_off.push_back((uint32_t)names.size()); lens.push_back(0); p = q + 1; } }
    const int32_t n_ref = (int32_t)lens.size();
    Refs F{n_ref, name_off.data(), names.data(), lens.data()};
    const uint64_t p0 = first; uint64_t nseg = n > p0 ? (n - p0 + SEG - 1) / SEG : 0, tail = n > p0 ? n : p0;
    std::vector<uint64_t> entry(nseg + 1), exit_(nseg), bad(nseg, ~0ull); std::vector<uint32_t> cnt(nseg);
    OneLane one;
    for (uint64_t s = 0; s < nseg; s++) entry[s] = s == 0 ? p0 : guess_entry(one, d.data(), n, p0 + s * SEG, n_ref, depth);
    std::vector<char> dirty(nseg, 1);
    for (bool changed = true; changed;) {
        changed = false;
        for (uint64_t s = 0; s < nseg; s++) if (dirty[s]) { bad[s] = ~0ull; exit_[s] = walk_chain(d.data(), n, entry[s], p0 + (s + 1) * SEG, &cnt[s], nullptr, &bad[s]); dirty[s] = 0; }
        std::vector<uint64_t> next(entry);
        for (uint64_t s = 0; s + 1 < nseg; s++) { bool ch = false; next[s + 1] = repaired_entry(entry.data(), exit_.data(), bad.data(), s, &ch); if (ch) { dirty[s + 1] = 1; changed = true; } }
        entry.swap(next);
    }
    uint64_t first_bad = ~0ull;
    for (uint64_t s = 0; s < nseg; s++) if (bad[s] < first_bad) first_bad = bad[s];
    if (first_bad != ~0ull) {
        const uint64_t o = first_bad; const uint32_t bs = o + 4 <= n ? ld32(d.data() + o) : 0;
        const bool cut = o + 4 > n || (bs >= 32 && o + 4 + (uint64_t)bs > n);
        if (!cut) { fprintf(stderr, "corrupt BAM record at %llu\n", (unsigned long long)o); return 3; }
        tail = o; nseg = o > p0 ? (o - p0) / SEG + 1 : 1;
    }
    std::vector<uint64_t> rec;
    for (uint64_t s = 0; s < nseg; s++) { std::vector<uint64_t> o(cnt[s]); uint32_t c; uint64_t b = ~0ull; walk_chain(d.data(), n, entry[s], p0 + (s + 1) * SEG, &c, o.data(), &b); rec.insert(rec.end(), o.begin(), o.end()); }
    std::string out;
    for (uint64_t o : rec) {
        Rec R; R.load(d.data() + o);
        if (!R.consistent()) { fprintf(stderr, "inconsistent record at %llu\n", (unsigned long long)o); return 3; }
        CountSink cs; format_record(R, F, cs);
        const size_t a2 = out.size(); out.resize(a2 + cs.n);
        WriteSink<OneLane> ws; ws.o = &out[a2]; format_record(R, F, ws);
    }
    fwrite(out.data(), 1, out.size(), stdout);
    fprintf(stderr, "nrec %zu tail %llu\n", rec.size(), (unsigned long long)tail);
    return 0;
}


// ---- table construction of the two-phase decoder against a plain canonical-code walk -----------------------------------------------
// N random complete code-length sets (literal/length: up to 286 symbols, distance: up to 30, lengths up to 15 bits, skewed so that
// long codes and large second-level tables occur): every symbol's code must decode to that symbol with that length through
// Decoder::probe (root + second level), in both memory layouts.  Sets whose tables exceed the arena must say E_FALLBACK.
static void random_lengths(std::mt19937_64 &rng, int n, int maxbits, uint8_t *out) {
    // split a unit of Kraft weight: start with one code of length 0, repeatedly split a random leaf (biased to deep ones) until n leaves
    std::vector<int> leaves{0};
    while ((int)leaves.size() < n) {
        size_t pick = rng() % leaves.size();
        if (rng() % 3) { size_t p2 = rng() % leaves.size(); if (leaves[p2] > leaves[pick]) pick = p2; }       // prefer deeper leaves: skew
        if (leaves[pick] >= maxbits) { bool any = false; for (size_t i = 0; i < leaves.size(); i++) if (leaves[i] < maxbits) { pick = i; any = true; break; } if (!any) break; }
        const int l = leaves[pick] + 1; leaves[pick] = l; leaves.push_back(l);
    }
    std::shuffle(leaves.begin(), leaves.end(), rng);
    for (int i = 0; i < n; i++) out[i] = i < (int)leaves.size() ? (uint8_t)leaves[i] : 0;
}
template <int SH>
static int check_tables(dflate2::Decoder<SH> &D, const uint8_t *ln, int nlen, int ndist, size_t *fallbacks) {
    const uint32_t S = (uint32_t)SH;
    for (int i = 0; i < nlen + ndist; i++) D.m.ln[(uint32_t)i << S] = ln[i];
    const int rc = D.both_tables(nlen, ndist);
    if (rc == dflate2::E_FALLBACK) { (*fallbacks)++; return 0; }
    if (rc != OK) { fprintf(stderr, "both_tables: rc %d on a complete code\n", rc); return 1; }
    for (int which = 0; which < 2; which++) {
        const int n = which ? ndist : nlen; const uint8_t *l = ln + (which ? nlen : 0);
        // canonical codes (RFC 1951 3.2.2)
        int cnt[16] = {0}, next[16] = {0};
        for (int i = 0; i < n; i++) cnt[l[i]]++;
        cnt[0] = 0; int code = 0;
        for (int b = 1; b <= 15; b++) { code = (code + cnt[b - 1]) << 1; next[b] = code; }
        for (int sy = 0; sy < n; sy++) {
            if (!l[sy]) continue;
            const uint32_t c = (uint32_t)next[l[sy]]++, rev = brev32(c) >> (32 - l[sy]);
            for (uint32_t hi = 0; hi < 4; hi++) {                          // the bits behind the code must not matter
                D.bb = (uint64_t)rev | ((uint64_t)(hi * 0x9e3779b9u) << l[sy]);
                const uint32_t e = which ? D.probe(D.dt_off, dflate2::DB) : D.probe(0, dflate2::LB);
                const uint32_t want = which ? dflate2::dist_entry((uint32_t)sy, l[sy]) : dflate2::litlen_entry((uint32_t)sy, l[sy]);
                if (e != want) { fprintf(stderr, "table %d symbol %d len %d: entry %08x, want %08x\n", which, sy, l[sy], e, want); return 1; }
            }
        }
    }
    return 0;
}
// the team's table construction (inflate3_core.cuh: team_tables) must leave the SAME arena as Decoder::both_tables: entry by entry
template <int NL>
static int check_team_tables(const uint8_t *ln, int nlen, int ndist, dflate2::Decoder<0> &ref, int ref_rc) {
    static dflate2::HostLane H; static EmuSharedN<NL> sh;
    for (int i = 0; i < nlen + ndist; i++) H.ln[i] = ln[i];
    int rcs[NL]; uint32_t dts[NL];
    if (NL == 1) rcs[0] = dflate3::team_tables(OneLane(), H.mem(), (uint32_t)nlen, (uint32_t)ndist, &dts[0]);
    else {
        std::vector<std::thread> th;
        for (int l = 0; l < NL; l++) th.emplace_back([&, l]() { rcs[l] = dflate3::team_tables(EmuLanesN<NL>{l, &sh}, H.mem(), (uint32_t)nlen, (uint32_t)ndist, &dts[l]); });
        for (auto &t : th) t.join();
        for (int l = 1; l < NL; l++) if (rcs[l] != rcs[0] || (rcs[0] == OK && dts[l] != dts[0])) { fprintf(stderr, "team_tables: lanes disagree\n"); return 1; }
    }
    if (rcs[0] != ref_rc) { fprintf(stderr, "team_tables (%d lanes): rc %d, one lane %d\n", NL, rcs[0], ref_rc); return 1; }
    if (ref_rc != OK) return 0;
    if (dts[0] != ref.dt_off) { fprintf(stderr, "team_tables (%d lanes): distance root at %u, one lane %u\n", NL, dts[0], ref.dt_off); return 1; }
    // every entry a probe can reach: walk both decoders over all codes (second-level offsets may differ: compare what probes return)
    for (int which = 0; which < 2; which++) {
        const int n = which ? ndist : nlen; const uint8_t *l = ln + (which ? nlen : 0);
        int cnt[16] = {0}, next[16] = {0};
        for (int i = 0; i < n; i++) cnt[l[i]]++;
        cnt[0] = 0; int code = 0;
        for (int b = 1; b <= 15; b++) { code = (code + cnt[b - 1]) << 1; next[b] = code; }
        dflate2::Decoder<0> D = ref; D.m = H.mem(); D.dt_off = dts[0];
        for (int sy = 0; sy < n; sy++) {
            if (!l[sy]) continue;
            const uint32_t c = (uint32_t)next[l[sy]]++, rev = brev32(c) >> (32 - l[sy]);
            for (uint32_t hi = 0; hi < 4; hi++) {
                D.bb = (uint64_t)rev | ((uint64_t)(hi * 0x9e3779b9u) << l[sy]);
                const uint32_t e = which ? D.probe(D.dt_off, dflate2::DB) : D.probe(0, dflate2::LB);
                const uint32_t want = which ? dflate2::dist_entry((uint32_t)sy, l[sy]) : dflate2::litlen_entry((uint32_t)sy, l[sy]);
                if (e != want) { fprintf(stderr, "team table %d (%d lanes) symbol %d len %d: entry %08x, want %08x\n", which, NL, sy, l[sy], e, want); return 1; }
            }
        }
    }
    return 0;
}
static int cmd_tables(long N, unsigned seed) {
    std::mt19937_64 rng(seed); size_t bad = 0, fallbacks = 0;
    static dflate2::HostLane H;
    static uint8_t dummy[64] = {3, 0};
    for (long t = 0; t < N; t++) {
        uint8_t ln[320] = {0};
        const int nlen = 257 + (int)(rng() % 30), ndist = 1 + (int)(rng() % 30);
        const int maxl = 9 + (int)(rng() % 7), maxd = 5 + (int)(rng() % 11);
        random_lengths(rng, nlen, maxl, ln);
        if (ndist >= 2) random_lengths(rng, ndist, maxd, ln + nlen); else ln[nlen] = 1;
        dflate2::Decoder<0> D0; D0.init(H.mem(), dummy, 2, dummy + 8, 0, nullptr);
        const size_t fb0 = fallbacks;
        bad += check_tables(D0, ln, nlen, ndist, &fallbacks);
        const int rc0 = fallbacks != fb0 ? dflate2::E_FALLBACK : OK;
        bad += check_team_tables<1>(ln, nlen, ndist, D0, rc0);
        if (t % 16 == 0) bad += check_team_tables<32>(ln, nlen, ndist, D0, rc0);
        if (t % 16 == 8) bad += check_team_tables<8>(ln, nlen, ndist, D0, rc0);
    }
    printf("tables %ld fallbacks %zu mismatches %zu\n", N, fallbacks, bad);
    return bad ? 1 : 0;
}

static int cmd_fmtg(long N, unsigned seed) {
    std::mt19937_64 rng(seed); size_t bad = 0;
    auto test = [&](uint32_t u) {
        float f; memcpy(&f, &u, 4);
        char a[40], b[40]; snprintf(a, sizeof a, "%g", f);
        const int k = bamcore::fmt_g(f, b); b[k] = 0;
        if (strcmp(a, b)) { if (bad < 20) fprintf(stderr, "%08x: printf '%s' core '%s'\n", u, a, b); bad++; }
    };
    const float edge[] = {0.f, -0.f, 1.f, -1.f, 0.5f, 0.1f, 100000.f, 999999.f, 999999.5f, 1000000.f, 1e-4f, 9.9999e-5f, 1e-5f, 123456.5f, 1234565.f, 0.25f, 2.f, 1.5f,
                          3.4028235e38f, 1.17549435e-38f, 1e-45f, 16777216.f, 8388608.5f, 0.000123456789f, 1e10f, 1e-10f, INFINITY, -INFINITY, NAN, 2.5f, 0.3f, 1e6f, 1e5f, 99999.95f, 0.00001f};
    for (float e : edge) { uint32_t u; memcpy(&u, &e, 4); test(u); }
    for (long i = 0; i < N; i++) test((uint32_t)rng());
    // values near the style switches and with short decimal expansions (tags are usually like 0.25 or 12.5)
    for (long i = 0; i < N / 4; i++) { const float v = (float)((double)(rng() % 20000000) / 16.0); uint32_t u; memcpy(&u, &v, 4); test(u); }
    for (long i = 0; i < N / 4; i++) { const float v = (float)((double)(rng() % 100000) / 1e3); uint32_t u; memcpy(&u, &v, 4); test(u); }
    printf("fmtg mismatches %zu\n", bad);
    return bad ? 1 : 0;
}

int main(int argc, char **argv) {
    for (uint32_t i = 0; i < 256; i++) g_crc_table[i] = crc_table_entry(i);
    for (uint32_t k = 0; k < 4; k++) for (uint32_t i = 0; i < 256; i++) g_crc4[k * 256 + i] = dflate2::crc_slice_entry(k, i);
    if (argc >= 3 && !strcmp(argv[1], "inflate")) return cmd_inflate(argv[2], argc > 3 ? atoi(argv[3]) : 0);
    if (argc >= 3 && !strcmp(argv[1], "inflate3")) return cmd_inflate3(argv[2]);
    if (argc >= 3 && !strcmp(argv[1], "nptags")) return cmd_nptags(argv[2]);
    if (argc >= 5 && !strcmp(argv[1], "view")) return cmd_view(argv[2], strtoull(argv[3], nullptr, 10), atoi(argv[4]), argc - 5, argv + 5);
    if (argc >= 4 && !strcmp(argv[1], "fmtg")) return cmd_fmtg(atol(argv[2]), (unsigned)atoi(argv[3]));
    if (argc >= 4 && !strcmp(argv[1], "tables")) return cmd_tables(atol(argv[2]), (unsigned)atoi(argv[3]));
    if (argc >= 9 && !strcmp(argv[1], "part")) return cmd_part(argv[2], strtoull(argv[3], nullptr, 10), atoi(argv[4]), strtoull(argv[5], nullptr, 10), strtoull(argv[6], nullptr, 10), strtoull(argv[7], nullptr, 10), argv[8]);
    fprintf(stderr, "usage: see the header of tests/bamdev_core_check.cpp\n");
    return 2;
}
