// TEST INFRASTRUCTURE: builds wgbs_tools_b200/csrc/pat_core.cuh -- the per-line logic of the staged two-pass tile parser for pat
// text (pat_tiles_k) -- as plain host C++ and runs the kernel's two passes sequentially (tile by tile, thread by thread, the same
// ownership / prefix / write logic), against the default parser's algorithm (pat_lines_k + pat_pack_k: byte loops per line).
//   usage: pat_core_check FILE.pat      ->  "lines N words W err E mismatches M"
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#define __host__
#define __device__
#define __forceinline__ inline
#include "../wgbs_tools_b200/csrc/pat_core.cuh"

struct Out { std::vector<uint32_t> idx, len, cnt, off, pool; uint32_t err = 0; };

// the default parser (pat.cu: pat_lines_k, scan, pat_pack_k), line by line
static bool stoi_field(const char *t, uint32_t s, uint32_t e, int32_t *out) {
    while (s < e && (t[s] == ' ' || (t[s] >= 9 && t[s] <= 13))) s++;
    bool neg = false;
    if (s < e && (t[s] == '+' || t[s] == '-')) { neg = t[s] == '-'; s++; }
    if (s >= e || t[s] < '0' || t[s] > '9') return false;
    long long v = 0;
    while (s < e && t[s] >= '0' && t[s] <= '9') { v = v * 10 + (t[s] - '0'); if (v > 0x80000000LL) return false; s++; }
    if (neg) v = -v;
    if (v > 0x7fffffffLL || v < -0x80000000LL) return false;
    *out = (int32_t)v; return true;
}
static Out reference(const std::string &x) {
    Out o; const char *t = x.data(); const uint32_t n = (uint32_t)x.size();
    uint32_t s = 0;
    while (s < n) {
        uint32_t e = s; while (e < n && t[e] != '\n') e++;
        uint32_t oi = 0, ol = 0, oc = 0, ps = s;
        if (e > s) {
            uint32_t tab[4]; int nt = 0;
            for (uint32_t p = s; p < e && nt < 4; p++) if (t[p] == '\t') tab[nt++] = p;
            if (nt < 3) o.err |= 1;
            else {
                const uint32_t cend = nt >= 4 ? tab[3] : e; int32_t vi, vc;
                if (!stoi_field(t, tab[0] + 1, tab[1], &vi) || !stoi_field(t, tab[2] + 1, cend, &vc)) o.err |= 2;
                else { oi = (uint32_t)vi; oc = (uint32_t)vc; ol = tab[2] - tab[1] - 1; ps = tab[1] + 1; }
            }
        }
        o.idx.push_back(oi); o.len.push_back(ol); o.cnt.push_back(oc); o.off.push_back((uint32_t)o.pool.size());
        for (uint32_t b = 0; b < ol; b += 16) { uint32_t w = 0, m = ol - b < 16 ? ol - b : 16; for (uint32_t k = 0; k < m; k++) w |= sym_code(t[ps + b + k]) << (30 - 2 * k); o.pool.push_back(w); }
        s = e + 1;
    }
    return o;
}

// pat_tiles_k, emulated: PASS 0 fills tile totals, PASS 1 writes with the exclusive prefixes
static Out tiles(const std::string &x) {
    const char *text = x.data(); const uint32_t n = (uint32_t)x.size();
    const uint32_t ntiles = (n + PS_TILE - 1) / PS_TILE;
    std::vector<uint32_t> tl(ntiles + 1, 0), tw(ntiles + 1, 0);
    Out o;
    for (int pass = 0; pass < 2; pass++) {
        if (pass == 1) {
            uint32_t a = 0, b = 0;
            for (uint32_t t = 0; t <= ntiles; t++) { const uint32_t x1 = tl[t], x2 = tw[t]; tl[t] = a; tw[t] = b; a += x1; b += x2; }
            o.idx.assign(tl[ntiles], 0xdeadbeef); o.len = o.cnt = o.off = o.idx; o.pool.assign(tw[ntiles], 0xdeadbeef);
        }
        for (uint32_t tile = 0; tile < ntiles; tile++) {
            std::vector<unsigned char> sm(PS_TILE, 0); std::vector<unsigned long long> nl(PS_T, 0), tb(PS_T, 0);
            const uint32_t t0 = tile * (uint32_t)PS_TILE, t1 = (n - t0 > (uint32_t)PS_TILE) ? t0 + PS_TILE : n;
            memcpy(sm.data(), text + t0, t1 - t0);
            for (int tid = 0; tid < PS_T; tid++) for (int b = 0; b < PS_SPAN; b++) {
                const unsigned char c = sm[tid * PS_SPAN + b];
                if (c == '\n') nl[tid] |= 1ull << b;
                if (c == '\t') tb[tid] |= 1ull << b;
            }
            PatTile T; T.g = text; T.n = n; T.sm = sm.data(); T.t0 = t0; T.t1 = t1; T.nl = nl.data(); T.tab = tb.data();
            std::vector<uint32_t> my_l(PS_T, 0), my_w(PS_T, 0);
            for (int tid = 0; tid < PS_T; tid++) {
                const uint32_t span0 = t0 + tid * PS_SPAN; const bool line0 = tile == 0 && tid == 0 && n > 0;
                if (line0) { my_l[tid]++; my_w[tid] += (pat_line(T, 0, false).len + 15) >> 4; }
                for (unsigned long long m = nl[tid]; m;) {
                    const int b = pat_ctz64(m); m &= m - 1;
                    const uint32_t s = span0 + (uint32_t)b + 1;
                    if (s < n) { my_l[tid]++; my_w[tid] += (pat_line(T, s, false).len + 15) >> 4; }
                }
            }
            if (pass == 0) { for (int tid = 0; tid < PS_T; tid++) { tl[tile] += my_l[tid]; tw[tile] += my_w[tid]; } continue; }
            uint32_t line = tl[tile], w = tw[tile];
            for (int tid = 0; tid < PS_T; tid++) {
                const uint32_t span0 = t0 + tid * PS_SPAN; bool first = tile == 0 && tid == 0 && n > 0;
                uint32_t ln = line, ow = w; unsigned long long m = nl[tid];
                while (first || m) {
                    uint32_t s = 0;
                    if (first) first = false;
                    else { const int b = pat_ctz64(m); m &= m - 1; s = span0 + (uint32_t)b + 1; if (s >= n) continue; }
                    const PatRec r = pat_line(T, s, true);
                    o.err |= r.err;
                    const uint32_t L = r.err ? 0u : r.len;
                    o.idx[ln] = r.idx; o.len[ln] = L; o.cnt[ln] = r.cnt; o.off[ln] = ow;
                    for (uint32_t b = 0; b < L; b += 16) {
                        uint32_t wd = 0; const uint32_t mm = L - b < 16 ? L - b : 16;
                        for (uint32_t k = 0; k < mm; k++) wd |= sym_code((char)T.byte(r.ps + b + k)) << (30 - 2 * k);
                        o.pool[ow + (b >> 4)] = wd;
                    }
                    ow += (r.len + 15) >> 4; ln++;
                }
                line += my_l[tid]; w += my_w[tid];
            }
        }
    }
    return o;
}

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    FILE *f = fopen(argv[1], "rb"); if (!f) { perror(argv[1]); return 2; }
    std::string x; char buf[1 << 16]; size_t k;
    while ((k = fread(buf, 1, sizeof buf, f)) > 0) x.append(buf, k);
    fclose(f);
    const Out a = reference(x), b = tiles(x);
    size_t bad = 0;
    if (a.err != b.err) bad++;
    if (a.idx.size() != b.idx.size()) bad++;
    else if (!a.err) {
        for (size_t i = 0; i < a.idx.size(); i++) bad += a.idx[i] != b.idx[i] || a.len[i] != b.len[i] || a.cnt[i] != b.cnt[i] || a.off[i] != b.off[i];
        if (a.pool != b.pool) bad++;
    }
    printf("lines %zu words %zu err %u mismatches %zu\n", b.idx.size(), b.pool.size(), b.err, bad);
    return bad ? 1 : 0;
}
