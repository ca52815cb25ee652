// TEST INFRASTRUCTURE: builds wgbs_tools_b200/csrc/cview_core.cuh -- the per-record logic the device kernels of view.cu
// run -- as plain host C++ (g++), so tests/test_view.py can compare it with the reference `cview` executable without a
// GPU.  Not part of libwgbs_b200.so.
//   usage: cview_core_check BLOCKS.tsv strict strip no_gaps min_cpgs [pre.tsv] < in.pat > out.pat
//   BLOCKS.tsv: "start\tend" per line, sorted by start; pre.tsv: closed "lo\thi" ranges of start indices
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../wgbs_tools_b200/csrc/cview_core.cuh"

static std::vector<int32_t> col(const char *path, int c) {
    std::vector<int32_t> v; FILE *f = fopen(path, "r"); if (!f) { perror(path); exit(2); }
    long a, b; while (fscanf(f, "%ld %ld", &a, &b) == 2) v.push_back((int32_t)(c ? b : a));
    fclose(f); return v;
}
struct Printer {
    const std::string *chrom, *pat, *tail; int32_t s;
    void operator()(int32_t start, uint32_t a, uint32_t len) const {
        printf("%s\t%d\t%s\t%s\n", chrom->c_str(), start, pat->substr(a, len).c_str(), tail->c_str());
    }
};
int main(int argc, char **argv) {
    if (argc < 6) return 2;
    std::vector<int32_t> bs = col(argv[1], 0), be = col(argv[1], 1), pm(bs.size()), pl, ph;
    int32_t run = INT32_MIN; for (size_t i = 0; i < bs.size(); i++) { run = be[i] > run ? be[i] : run; pm[i] = run; }
    if (argc > 6) { pl = col(argv[6], 0); ph = col(argv[6], 1); }
    CviewParams p{bs.data(), be.data(), pm.data(), (int32_t)bs.size(), pl.data(), ph.data(), (int32_t)pl.size(), atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5])};
    char *line = nullptr; size_t cap = 0; ssize_t n;
    while ((n = getline(&line, &cap, stdin)) > 0) {
        if (line[n - 1] == '\n') line[--n] = 0;
        if (!n) continue;
        std::string l(line), chrom, idx, pat, tail;
        size_t t1 = l.find('\t'), t2 = l.find('\t', t1 + 1), t3 = l.find('\t', t2 + 1);
        chrom = l.substr(0, t1); idx = l.substr(t1 + 1, t2 - t1 - 1); pat = l.substr(t2 + 1, t3 - t2 - 1); tail = l.substr(t3 + 1);
        std::vector<uint32_t> w((pat.size() + 15) / 16 + 1, 0u);
        for (size_t k = 0; k < pat.size(); k++) {
            uint32_t c = pat[k] == 'C' ? 1u : pat[k] == 'H' ? 2u : pat[k] == 'T' ? 3u : 0u;
            w[k >> 4] |= c << (30 - 2 * (k & 15));
        }
        Printer pr{&chrom, &pat, &tail, atoi(idx.c_str())};
        cview_record(p, pr.s, (uint32_t)pat.size(), w.data(), pr);
    }
    return 0;
}
