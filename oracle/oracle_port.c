/*
 * oracle_port.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C CPU restatement of the wgbstools per-CpG hot path (reference nloyfer/wgbs_tools @ v0.3.0),
 * written to be read next to the reference and to serve as the checker for the CUDA path.
 * It is NOT part of the product: nothing in wgbs_tools_b200/ imports, links or executes it.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may load liboracle_port.so.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py checks every function here byte-for-byte against the
 * UNMODIFIED reference executables compiled from /root/reference/src by oracle/Makefile into oracle/_ref/
 * (the reference ships no golden vectors for this path whose inputs are available).
 *
 * Each function cites the reference file:line it restates (paths relative to /root/reference/src).
 */
#include <ctype.h>
#include <errno.h>
#include <limits.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------
 * small helpers
 * ---------------------------------------------------------------------------------------------- */
typedef struct { char *p; long n, cap; } sbuf;
static void sb_put(sbuf *b, const char *s, long n) {
    if (b->n + n + 1 > b->cap) { b->cap = (b->n + n + 1) * 2 + 64; b->p = (char *)realloc(b->p, b->cap); }
    memcpy(b->p + b->n, s, n); b->n += n; b->p[b->n] = 0;
}
static void sb_putc(sbuf *b, char c) { sb_put(b, &c, 1); }

/* std::stoi semantics (strtol base 10, throw if no digits / out of int range). returns 0 on "throw". */
static int cxx_stoi(const char *s, long n, long *out) {
    char tmp[64]; if (n > 63) n = 63; memcpy(tmp, s, n); tmp[n] = 0;
    char *end; long v = strtol(tmp, &end, 10);
    if (end == tmp) return 0;
    if (v > INT_MAX || v < INT_MIN) return 0;
    *out = v; return 1;
}

/* std::stoul semantics (strtoul base 10: optional blanks and sign, wraps negatives, throws when no digits / > ULONG_MAX) */
static int cxx_stoul(const char *s, long n, unsigned long *out) {
    char tmp[64]; if (n > 63) n = 63; memcpy(tmp, s, n); tmp[n] = 0;
    char *end; errno = 0; unsigned long v = strtoul(tmp, &end, 10);
    if (end == tmp || errno == ERANGE) return 0;
    *out = v; return 1;
}

typedef struct { const char *s; long n; } tok;
/* line2tokens: pipeline_wgbs/patter_utils.cpp:9-18 (getline on '\t': no trailing empty token) */
static int split_tabs(const char *line, long n, tok *t, int maxt) {
    int k = 0; long st = 0;
    for (long i = 0; i <= n; i++) {
        if (i == n || line[i] == '\t') {
            if (i == n && st == n) break;               /* nothing after the last delimiter */
            if (k < maxt) { t[k].s = line + st; t[k].n = i - st; }
            k++; st = i + 1;
        }
    }
    return k;
}
static int tok_eq(tok a, const char *lit) { long l = (long)strlen(lit); return a.n == l && !memcmp(a.s, lit, l); }
static int tok_pref(tok a, const char *lit) { long l = (long)strlen(lit); return a.n >= l && !memcmp(a.s, lit, l); }

/* ------------------------------------------------------------------------------------------------
 * a4  clean_CIGAR  -- pipeline_wgbs/patter_utils.cpp:209-251
 * returns 0 on any condition where the reference throws (read becomes "invalid").
 * ---------------------------------------------------------------------------------------------- */
static int clean_cigar(const char *seq, long slen, const char *cig, long clen, sbuf *adj) {
    adj->n = 0; if (adj->p) adj->p[0] = 0; else sb_put(adj, "", 0);
    /* phase 1: tokenise (stoi on the digit run before every op char) */
    long nops = 0; long *nums = (long *)malloc(sizeof(long) * (clen + 1)); char *ops = (char *)malloc(clen + 1);
    long ds = 0;
    for (long i = 0; i < clen; i++) {
        if (isdigit((unsigned char)cig[i])) continue;
        long v; if (!cxx_stoi(cig + ds, i - ds, &v) || i == ds) { free(nums); free(ops); return 0; }
        nums[nops] = v; ops[nops++] = cig[i]; ds = i + 1;
    }
    long pos = 0; int ok = 1;
    for (long k = 0; k < nops && ok; k++) {
        char ch = ops[k]; long num = nums[k];
        if (ch == 'M' || ch == '=' || ch == 'X') {
            long take = num; if (take > slen - pos) take = slen - pos;
            sb_put(adj, seq + pos, take);
            if (num > slen - pos) ok = 0;               /* seq.substr(num, ..) -> out_of_range */
            pos += num;
        } else if (ch == 'D' || ch == 'N') {
            for (long j = 0; j < num; j++) sb_putc(adj, 'N');
        } else if (ch == 'I' || ch == 'S') {
            if (num > slen - pos) ok = 0;
            pos += num;
        } else if (ch == 'H') {
        } else ok = 0;                                   /* Unknown CIGAR character */
    }
    free(nums); free(ops);
    return ok;
}

/* strip_pat -- patter_utils.cpp:270-280. returns -1 if empty/all dots, else #leading dots removed */
static int strip_pat(sbuf *p) {
    long e = p->n; while (e > 0 && p->p[e - 1] == '.') e--;
    p->n = e; p->p[e] = 0;
    if (e == 0) return -1;
    long s = 0; while (p->p[s] == '.') s++;
    if (s > 0) { memmove(p->p, p->p + s, e - s); p->n = e - s; p->p[p->n] = 0; }
    return (int)s;
}

/* CpG dictionary of the region: sorted loci[], idx[] (a3: patter.cpp:14-42, 81-93) */
typedef struct { const int *loci; const int *idx; int n; long bsize; } dict_t;
static int dict_find(const dict_t *d, long locus) {     /* conv[locus] ? position : -1 */
    int lo = 0, hi = d->n;
    while (lo < hi) { int m = (lo + hi) >> 1; if (d->loci[m] < locus) lo = m + 1; else hi = m; }
    return (lo < d->n && d->loci[lo] == locus) ? lo : -1;
}

typedef struct {
    int min_cpg, clip, nanopore, combine_mods; float np_thresh; char cpc_call;
    int is_pe;
    long line_i, nr_pairs, nr_empty, nr_short, nr_invalid;
} patter_t;

/* a5 is_bottom -- patter_utils.cpp:163-168 */
static int is_bottom(int flag, int pe) {
    if (pe) return ((flag & 0x53) == 83) || ((flag & 0xA3) == 163);
    return (flag & 0x10) == 16;
}

/* a6 compareSeqToRef -- patter.cpp:96-184.  returns first CpG index or -1 */
static int compare_seq_to_ref(const patter_t *P, const dict_t *d, const char *seq, long len, long start_locus,
                              int flag, sbuf *pat) {
    int bottom = is_bottom(flag, P->is_pe);
    int shift = bottom ? 1 : 0;
    char ref_chr = bottom ? 'G' : 'C', unmeth = bottom ? 'A' : 'T';
    int first_ind = -1;
    pat->n = 0; sb_put(pat, "", 0);
    for (long i = 0; i < len; i++) {
        if (start_locus + i > d->bsize - 1) continue;
        int di = dict_find(d, start_locus + i);
        if (di < 0) continue;
        long j = i + shift;
        char s = (j < len) ? seq[j] : 0;                 /* std::string::operator[](size()) == '\0' */
        char st = '.';
        int cpg;
        if (!shift) cpg = (j < len - 1) && (seq[j] == 'C' || seq[j] == 'T') && (seq[j + 1] == 'G');
        else        cpg = (j > 0) && (s == 'G' || s == 'A') && (seq[j - 1] == 'C');
        if (cpg) { if (s == unmeth) st = 'T'; else if (s == ref_chr) st = 'C'; }
        if (!((j >= P->clip) && (j < len - P->clip))) st = '.';
        if (first_ind < 0 && st != '.') first_ind = d->idx[di];
        if (first_ind > 0) sb_putc(pat, st);
    }
    if (strip_pat(pat)) return -1;                        /* nonzero: -1 (empty) or >0 (cannot happen) */
    return first_ind;
}

/* ------------------------------------------------------------------------------------------------
 * a9-a11  MM/ML path -- pipeline_wgbs/ont.cpp
 * ---------------------------------------------------------------------------------------------- */
typedef struct { int *v; int n, cap; } ivec;
static void iv_push(ivec *a, int x) { if (a->n == a->cap) { a->cap = a->cap * 2 + 16; a->v = (int *)realloc(a->v, sizeof(int) * a->cap); } a->v[a->n++] = x; }
static void iv_insert(ivec *a, int at, int x) { iv_push(a, 0); memmove(a->v + at + 1, a->v + at, sizeof(int) * (a->n - 1 - at)); a->v[at] = x; }

/* split_by_comma -- patter_utils.cpp:83-94 (istream >> int, skip one ',' after each) */
static void split_by_comma(const char *s, long n, ivec *out) {
    out->n = 0; long i = 0;
    while (1) {
        while (i < n && isspace((unsigned char)s[i])) i++;
        long st = i; if (i < n && (s[i] == '+' || s[i] == '-')) i++;
        long ds = i; while (i < n && isdigit((unsigned char)s[i])) i++;
        if (i == ds) return;
        char tmp[32]; long l = i - st; if (l > 31) return; memcpy(tmp, s + st, l); tmp[l] = 0;
        long v = strtol(tmp, NULL, 10); if (v > INT_MAX || v < INT_MIN) return;  /* failbit */
        iv_push(out, (int)v);
        if (i < n && s[i] == ',') i++;
    }
}
/* trim_from_first_comma -- ont.cpp:335-344 */
static void after_first_comma(const char *s, long n, const char **o, long *on) {
    const char *c = (const char *)memchr(s, ',', n);
    if (!c) { *o = s; *on = 0; return; }
    *o = c + 1; *on = n - (c + 1 - s);
}

typedef struct { ivec mm, ml, mm_h, ml_h; int np_dot; } npfields;

/* parse_np_fields_by_mod + subset_to_Cm_section + find_Cm_substring + get_np_tags
 * -- ont.cpp:269-308, 361-416, 310-333, 418-438.  returns 0 when the reference throws. */
static int parse_by_mod(const tok *t, int nt, char mod, int *np_dot, ivec *MM, ivec *ML) {
    MM->n = 0; ML->n = 0;
    const char *mm = NULL, *ml = NULL; long mmn = 0, mln = 0; int has_ml = 0;
    for (int j = 11; j < nt; j++) {
        if (tok_pref(t[j], "MM:Z:") || tok_pref(t[j], "Mm:Z:")) { mm = t[j].s + 5; mmn = t[j].n - 5; }
        else if (tok_pref(t[j], "ML:B:C") || tok_pref(t[j], "Ml:B:C")) { ml = t[j].s + 6; mln = t[j].n - 6; has_ml = 1; }
    }
    if (!mm || mmn == 0) return 1;
    if (!has_ml) { ml = ""; mln = 0; }
    /* find section */
    long st = 0; int pos = 0, found = 0; const char *sec = NULL; long secn = 0;
    for (long i = 0; i <= mmn; i++) {
        if (i == mmn || mm[i] == ';') {
            if (i == mmn && st == mmn) break;
            if (i - st >= 3 && mm[st] == 'C' && mm[st + 1] == '+' && mm[st + 2] == mod) { found = 1; sec = mm + st; secn = i - st; break; }
            pos++; st = i + 1;
        }
    }
    ivec orig = {0}, mlv = {0}; int ok = 1;
    char *ml_own = NULL;
    if (!found) { mln = 0; ml = ""; sec = ""; secn = 0; }
    else {
        *np_dot = !((secn > 3) && (sec[3] == '?'));
        const char *a; long an; after_first_comma(sec, secn, &a, &an);
        split_by_comma(a, an, &orig);
        if (mln != 0) {
            const char *b; long bn; after_first_comma(ml, mln, &b, &bn);
            split_by_comma(b, bn, &mlv);
            int nr = orig.n;
            if (nr == 0) { mln = 0; ml = ""; }
            else {
                if ((mlv.n % nr != 0) && mlv.n > 0) { ok = 0; goto done; }
                if (mlv.n >= (pos + 1) * nr) {
                    /* ML_str = "," + slice */
                    ivec sl = {0}; for (int q = pos * nr; q < (pos + 1) * nr; q++) iv_push(&sl, mlv.v[q]);
                    free(mlv.v); mlv = sl; ml_own = (char *)1;     /* marks: mlv already is the final ML */
                }
            }
        }
    }
    {
        /* back in parse_np_fields_by_mod (ont.cpp:288-307) */
        const char *a; long an; ivec o2 = {0};
        if (secn) { after_first_comma(sec, secn, &a, &an); split_by_comma(a, an, &o2); }
        if (mln == 0) { for (int q = 0; q < o2.n; q++) iv_push(ML, 255); }
        else {
            if (!ml_own) { const char *b; long bn; after_first_comma(ml, mln, &b, &bn); split_by_comma(b, bn, &mlv); }
            for (int q = 0; q < mlv.n; q++) iv_push(ML, mlv.v[q]);
            if (o2.n != ML->n) { ok = 0; free(o2.v); goto done; }
        }
        int p = 0; for (int q = 0; q < o2.n; q++) { p += o2.v[q]; iv_push(MM, p++); }
        free(o2.v);
    }
done:
    free(orig.v); free(mlv.v);
    return ok;
}

/* parse_np_fields -- ont.cpp:223-267 */
static int parse_np_fields(const patter_t *P, const tok *t, int nt, npfields *F) {
    ivec cm = {0}, cl = {0}; int ok = 1;
    F->np_dot = 0;
    if (!parse_by_mod(t, nt, 'h', &F->np_dot, &F->mm_h, &F->ml_h)) { ok = 0; goto out; }
    if (!parse_by_mod(t, nt, 'm', &F->np_dot, &F->mm, &F->ml)) { ok = 0; goto out; }
    {
        int np_dot_m = F->np_dot;
        if (!parse_by_mod(t, nt, 'C', &F->np_dot, &cm, &cl)) { ok = 0; goto out; }
        if (cm.n && P->cpc_call != '.') {
            ivec *tv = (P->cpc_call == 'H') ? &F->mm_h : &F->mm;
            ivec *tl = (P->cpc_call == 'H') ? &F->ml_h : &F->ml;
            int n0 = tv->n; int *ex = (int *)malloc(sizeof(int) * (n0 + 1)); memcpy(ex, tv->v, sizeof(int) * n0);
            for (int q = 0; q < cm.n; q++) {
                int p = cm.v[q], present = 0;
                for (int r = 0; r < n0; r++) if (ex[r] == p) { present = 1; break; }
                if (present) continue;
                int at = 0; while (at < tv->n && tv->v[at] < p) at++;      /* lower_bound */
                iv_insert(tv, at, p); iv_insert(tl, at, 255);
            }
            free(ex);
        }
        F->np_dot = np_dot_m;
    }
out:
    free(cm.v); free(cl.v);
    return ok;
}

/* make_meth_mask -- ont.cpp:22-87 */
static void make_meth_mask(const patter_t *P, const npfields *F, const char *ws, long n, char *mask) {
    int C = 0, mi = 0, hi = 0; float th = P->np_thresh;
    memset(mask, 'E', n);
    for (long i = 0; i < n; i++) {
        if (ws[i] != 'C') continue;
        char cur = 'N';
        if (P->combine_mods) {
            int hp = 0, mp = 0;
            int has_h = (hi < F->mm_h.n) && (C == F->mm_h.v[hi]);
            int has_m = (mi < F->mm.n) && (C == F->mm.v[mi]);
            if (has_h) { hp = F->ml_h.v[hi]; hi++; }
            if (has_m) { mp = F->ml.v[mi]; mi++; }
            if (has_h || has_m) {
                int comb = hp + mp; if (comb > 255) comb = 255;
                if (comb > (255 * th)) cur = 'M'; else if (comb < (255 * (1 - th))) cur = 'U';
                mask[i] = cur;
            }
        } else {
            if ((hi < F->mm_h.n) && (C == F->mm_h.v[hi])) {
                if (F->ml_h.v[hi] > (255 * th)) cur = 'H'; else if (F->ml_h.v[hi] < (255 * (1 - th))) cur = 'U';
                mask[i] = cur; hi++;
            }
            if ((mi < F->mm.n) && (C == F->mm.v[mi])) {
                if (F->ml.v[mi] > (255 * th)) cur = 'M';
                else if (F->ml.v[mi] < (255 * (1 - th))) { if (cur != 'H') cur = 'U'; }
                else if (cur != 'H') cur = 'N';
                mask[i] = cur; mi++;
            }
        }
        C++;
    }
}

/* reverse_comp -- patter_utils.cpp:179-201; returns 0 on unsupported base */
static int reverse_comp(const char *s, long n, char *o) {
    for (long i = 0; i < n; i++) {
        char c = s[n - 1 - i], r;
        if (c == 'A') r = 'T'; else if (c == 'C') r = 'G'; else if (c == 'G') r = 'C'; else if (c == 'T') r = 'A';
        else if (c == 'N') r = 'N'; else return 0;
        o[i] = r;
    }
    return 1;
}

/* np_samLineToPatVec -- ont.cpp:90-221.  ret: 1 ok(pattern, *first), 0 empty, -1 invalid */
static int np_line_to_pat(patter_t *P, const dict_t *d, const tok *t, int nt, sbuf *pat, int *first) {
    npfields F; memset(&F, 0, sizeof F);
    int rc = 0; sbuf seq = {0}, m2 = {0}; char *orig = NULL, *mask = NULL;
    if (!parse_np_fields(P, t, nt, &F)) { rc = -1; goto out; }
    if ((F.mm.n == 0 && (F.mm_h.n == 0 && !F.np_dot)) || tok_eq(t[9], "*")) { P->nr_empty++; rc = 0; goto out; }
    long sl, fl; unsigned long usl;
    /* unsigned long start_locus = stoul(tokens[3]) (ont.cpp:102): kept 64-bit; arithmetic below is the same modulo 2^64 */
    if (!cxx_stoul(t[3].s, t[3].n, &usl) || !cxx_stoi(t[1].s, t[1].n, &fl)) { rc = -1; goto out; }
    sl = (long)usl;
    int bottom = ((fl & 0x10) == 16);
    if (!clean_cigar(t[9].s, t[9].n, t[5].s, t[5].n, &seq)) { rc = -1; goto out; }
    long n = t[9].n; orig = (char *)malloc(n + 1); mask = (char *)malloc(n + 1);
    if (bottom) { if (!reverse_comp(t[9].s, n, orig)) { rc = -1; goto out; } } else memcpy(orig, t[9].s, n);
    make_meth_mask(P, &F, orig, n, mask);
    if (bottom) for (long i = 0; i < n / 2; i++) { char c = mask[i]; mask[i] = mask[n - 1 - i]; mask[n - 1 - i] = c; }
    if (!clean_cigar(mask, n, t[5].s, t[5].n, &m2)) { rc = -1; goto out; }
    pat->n = 0; sb_put(pat, "", 0);
    int start_site = -1;
    for (long i = bottom ? -1 : 0; i < seq.n; i++) {
        long di = bottom ? i + 1 : i;
        if (sl + i > d->bsize - 1) continue;
        if (sl + i < 0) continue;
        int dj = dict_find(d, sl + i);
        if (dj < 0) continue;
        if (di >= m2.n) continue;
        char st;
        if (m2.p[di] == 'N') st = '.';
        else if (m2.p[di] == 'E') {
            int has_base = (di >= 0 && di < seq.n) && (bottom ? seq.p[di] == 'G' : seq.p[di] == 'C');
            st = (F.np_dot && has_base) ? 'T' : '.';
        } else {
            st = '.';
            if (m2.p[di] == 'M') st = 'C'; else if (m2.p[di] == 'U') st = 'T'; else if (m2.p[di] == 'H') st = 'H';
        }
        long clip_pos = bottom ? di : i;
        if (!((clip_pos >= P->clip) && (clip_pos < (long)seq.n - P->clip))) st = '.';
        if (start_site < 0 && st != '.') start_site = d->idx[dj];
        if (start_site > 0) sb_putc(pat, st);
    }
    if (strip_pat(pat)) { P->nr_empty++; rc = 0; goto out; }
    if (start_site < 1) { P->nr_empty++; rc = 0; goto out; }
    *first = start_site; rc = 1;
out:
    free(seq.p); free(m2.p); free(orig); free(mask);
    free(F.mm.v); free(F.ml.v); free(F.mm_h.v); free(F.ml_h.v);
    return rc;
}

/* a8 samLineToPatVec -- patter.cpp:187-245.  ret 1 ok, 0 empty(vector) */
typedef struct { int ok; int first; sbuf pat; tok chr; } patvec;
static void line_to_patvec(patter_t *P, const dict_t *d, const tok *t, int nt, patvec *o) {
    o->ok = 0; o->pat.n = 0;
    if (nt == 0) return;
    if (nt < 11) { P->nr_invalid++; return; }
    if (P->nanopore) {
        int f; int r = np_line_to_pat(P, d, t, nt, &o->pat, &f);
        if (r < 0) P->nr_invalid++;
        if (r == 1) { o->ok = 1; o->first = f; o->chr = t[2]; }
        return;
    }
    long sl, fl; unsigned long usl;
    /* stoul(tokens[3]) then passed as `int start_locus` to compareSeqToRef (patter.cpp:208,105): low 32 bits, signed */
    if (!cxx_stoul(t[3].s, t[3].n, &usl) || !cxx_stoi(t[1].s, t[1].n, &fl)) { P->nr_invalid++; return; }
    sl = (long)(int)(unsigned int)usl;
    sbuf adj = {0};
    if (!clean_cigar(t[9].s, t[9].n, t[5].s, t[5].n, &adj)) { P->nr_invalid++; free(adj.p); return; }
    int f = compare_seq_to_ref(P, d, adj.p, adj.n, sl, (int)fl, &o->pat);
    free(adj.p);
    if (f < 1) { P->nr_empty++; return; }
    o->ok = 1; o->first = f; o->chr = t[2];
}

/* a7 merge_PE -- patter_utils.cpp:292-342 + proc2lines patter.cpp:247-290. Appends output line. */
static void proc2(patter_t *P, const dict_t *d, const tok *t1, int n1, const tok *t2, int n2, sbuf *out) {
    patvec a = {0}, b = {0};
    line_to_patvec(P, d, t1, n1, &a);
    line_to_patvec(P, d, t2, n2, &b);
    patvec *l1 = &a, *l2 = &b, res = {0};
    sbuf merged = {0};
    if (!a.ok && !b.ok) goto out;
    if (!a.ok) { res = b; } else if (!b.ok) { res = a; }
    else {
        if (l1->first > l2->first) { patvec *tmp = l1; l1 = l2; l2 = tmp; }
        int s1 = l1->first, s2 = l2->first;
        long e1 = s1 + l1->pat.n, e2 = s2 + l2->pat.n; int last = (int)(e1 > e2 ? e1 : e2);
        if (last - s1 > 300) goto out;                   /* throws "merged read is too long": nothing printed */
        for (int i = s1; i < last; i++) sb_putc(&merged, '.');
        memcpy(merged.p, l1->pat.p, l1->pat.n);
        for (long i = 0; i < l2->pat.n; i++) {
            long ai = i + s2 - s1;
            if (merged.p[ai] == '.') merged.p[ai] = l2->pat.p[i];
            else if (l2->pat.p[i] != '.' && merged.p[ai] != l2->pat.p[i]) merged.p[ai] = '.';
        }
        int sp = strip_pat(&merged);
        if (sp < 0) goto out;                            /* strip_read clears -> res.empty() */
        res.ok = 1; res.first = s1 + sp; res.pat = merged; res.chr = l1->chr;
    }
    if (res.pat.n < P->min_cpg) { P->nr_short++; goto out; }
    {
        char num[32];
        sb_put(out, res.chr.s, res.chr.n); sb_putc(out, '\t');
        int l = snprintf(num, sizeof num, "%d", res.first); sb_put(out, num, l); sb_putc(out, '\t');
        sb_put(out, res.pat.p, res.pat.n); sb_putc(out, '\n');
    }
out:
    free(a.pat.p); free(b.pat.p); free(merged.p);
}

/*
 * port_patter: parse_reads_from_stdin -- patter.cpp:381-416 (+ first_line :324-350).
 * sam: SAM text (no header), mates adjacent (i.e. after match_maker).  loci/idx: dictionary of REGION.
 * stats out: [lines, pairs, empty, short, invalid, is_pe]
 * returns bytes written to *out (malloc'ed, caller frees with port_free), or -1 on fatal (PE+nanopore).
 */
long port_patter(const char *sam, long n, const int *loci, const int *idx, int ncpg, int min_cpg, int clip,
                 int nanopore, float np_thresh, char cpc_call, int combine_mods, char **outp, long *stats) {
    patter_t P; memset(&P, 0, sizeof P);
    P.min_cpg = min_cpg; P.clip = clip; P.nanopore = nanopore; P.np_thresh = np_thresh; P.cpc_call = cpc_call;
    P.combine_mods = combine_mods;
    dict_t d = { loci, idx, ncpg, ncpg ? (long)loci[ncpg - 1] + 1 : 0 };
    sbuf out = {0}; sb_put(&out, "", 0);
    enum { MAXT = 64 };
    tok t1[MAXT], t2[MAXT]; int n1 = 0, n2 = 0, init = 0;
    long st = 0;
    for (long i = 0; i <= n; i++) {
        if (i < n && sam[i] != '\n') continue;
        if (i == n && st == n) break;
        const char *line = sam + st; long ln = i - st; st = i + 1;
        if (ln == 0) { P.line_i++; continue; }
        if (!init) {
            tok tt[MAXT]; int k = split_tabs(line, ln, tt, MAXT); if (k > MAXT) k = MAXT;
            long fl = 0; if (k < 3 || !cxx_stoi(tt[1].s, tt[1].n, &fl)) { free(out.p); return -1; }
            P.is_pe = (int)(((uint16_t)fl) & 1);
            int has_np = 0;
            for (int j = 11; j < k; j++) if ((tok_pref(tt[j], "MM:Z:") || tok_pref(tt[j], "Mm:Z:")) && tt[j].n > 5) has_np = 1;
            P.nanopore = P.nanopore || has_np;
            if (P.is_pe && P.nanopore) { free(out.p); return -1; }
            init = 1;
        }
        if (n1 == 0) {
            n1 = split_tabs(line, ln, t1, MAXT); if (n1 > MAXT) n1 = MAXT;
            if (!P.is_pe) { proc2(&P, &d, t1, n1, NULL, 0, &out); n1 = 0; }
            P.line_i++; continue;
        }
        n2 = split_tabs(line, ln, t2, MAXT); if (n2 > MAXT) n2 = MAXT;
        if (n1 && n2 && t1[0].n == t2[0].n && !memcmp(t1[0].s, t2[0].s, t1[0].n)) {
            proc2(&P, &d, t1, n1, t2, n2, &out); P.nr_pairs++; n1 = 0;
        } else {
            proc2(&P, &d, t1, n1, NULL, 0, &out);
            memcpy(t1, t2, sizeof(tok) * n2); n1 = n2;
        }
        P.line_i++;
    }
    if (n1) proc2(&P, &d, t1, n1, NULL, 0, &out);
    stats[0] = P.line_i; stats[1] = P.nr_pairs; stats[2] = P.nr_empty; stats[3] = P.nr_short; stats[4] = P.nr_invalid;
    stats[5] = P.is_pe;
    *outp = out.p; return out.n;
}
void port_free(void *p) { free(p); }

/* ------------------------------------------------------------------------------------------------
 * a1 match_maker -- pipeline_wgbs/match_maker.cpp:48-183 (output_singles = true)
 * ---------------------------------------------------------------------------------------------- */
typedef struct { const char *s; long n; } lref;
static int lref_cmp(const void *a, const void *b) {      /* std::string operator< */
    const lref *x = (const lref *)a, *y = (const lref *)b; long m = x->n < y->n ? x->n : y->n;
    int c = memcmp(x->s, y->s, m); if (c) return c; return (x->n > y->n) - (x->n < y->n);
}
static long field_int(lref l, int f) {
    tok t[16]; int k = split_tabs(l.s, l.n, t, 16); long v = 0; if (f < k && f < 16) cxx_stoi(t[f].s, t[f].n, &v); return v;
}
static tok field_tok(lref l, int f) { tok t[16]; tok e = {"", 0}; int k = split_tabs(l.s, l.n, t, 16); return (f < k && f < 16) ? t[f] : e; }
typedef struct { long key; lref r1, r2; long ord; } pe_t;
static int pe_cmp(const void *a, const void *b) {        /* key only; ties broken by discovery order (std::sort is unstable: ties unspecified) */
    const pe_t *x = (const pe_t *)a, *y = (const pe_t *)b;
    if (x->key != y->key) return (x->key > y->key) - (x->key < y->key);
    return (x->ord > y->ord) - (x->ord < y->ord);
}
static long mm_flush(lref *data, long nd, int last_chunk, sbuf *out, lref *opt) {
    if (!nd) return 0;
    tok lastchrom = field_tok(data[nd - 1], 2);
    qsort(data, nd, sizeof(lref), lref_cmp);
    char *fl = (char *)calloc(nd, 1); pe_t *pv = (pe_t *)malloc(sizeof(pe_t) * nd); long np = 0;
    long last_pos = field_int(data[nd - 1], 3);
    for (long i = 0; i + 1 < nd; i++) {
        tok a = field_tok(data[i], 0), b = field_tok(data[i + 1], 0);
        if (a.n == b.n && !memcmp(a.s, b.s, a.n)) {
            pe_t p; p.r1 = data[i]; p.r2 = data[i + 1]; p.key = field_int(data[i], 3); p.ord = np;
            long k2 = field_int(data[i], 7);
            if (k2 < p.key) { p.key = k2; p.r1 = data[i + 1]; p.r2 = data[i]; }
            pv[np++] = p; fl[i] = fl[i + 1] = 1; i++;
        } else if (field_int(data[i], 7) < last_pos || last_chunk) {
            pe_t p; p.r1 = data[i]; p.r2.s = NULL; p.r2.n = 0; p.key = field_int(data[i], 3); p.ord = np; pv[np++] = p; fl[i] = 1;
        }
    }
    long no = 0;
    for (long i = 0; i < nd; i++) if (!fl[i]) {
        if (last_chunk) { pe_t p; p.r1 = data[i]; p.r2.s = NULL; p.r2.n = 0; p.key = field_int(data[i], 3); p.ord = np; pv[np++] = p; }
        else { tok c = field_tok(data[i], 2); if (c.n == lastchrom.n && !memcmp(c.s, lastchrom.s, c.n)) opt[no++] = data[i]; }
    }
    qsort(pv, np, sizeof(pe_t), pe_cmp);
    for (long i = 0; i < np; i++) {
        sb_put(out, pv[i].r1.s, pv[i].r1.n); sb_putc(out, '\n');
        if (pv[i].r2.s) { sb_put(out, pv[i].r2.s, pv[i].r2.n); sb_putc(out, '\n'); }
    }
    free(fl); free(pv);
    return no;
}
long port_match_maker(const char *sam, long n, char **outp) {
    sbuf out = {0}; sb_put(&out, "", 0);
    long cap = 1 << 16, nd = 0; lref *data = (lref *)malloc(sizeof(lref) * cap), *opt = (lref *)malloc(sizeof(lref) * cap);
    long line_i = 0, st = 0; int chrom_set = 0;
    for (long i = 0; i <= n; i++) {
        if (i < n && sam[i] != '\n') continue;
        if (i == n && st == n) break;
        lref l = { sam + st, i - st }; st = i + 1;
        if (!chrom_set && l.n && l.s[0] == '@') { sb_put(&out, l.s, l.n); sb_putc(&out, '\n'); line_i++; continue; }
        chrom_set = 1;
        if (nd + 1 >= cap) { cap *= 2; data = (lref *)realloc(data, sizeof(lref) * cap); opt = (lref *)realloc(opt, sizeof(lref) * cap); }
        data[nd++] = l;
        tok c = field_tok(l, 2), c0 = field_tok(data[0], 2);
        if (line_i && !(c.n == c0.n && !memcmp(c.s, c0.s, c.n))) { nd = mm_flush(data, nd, 1, &out, opt); memcpy(data, opt, sizeof(lref) * nd); }
        if (line_i && (line_i % 50000 == 0) && nd) {
            if (field_int(l, 3) - field_int(data[0], 3) > 160) { nd = mm_flush(data, nd, 0, &out, opt); memcpy(data, opt, sizeof(lref) * nd); }
        }
        line_i++;
    }
    mm_flush(data, nd, 1, &out, opt);
    free(data); free(opt);
    *outp = out.p; return out.n;
}

/* ------------------------------------------------------------------------------------------------
 * a12 collapse -- python/bam2pat.py:99-106:  sort -k2,2n -k3,3 | uniq -c | awk '{print $2,$3,$4,$1}'  (C locale)
 * input: lines "chr\tidx\tpattern\n"
 * ---------------------------------------------------------------------------------------------- */
static int pat_line_cmp(const void *a, const void *b) {
    const lref *x = (const lref *)a, *y = (const lref *)b;
    tok tx[4], ty[4]; split_tabs(x->s, x->n, tx, 4); split_tabs(y->s, y->n, ty, 4);
    long ix = 0, iy = 0; cxx_stoi(tx[1].s, tx[1].n, &ix); cxx_stoi(ty[1].s, ty[1].n, &iy);
    if (ix != iy) return (ix > iy) - (ix < iy);
    lref px = { tx[2].s, tx[2].n }, py = { ty[2].s, ty[2].n };
    int c = lref_cmp(&px, &py); if (c) return c;
    return lref_cmp(x, y);                                 /* last-resort whole-line comparison */
}
long port_collapse(const char *txt, long n, char **outp) {
    long cap = 1 << 16, nl = 0; lref *L = (lref *)malloc(sizeof(lref) * cap); long st = 0;
    for (long i = 0; i <= n; i++) {
        if (i < n && txt[i] != '\n') continue;
        if (i == n && st == n) break;
        if (nl == cap) { cap *= 2; L = (lref *)realloc(L, sizeof(lref) * cap); }
        L[nl].s = txt + st; L[nl].n = i - st; nl++; st = i + 1;
    }
    qsort(L, nl, sizeof(lref), pat_line_cmp);
    sbuf out = {0}; sb_put(&out, "", 0); char num[32];
    for (long i = 0; i < nl;) {
        long j = i + 1; while (j < nl && !lref_cmp(&L[i], &L[j])) j++;
        sb_put(&out, L[i].s, L[i].n); sb_putc(&out, '\t');
        int l = snprintf(num, sizeof num, "%ld", j - i); sb_put(&out, num, l); sb_putc(&out, '\n');
        i = j;
    }
    free(L); *outp = out.p; return out.n;
}

/* ------------------------------------------------------------------------------------------------
 * pat text reader shared by pat2beta / homog ports: "chr\tidx\tpattern\tcount[\t...]"
 * ---------------------------------------------------------------------------------------------- */
typedef int (*patline_fn)(void *u, long site, const char *pat, long plen, long count);
static int for_pat_lines(const char *txt, long n, patline_fn fn, void *u) {   /* ret -1 on "throw" */
    long st = 0;
    for (long i = 0; i <= n; i++) {
        if (i < n && txt[i] != '\n') continue;
        if (i == n && st == n) break;
        const char *line = txt + st; long ln = i - st; st = i + 1;
        if (!ln) continue;
        tok t[8]; int k = split_tabs(line, ln, t, 8);
        if (k < 4) return -1;
        long site, cnt; if (!cxx_stoi(t[1].s, t[1].n, &site) || !cxx_stoi(t[3].s, t[3].n, &cnt)) return -1;
        int r = fn(u, site, t[2].s, t[2].n, cnt); if (r) return r > 0 ? 0 : -1;
    }
    return 0;
}

/* a13 stdin2beta -- pat2beta/stdin2beta.cpp:59-93 */
typedef struct { int start, end, n; int32_t *mc; } p2b_t;
static int p2b_line(void *u, long site, const char *pat, long plen, long count) {
    p2b_t *b = (p2b_t *)u;
    if ((site + plen - 1 < b->start) || (site >= b->end)) return 0;
    for (long i = 0; i < plen; i++) {
        long k = site - b->start + i; char c = pat[i];
        if (k >= b->n || k < 0) continue;
        if (!(c == 'T' || c == 'C' || c == 'H')) continue;
        b->mc[2 * k + 1] += (int32_t)count;
        if (c == 'C' || c == 'H') b->mc[2 * k] += (int32_t)count;
    }
    return 0;
}
/* meth_cov: int32[n,2] (meth, cover), zero-initialised here. returns 0 ok, -1 parse failure ("failed calculating beta") */
int port_pat2beta(const char *txt, long n, int start, int end, int32_t *meth_cov) {
    p2b_t b = { start, end, end - start, meth_cov };
    memset(meth_cov, 0, sizeof(int32_t) * 2 * (size_t)(end - start));
    return for_pat_lines(txt, n, p2b_line, &b);
}

/* a14 trim_to_uint8 -- python/utils_wgbs.py:277-290 (float64 divide then multiply then truncate) */
void port_trim(const int64_t *mc, long n, int nbits, void *out) {
    int64_t maxv = (1 << nbits) - 1;
    for (long i = 0; i < n; i++) {
        int64_t m = mc[2 * i], c = mc[2 * i + 1];
        if (c > maxv) { m = (int64_t)(((double)m / (double)c) * (double)maxv); c = maxv; }
        if (nbits == 8) { ((uint8_t *)out)[2 * i] = (uint8_t)m; ((uint8_t *)out)[2 * i + 1] = (uint8_t)c; }
        else { ((uint16_t *)out)[2 * i] = (uint16_t)m; ((uint16_t *)out)[2 * i + 1] = (uint16_t)c; }
    }
}

/* a15 homog -- homog/homog.cpp:154-260 (blocks already loaded/sorted/filtered; see test harness) */
typedef struct { const int *bs, *be; long nb; const float *range; int nbins, min_cpgs, inclusive; int32_t *counts; long cur; } hg_t;
static void hg_update(hg_t *h, long bi, const char *pat, long off, long len, long count) {
    long nC = 0, nT = 0;
    for (long i = off; i < off + len; i++) { if (pat[i] == 'C' || pat[i] == 'H') nC++; else if (pat[i] == 'T') nT++; }
    if (nC + nT < h->min_cpgs) return;
    float m = (float)nC / (float)(nC + nT);
    if (m < h->range[0]) return;
    int b; for (b = 0; b < h->nbins; b++) if (m >= h->range[b] && m < h->range[b + 1]) break;
    if (b == h->nbins) b--;
    h->counts[bi * h->nbins + b] += (int32_t)count;
}
static int hg_line(void *u, long rs, const char *pat, long plen, long count) {
    hg_t *h = (hg_t *)u;
    long re = rs + plen - 1;
    if (h->nb && rs >= h->be[h->nb - 1]) return 1;        /* homog.cpp:218-223 (pipe exhausted whenever this holds) */
    while (h->cur < h->nb && rs >= h->be[h->cur]) h->cur++;
    if (h->cur >= h->nb) return 1;
    if (re < h->bs[h->cur]) return 0;
    for (long bi = h->cur; bi < h->nb; bi++) {
        if (h->bs[bi] > re) break;
        long os = rs > h->bs[bi] ? rs : h->bs[bi];
        long oe = (rs + plen) < h->be[bi] ? (rs + plen) : h->be[bi];
        if (os >= oe) continue;
        if (h->inclusive) { if (plen < h->min_cpgs) continue; hg_update(h, bi, pat, 0, plen, count); }
        else { if (oe - os < h->min_cpgs) continue; hg_update(h, bi, pat, os - rs, oe - os, count); }
    }
    return 0;
}
int port_homog(const char *txt, long n, const int *bstart, const int *bend, long nblocks, const float *range, int nbins,
               int min_cpgs, int inclusive, int32_t *counts) {
    hg_t h = { bstart, bend, nblocks, range, nbins, min_cpgs, inclusive, counts, 0 };
    memset(counts, 0, sizeof(int32_t) * (size_t)nblocks * nbins);
    return for_pat_lines(txt, n, hg_line, &h);
}

/* a17 segmentor::dp + traceback -- segment_betas/segmentor.cpp:50-159.
 * betas: K pointers to uint8[n,2] (already offset to the chunk start).  borders out (ascending), returns count.
 * NB: the float/double mix is the reference's; this file is compiled with -ffp-contract=off. */
int port_segment(const uint8_t *const *betas, int K, const uint32_t *dists, int n, int max_cpg, uint32_t max_bp,
                 float pseudo, int32_t *borders) {
    float *nm = (float *)malloc(sizeof(float) * K), *nt = (float *)malloc(sizeof(float) * K);
    int ring = 1; while (ring < max_cpg) ring <<= 1; int mask = ring - 1;
    double *mem = (double *)calloc((size_t)ring * max_cpg, sizeof(double));
    double *M = (double *)calloc(n + 1, sizeof(double)); int *T = (int *)calloc(n + 1, sizeof(int));
    const double NEG = -(double)INFINITY;
    for (int i = 0; i < n; i++) {
        double *row = &mem[(size_t)(i & mask) * max_cpg];
        for (int j = 0; j < max_cpg; j++) row[j] = 0.0;
        memset(nm, 0, sizeof(float) * K); memset(nt, 0, sizeof(float) * K);
        int window = (n - i) < max_cpg ? (n - i) : max_cpg;
        for (int j = 0; j < window; j++) {
            if ((dists[i + j] - dists[i] > max_bp) || (dists[i + j] < dists[i])) { row[j] = NEG; continue; }
            double ll_sum = 0;
            for (int k = 0; k < K; k++) {
                nm[k] += (float)betas[k][(i + j) * 2]; nt[k] += (float)betas[k][(i + j) * 2 + 1];
                float ntk = nt[k], nmk = nm[k];
                if (!ntk) continue;
                float p = (nmk + pseudo) / (ntk + (2 * pseudo));
                float ll = 0;
                if (p > 0.0) ll += (nmk * log2f(p));
                if (p < 1.0) ll += (ntk - nmk) * log2(1.0 - p);
                ll_sum += ll;
            }
            if (ll_sum) row[j] = ll_sum;
        }
        double best = NEG; int bi = -1; int sk = (i + 1 - max_cpg) > 0 ? (i + 1 - max_cpg) : 0;
        for (int k = sk; k < i + 1; k++) {
            double tmp = M[k] + mem[(size_t)(k & mask) * max_cpg + (i - k)];
            if (tmp > best) { best = tmp; bi = k; }
        }
        M[i + 1] = best; T[i + 1] = bi;
    }
    int nb = 0; int i = n; borders[nb++] = i;
    while (i > 0) { i = T[i] > 0 ? T[i] : 0; borders[nb++] = i; }
    for (int a = 0, b = nb - 1; a < b; a++, b--) { int t = borders[a]; borders[a] = borders[b]; borders[b] = t; }
    free(nm); free(nt); free(mem); free(M); free(T);
    return nb;
}

/* host libm probes: used by the exhaustive GPU-vs-glibc log2f/log2 sweep */
void port_log2f_array(const float *x, long n, float *y) { for (long i = 0; i < n; i++) y[i] = log2f(x[i]); }
void port_log2_1m_array(const float *p, long n, double *y) { for (long i = 0; i < n; i++) y[i] = log2(1.0 - (double)p[i]); }
