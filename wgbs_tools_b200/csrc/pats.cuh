// pats.cuh -- device representation of pat records (see include/wgbs_b200.h for the layout contract).
#pragma once
#include <stdint.h>

enum : uint32_t { SYM_DOT = 0, SYM_C = 1, SYM_H = 2, SYM_T = 3 };

struct wgbs_pats {
    size_t n = 0;            // records
    size_t pool_words = 0;
    uint32_t *idx = nullptr;    // CpG index of symbol 0 (int32 semantics)
    uint32_t *len = nullptr;    // symbols
    uint32_t *count = nullptr;  // multiplicity (int32 semantics)
    uint32_t *off = nullptr;    // n+1 word offsets into pool
    uint32_t *pool = nullptr;   // 16 two-bit symbols per word, first symbol in bits 31:30
    // --long only: read name of every record (patter --long prints it as a 4th column)
    uint32_t *name_off = nullptr, *name_len = nullptr;
    char *names = nullptr; size_t names_bytes = 0;
};

struct PatsView {
    size_t n;
    const uint32_t *idx, *len, *count, *off, *pool;
    const uint32_t *name_off, *name_len; const char *names;    // null unless --long
};
static inline PatsView view_of(const wgbs_pats *P) { return PatsView{P->n, P->idx, P->len, P->count, P->off, P->pool, P->name_off, P->name_len, P->names}; }

// ASCII -> 2-bit code; anything that is not C/H/T is "unknown" ('.'), which is how every consumer in the reference
// treats it (stdin2beta.cpp:82-84, homog.cpp:158-163).
__host__ __device__ __forceinline__ uint32_t sym_code(char c) { return c == 'C' ? SYM_C : c == 'T' ? SYM_T : c == 'H' ? SYM_H : SYM_DOT; }
__host__ __device__ __forceinline__ char sym_char(uint32_t code) { return code == SYM_C ? 'C' : code == SYM_T ? 'T' : code == SYM_H ? 'H' : '.'; }
