/*
 * formats.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 * Decoders for the reference's two on-disk formats, written independently of the product's
 * parser (regex_fpga_b200/csrc/formats.cpp) so that the two can be cross-checked.
 */
#include "oracle.h"
#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static char *slurp(const char *path, size_t *len) {
    FILE *f = fopen(path, "rb");
    if (!f) return NULL;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    char *buf = (char *)malloc((size_t)n + 1);
    if (!buf) { fclose(f); return NULL; }
    if (fread(buf, 1, (size_t)n, f) != (size_t)n) { fclose(f); free(buf); return NULL; }
    fclose(f);
    buf[n] = 0;
    *len = (size_t)n;
    return buf;
}

static int hexval(int c) {
    if (c >= '0' && c <= '9') return c - '0';
    if (c >= 'a' && c <= 'f') return c - 'a' + 10;
    if (c >= 'A' && c <= 'F') return c - 'A' + 10;
    return -1;
}

/* Block_Mem/CSR_BlockMem_snort_16.coe:1-2 / CSR_BlockMem.coe:1-2:
 *   memory_initialization_radix=16;
 *   memory_initialization_vector=<32 hex digits><sep>... [;]
 * One 32-digit word = one 128-bit BRAM line; the leftmost 8 digits are rd_bus[127:96] = cache[0]
 * (Design/FPGA.v:884), i.e. entry 4*line+0.  Separators: newline (snort_16), blank (l7), commas
 * permitted by the COE grammar.  The closing ';' is optional (snort_16 has none). */
int orc_coe_parse(const char *path, uint32_t **entries, size_t *n_entries) {
    size_t len;
    char *txt = slurp(path, &len);
    if (!txt) return -1;
    const char *key = "memory_initialization_vector";
    char *p = strstr(txt, key);
    if (!p) { free(txt); return -2; }
    /* radix must be 16 */
    char *r = strstr(txt, "memory_initialization_radix");
    if (r) {
        r = strchr(r, '=');
        if (!r || strtol(r + 1, NULL, 10) != 16) { free(txt); return -3; }
    }
    p = strchr(p, '=');
    if (!p) { free(txt); return -2; }
    p++;
    size_t cap = len / 8 + 4, n = 0;
    uint32_t *E = (uint32_t *)malloc(cap * sizeof(uint32_t));
    while (*p) {
        while (*p && (isspace((unsigned char)*p) || *p == ',')) p++;
        if (!*p || *p == ';') break;
        /* one word: exactly 32 hex digits */
        uint32_t w[4] = {0, 0, 0, 0};
        int d = 0;
        while (hexval((unsigned char)*p) >= 0) {
            if (d >= 32) { free(E); free(txt); return -4; }
            w[d >> 3] = (w[d >> 3] << 4) | (uint32_t)hexval((unsigned char)*p);
            d++; p++;
        }
        if (d != 32) { free(E); free(txt); return -4; }
        for (int k = 0; k < 4; k++) E[n++] = w[k];
    }
    free(txt);
    *entries = E;
    *n_entries = n;
    return 0;
}

/* The image stores row_ptr[0..size] then nnz transitions, zero-padded to a whole line
 * (Design/FPGA.v:773,782,793): size is the unique value with E[0]==0, E[0..size] non-decreasing,
 * pad = n - (size+1+E[size]) in 0..3 and all pad entries zero. */
int64_t orc_detect_size(const uint32_t *E, size_t n) {
    if (n == 0 || E[0] != 0) return -1;
    int64_t found = -1;
    for (size_t size = 1; size < n; size++) {
        if (E[size] < E[size - 1]) break;
        uint64_t used = (uint64_t)size + 1 + E[size];
        if (used > n || n - used > 3) continue;
        int ok = 1;
        for (size_t j = used; j < n; j++) if (E[j]) ok = 0;
        if (!ok) continue;
        if (found >= 0) return -1; /* ambiguous */
        found = (int64_t)size;
    }
    return found;
}

/* $readmemh text without address directives: token k -> array index k, lowest address first
 * (testbench_BLK_Mem.sv:16-17,34-35).  Tokens are 1-2 hex digits. */
int orc_mem_parse(const char *path, uint8_t **bytes, size_t *n_out) {
    size_t len;
    char *txt = slurp(path, &len);
    if (!txt) return -1;
    uint8_t *b = (uint8_t *)malloc(len / 2 + 2);
    size_t n = 0;
    const char *p = txt;
    while (*p) {
        while (*p && isspace((unsigned char)*p)) p++;
        if (!*p) break;
        int v = 0, d = 0;
        while (hexval((unsigned char)*p) >= 0) { v = v * 16 + hexval((unsigned char)*p); d++; p++; }
        if (d == 0 || d > 2 || (*p && !isspace((unsigned char)*p))) { free(b); free(txt); return -4; }
        b[n++] = (uint8_t)v;
    }
    free(txt);
    *bytes = b;
    *n_out = n;
    return 0;
}

void orc_free(void *p) { free(p); }
