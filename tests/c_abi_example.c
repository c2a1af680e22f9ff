/* c_abi_example.c -- the C host program of INTEGRATION.md, compiled by tests/test_c_abi.py with plain gcc
 * against include/regex_fpga_b200.h and regex_fpga_b200/lib/librfb200.so (no Python, no torch in the process).
 * usage: c_abi_example <coe> <lo.mem> <hi.mem> <trace_entries>      prints the testbench's report (TB:75-84). */
#include "regex_fpga_b200.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int main(int argc, char **argv) {
    if (argc < 5) { fprintf(stderr, "usage\n"); return 2; }
    uint32_t *entries; size_t n_entries;
    if (rfb_coe_parse(argv[1], &entries, &n_entries)) { fprintf(stderr, "%s\n", rfb_last_error(NULL)); return 1; }
    printf("size_range = %lld\n", (long long)rfb_coe_detect_size(entries, n_entries));
    rfb_free(entries);
    if (argc > 5 && !strcmp(argv[5], "--host-only")) return 0;      /* formats only: no GPU needed */

    rfb_ctx *ctx; rfb_nfa *nfa;
    if (rfb_ctx_create(0, &ctx)) { fprintf(stderr, "%s\n", rfb_last_error(NULL)); return 1; }
    if (rfb_nfa_load_coe(ctx, argv[1], -1, &nfa)) { fprintf(stderr, "%s\n", rfb_last_error(ctx)); return 1; }
    uint8_t *lo, *hi; size_t nlo, nhi;
    if (rfb_trace_load_mem(argv[2], &lo, &nlo) || rfb_trace_load_mem(argv[3], &hi, &nhi)) { fprintf(stderr, "%s\n", rfb_last_error(NULL)); return 1; }
    const uint32_t M = (uint32_t)atoi(argv[4]);                      /* TB:71 uses 200000 */
    if (nlo < M || nhi < M) { fprintf(stderr, "traces too short\n"); return 1; }
    uint8_t *buf = malloc(2 * (size_t)M);
    memcpy(buf, lo, M); memcpy(buf + M, hi, M);                       /* stream 0 = lo = input_char, 1 = hi (TB:56-57) */
    rfb_nfa_info info; rfb_nfa_get_info(nfa, &info);
    rfb_batch b; memset(&b, 0, sizeof b);
    b.data = buf; b.data_bytes = 2 * (uint64_t)M; b.n_streams = 2; b.stride = M; b.n_steps = rfb_tb_steps(M);
    rfb_result r; memset(&r, 0, sizeof r);
    r.counts = calloc(info.n_states, 8);
    r.record_capacity = 1 << 20; r.records = malloc(r.record_capacity * sizeof(rfb_match));
    if (rfb_scan(ctx, nfa, &b, RFB_SCAN_SORT_RECORDS, &r)) { fprintf(stderr, "%s\n", rfb_last_error(ctx)); return 1; }
    for (int s = 0; s < 2; s++) {                                     /* TB:75-81 */
        unsigned *mc = calloc(info.n_states, sizeof *mc);
        for (uint64_t k = 0; k < r.n_records; k++) if (r.records[k].stream == (uint32_t)s) mc[r.records[k].state]++;
        for (int p = (int)info.n_states - 1; p >= 0; p--)
            if (mc[p]) printf(s ? "match_count_2[%d] = %u\n" : "match_count[%d] = %u\n", p, mc[p] & 0x3FF);
        free(mc);
    }
    /* the same run through a one-process group of GPUs (here: one): shards + NCCL all-reduce of the counts */
    {
        const int dev = 0;
        rfb_group *grp; rfb_group_nfa *gnfa;
        if (rfb_group_create(&dev, 1, &grp)) { fprintf(stderr, "%s\n", rfb_last_error(NULL)); return 1; }
        if (rfb_group_nfa_load_coe(grp, argv[1], -1, &gnfa)) { fprintf(stderr, "%s\n", rfb_group_last_error(grp)); return 1; }
        rfb_result g; memset(&g, 0, sizeof g);
        g.counts = calloc(info.n_states, 8);
        g.record_capacity = 1 << 20; g.records = malloc(g.record_capacity * sizeof(rfb_match));
        if (rfb_group_scan(grp, gnfa, &b, RFB_SCAN_SORT_RECORDS, &g)) { fprintf(stderr, "%s\n", rfb_group_last_error(grp)); return 1; }
        if (g.n_records != r.n_records || memcmp(g.records, r.records, r.n_records * sizeof(rfb_match)) ||
            memcmp(g.counts, r.counts, info.n_states * 8)) { fprintf(stderr, "group scan differs\n"); return 1; }
        free(g.counts); free(g.records);
        rfb_group_nfa_destroy(gnfa); rfb_group_destroy(grp);
    }
    uint64_t cycles = 0;
    if (rfb_fpga_cycles(ctx, nfa, lo, hi, M, &cycles)) { fprintf(stderr, "%s\n", rfb_last_error(ctx)); return 1; }
    printf("Total no. cycles: %llu\n", (unsigned long long)cycles);   /* TB:84 */
    rfb_free(lo); rfb_free(hi); free(buf); free(r.counts); free(r.records);
    rfb_nfa_destroy(nfa); rfb_ctx_destroy(ctx);
    return 0;
}
