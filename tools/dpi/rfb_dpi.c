/*
 * rfb_dpi.c -- DPI-C shim: lets the reference's own SystemVerilog testbench (Simulation/testbench_BLK_Mem.sv) keep
 * its file reads, counters and report while the instantiation of CSR_traversal + design_1_wrapper (TB:89-106) is
 * replaced by one imported call.  The SystemVerilog side is tools/dpi/tb_dpi.sv.
 *
 * Only fixed-size unpacked arrays of 2-state bytes / ints cross the boundary, which IEEE 1800 (35.5.6) maps to plain
 * C pointers -- no svdpi.h, no simulator headers: this file compiles with any C compiler against
 * include/regex_fpga_b200.h and links to librfb200.so.
 *
 *   rfb_dpi_open  (coe_path, size_range)            BRAM initialisation + `size` port (Design/top.v:10-13, TB:20)
 *   rfb_dpi_scan2 (lo, hi, trace_entries, ...)      the clk / input_char(_2) / input_char_flag loop of TB:49-73 for the
 *                                                   two lock-stepped streams; returns the accept pulses (TB:61-69)
 *   rfb_dpi_cycles(lo, hi, trace_entries, cycles)   the testbench's "Total no. cycles" (TB:52,84)
 *   rfb_dpi_close ()
 */
#include "regex_fpga_b200.h"
#include <stdlib.h>
#include <string.h>

static rfb_ctx *g_ctx;
static rfb_nfa *g_nfa;

const char *rfb_dpi_error(void) { return rfb_last_error(g_ctx); }

int rfb_dpi_open(const char *coe_path, long long size_range) {
    int rc;
    if (g_ctx) return RFB_E_INVALID;
    if ((rc = rfb_ctx_create(0, &g_ctx)) != RFB_OK) return rc;
    if ((rc = rfb_nfa_load_coe(g_ctx, coe_path, size_range, &g_nfa)) != RFB_OK) { rfb_ctx_destroy(g_ctx); g_ctx = NULL; return rc; }
    return RFB_OK;
}

void rfb_dpi_close(void) {
    rfb_nfa_destroy(g_nfa); g_nfa = NULL;
    rfb_ctx_destroy(g_ctx); g_ctx = NULL;
}

/* lo feeds input_char / match_count (stream 0), hi feeds input_char_2 / match_count_2 (stream 1): TB:56-57,61-69.
 * An M-entry trace pair executes M - 1 symbol steps (TB:71-86).  Records come back in canonical
 * (stream, pos, state) order; *n_records is the number of pulses, of which min(*n_records, capacity) are stored. */
int rfb_dpi_scan2(const unsigned char *lo, const unsigned char *hi, int trace_entries, int capacity,
                  unsigned int *n_records, unsigned int *rec_stream, unsigned int *rec_pos, unsigned int *rec_state) {
    if (!g_ctx || !g_nfa || !lo || !hi || trace_entries < 1 || capacity < 0 || !n_records) return RFB_E_INVALID;
    const size_t M = (size_t)trace_entries;
    unsigned char *buf = (unsigned char *)malloc(2 * M);
    rfb_match *recs = (rfb_match *)malloc(((size_t)capacity + 1) * sizeof(rfb_match));
    if (!buf || !recs) { free(buf); free(recs); return RFB_E_NOMEM; }
    memcpy(buf, lo, M);
    memcpy(buf + M, hi, M);
    rfb_batch b;
    memset(&b, 0, sizeof b);
    b.data = buf; b.data_bytes = 2 * M; b.n_streams = 2; b.stride = M; b.n_steps = rfb_tb_steps((uint32_t)trace_entries);
    rfb_result r;
    memset(&r, 0, sizeof r);
    r.records = recs; r.record_capacity = (uint64_t)capacity;
    const int rc = rfb_scan(g_ctx, g_nfa, &b, RFB_SCAN_SORT_RECORDS | RFB_SCAN_NO_COUNTS, &r);
    if (rc == RFB_OK) {
        *n_records = (unsigned int)r.n_matches;
        for (uint64_t k = 0; k < r.n_records; k++) {
            if (rec_stream) rec_stream[k] = recs[k].stream;
            if (rec_pos) rec_pos[k] = recs[k].pos;
            if (rec_state) rec_state[k] = recs[k].state;
        }
    }
    free(buf); free(recs);
    return rc;
}

int rfb_dpi_cycles(const unsigned char *lo, const unsigned char *hi, int trace_entries, unsigned long long *cycles) {
    if (!g_ctx || !g_nfa || !cycles || trace_entries < 1) return RFB_E_INVALID;
    uint64_t c = 0;
    const int rc = rfb_fpga_cycles(g_ctx, g_nfa, lo, hi, (uint32_t)trace_entries, &c);
    *cycles = c;
    return rc;
}
