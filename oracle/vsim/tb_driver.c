/*
 * tb_driver.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Clocks the C model that v2c.py generates from the reference's unmodified Design/FPGA.v exactly as the
 * reference's testbench does (Simulation/testbench_BLK_Mem.sv; line numbers below are that file's), with a
 * behavioural stand-in for the one piece the reference does not ship: design_1_wrapper (TB:89-92,
 * Design/top.v:10-13), a Vivado block-memory ROM initialised from the .coe.  It is modelled as
 *     always @(posedge clk) dout <= mem[addr];
 * (no output register, always enabled, unwritten locations 0): the only read latency under which FPGA.v's own
 * timing works (address registered at edge k is consumed from rd_bus at edge k+2, FPGA.v:161 -> :182).
 *
 * Event order per clock period, from the testbench:
 *   posedge        CSR_traversal and the ROM update from pre-edge values (non-blocking)
 *   posedge + 1 ns cycles++ (TB:52); input_char_flag sampled (TB:53)
 *   posedge + 2 ns if the flag was set: input_char <= lo[m], input_char_2 <= hi[m], m++ (TB:55-58)
 *   same activation: match_count[i]++ / match_count_2[i]++ on the flags (TB:61-69, 10-bit counters TB:21-22);
 *                  m == 200000 ends the run (TB:71-86; the #20 that follows only delays the printout:
 *                  the always block is suspended during it, so nothing more is counted)
 * Reset: reset = 1 at t = 2, released at t = 12 together with size = size_range and the counter clear
 * (TB:31-45); the first posedge is at t = 10 (tb_clk starts at 1 and toggles every 5 ns, TB:11,26), so exactly
 * one edge sees reset = 1, with `size` still unassigned.
 */
#include "vsim_rt.h"
#include "../oracle.h"
#include <pthread.h>

#define ROM_LINES 65536u          /* rd_address is 16 bits wide: Design/FPGA.v:33 */

typedef struct {
    vs_model *m;
    uint64_t *rd_bus;             /* 512-bit input port; the testbench drives bits 127:0 (TB:12) */
    uint64_t *p_clk, *p_reset, *p_size, *p_c1, *p_c2;
    uint64_t *p_icf, *p_amf, *p_amf2, *p_i, *p_addr, *p_state;
    const uint32_t *E;
    size_t n_entries;
} tb_t;

static int tb_open(tb_t *t, const uint32_t *E, size_t n_entries, uint32_t size_range, int xfill) {
    const char *pn[1] = {"size_range"};
    const uint64_t pv[1] = {size_range};
    memset(t, 0, sizeof *t);
    t->m = vs_new(pn, pv, 1, xfill);                     /* CSR_traversal #(.size_range(size_range)) C1: TB:94 */
    if (!t->m) return -1;
    uint64_t nbits = 0;
    t->rd_bus = vs_get_wide(t->m, "rd_bus", &nbits);
    t->p_clk = vs_ptr(t->m, "clk"); t->p_reset = vs_ptr(t->m, "reset"); t->p_size = vs_ptr(t->m, "size");
    t->p_c1 = vs_ptr(t->m, "input_char"); t->p_c2 = vs_ptr(t->m, "input_char_2");
    t->p_icf = vs_ptr(t->m, "input_char_flag"); t->p_amf = vs_ptr(t->m, "accepting_match_flag");
    t->p_amf2 = vs_ptr(t->m, "accepting_match_flag_2"); t->p_i = vs_ptr(t->m, "i");
    t->p_addr = vs_ptr(t->m, "rd_address"); t->p_state = vs_ptr(t->m, "state");
    if (!t->rd_bus || nbits < 128 || !t->p_reset || !t->p_size || !t->p_c1 || !t->p_c2 || !t->p_icf || !t->p_amf ||
        !t->p_amf2 || !t->p_i || !t->p_addr || !t->p_state) { vs_delete(t->m); return -2; }
    t->E = E; t->n_entries = n_entries;
    return 0;
}

/* ROM word at `line` onto rd_bus[127:0]: entry 4*line + slot sits in bits 127-32*slot .. 96-32*slot
 * (the .coe prints each 128-bit word most significant digit first; Design/FPGA.v:881-884 unpacks it). */
static inline void rom_drive(tb_t *t, uint64_t line) {
    uint32_t e[4];
    for (int k = 0; k < 4; k++) {
        const size_t idx = (size_t)line * 4 + (size_t)k;
        e[k] = (line < ROM_LINES && idx < t->n_entries) ? t->E[idx] : 0u;
    }
    t->rd_bus[0] = (uint64_t)e[3] | ((uint64_t)e[2] << 32);
    t->rd_bus[1] = (uint64_t)e[1] | ((uint64_t)e[0] << 32);
}

/* one posedge of tb_clk: the DUT evaluates with the bus value the ROM registered at the PREVIOUS edge; the ROM
 * registers mem[rd_address as it was before this edge] */
static inline void tb_posedge(tb_t *t) {
    const uint64_t addr_pre = *t->p_addr;
    vs_posedge(t->m);
    rom_drive(t, addr_pre);
}

int ref_tb_run(const uint32_t *E, size_t n_entries, uint32_t size_range, const uint8_t *lo, const uint8_t *hi,
               uint64_t M, int xfill, uint16_t *mc1, uint16_t *mc2, uint64_t *cnt1, uint64_t *cnt2, orc_rec *recs,
               uint64_t cap, uint64_t *n_recs, uint64_t *cycles_out, uint64_t *trace, uint64_t trace_cap) {
    if (size_range == 0 || M == 0) return -1;
    tb_t t;
    int rc = tb_open(&t, E, n_entries, size_range, xfill);
    if (rc) return rc;
    if (t.p_clk) *t.p_clk = 1;
    uint64_t cycles = 0, m = 0, nr = 0;                  /* int m = 0; int cycles = 0;  TB:18-19 */
    *t.p_reset = 1;                                      /* t = 2: TB:31-32 */
    int released = 0;
    for (;;) {
        tb_posedge(&t);                                  /* always #5 tb_clk = ~tb_clk: TB:26 */
        cycles++;                                        /* TB:52 */
        if (trace && cycles <= trace_cap)
            trace[cycles - 1] = *t.p_i | (*t.p_icf << 20) | (*t.p_amf << 21) | (*t.p_amf2 << 22) | (*t.p_state << 24) | (*t.p_addr << 32);
        if (*t.p_icf == 1) {                             /* TB:53-59 */
            *t.p_c1 = lo[m];
            *t.p_c2 = hi[m];
            m++;
        }
        if (*t.p_amf == 1) {                             /* TB:61-64 */
            const uint64_t i = *t.p_i;
            if (i < size_range) {
                if (mc1) mc1[i] = (uint16_t)((mc1[i] + 1u) & 0x3FFu);
                if (cnt1) cnt1[i]++;
            }
            if (recs && nr < cap) { recs[nr].stream = 0; recs[nr].pos = (uint32_t)(m - 1); recs[nr].state = (uint32_t)i; }
            nr++;
        }
        if (*t.p_amf2 == 1) {                            /* TB:66-69 */
            const uint64_t i = *t.p_i;
            if (i < size_range) {
                if (mc2) mc2[i] = (uint16_t)((mc2[i] + 1u) & 0x3FFu);
                if (cnt2) cnt2[i]++;
            }
            if (recs && nr < cap) { recs[nr].stream = 1; recs[nr].pos = (uint32_t)(m - 1); recs[nr].state = (uint32_t)i; }
            nr++;
        }
        if (!released) {                                 /* t = 12: TB:37-45 */
            *t.p_reset = 0;
            *t.p_size = size_range;
            if (mc1) memset(mc1, 0, sizeof(uint16_t) * size_range);
            if (mc2) memset(mc2, 0, sizeof(uint16_t) * size_range);
            released = 1;
        }
        if (*t.p_reset == 0 && m == M) break;            /* TB:71 (200000 there) */
    }
    vs_delete(t.m);
    if (n_recs) *n_recs = nr;
    if (cycles_out) *cycles_out = cycles;
    return 0;
}

/* ---- many (lo, hi) pairs on host threads: the CPU arm of bench.py --------------------------------------- */
typedef struct {
    const uint32_t *E; size_t n_entries; uint32_t size; const uint8_t *data;
    uint64_t p0, p1, stride, M;
    uint64_t *counts; uint64_t cycles, symbols; int rc;
} rjob;

static void *rworker(void *arg) {
    rjob *j = (rjob *)arg;
    j->counts = (uint64_t *)calloc(j->size, sizeof(uint64_t));
    uint64_t *c2 = (uint64_t *)calloc(j->size, sizeof(uint64_t));
    j->cycles = 0; j->symbols = 0; j->rc = 0;
    for (uint64_t p = j->p0; p < j->p1; p++) {
        uint64_t cyc = 0;
        const int rc = ref_tb_run(j->E, j->n_entries, j->size, j->data + (2 * p) * j->stride, j->data + (2 * p + 1) * j->stride,
                                  j->M, 0, NULL, NULL, j->counts, c2, NULL, 0, NULL, &cyc, NULL, 0);
        if (rc) { j->rc = rc; break; }
        j->cycles += cyc;
        j->symbols += 2 * (j->M - 1);
    }
    for (uint32_t q = 0; q < j->size; q++) j->counts[q] += c2[q];
    free(c2);
    return NULL;
}

int ref_tb_run_many(const uint32_t *E, size_t n_entries, uint32_t size, const uint8_t *data, uint64_t n_pairs,
                    uint64_t stride, uint64_t M, int n_threads, uint64_t *counts, uint64_t *total_cycles,
                    uint64_t *total_symbols) {
    if (n_threads < 1) n_threads = 1;
    if ((uint64_t)n_threads > n_pairs && n_pairs > 0) n_threads = (int)n_pairs;
    rjob *jobs = (rjob *)calloc((size_t)n_threads, sizeof(rjob));
    pthread_t *th = (pthread_t *)calloc((size_t)n_threads, sizeof(pthread_t));
    for (int k = 0; k < n_threads; k++) {
        jobs[k].E = E; jobs[k].n_entries = n_entries; jobs[k].size = size; jobs[k].data = data;
        jobs[k].stride = stride; jobs[k].M = M;
        jobs[k].p0 = n_pairs * (uint64_t)k / (uint64_t)n_threads;
        jobs[k].p1 = n_pairs * (uint64_t)(k + 1) / (uint64_t)n_threads;
        pthread_create(&th[k], NULL, rworker, &jobs[k]);
    }
    uint64_t cyc = 0, sym = 0;
    int rc = 0;
    for (int k = 0; k < n_threads; k++) {
        pthread_join(th[k], NULL);
        if (jobs[k].rc) rc = jobs[k].rc;
        if (counts) for (uint32_t q = 0; q < size; q++) counts[q] += jobs[k].counts[q];
        cyc += jobs[k].cycles; sym += jobs[k].symbols;
        free(jobs[k].counts);
    }
    free(jobs); free(th);
    if (total_cycles) *total_cycles = cyc;
    if (total_symbols) *total_symbols = sym;
    return rc;
}

const char *ref_tb_module(void) { return vs_module_name(); }
