/*
 * oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the reference design (linfenghuaster/Regex-FPGA) used as the
 * parity checker for the CUDA path.  Nothing under oracle/ is linked into, called by,
 * or shipped with the product library (regex_fpga_b200/lib/librfb200.so); only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
 *
 * PARITY PINNED BY THE REFERENCE'S OWN SOURCE (round 2).  The reference ships no golden vectors, no expected log
 * and no self-checking testbench, and no HDL simulator exists in this image -- but its
 * hot path is one Verilog module in a small language subset, so it is executed anyway: oracle/vsim/v2c.py
 * translates the UNMODIFIED Design/FPGA.v (read where it lies under /root/reference, never copied) to C with
 * the language's simulation semantics, oracle/vsim/tb_driver.c clocks it exactly as
 * Simulation/testbench_BLK_Mem.sv does (plus a behavioural 1-cycle ROM for the absent design_1_wrapper), and
 * `make -C oracle _ref` builds oracle/_ref/libref.so from it.  tests/test_ref_vsim.py then requires
 *   (1) oracle A (cycle-level restatement, oracle_a.c) == the executed reference EDGE BY EDGE -- ports i,
 *       input_char_flag, accepting_match_flag(_2), FSM state, rd_address -- on 2000-entry prefixes of both shipped
 *       trace pairs (19 667 105 and 5 952 000 clock edges), and in counters, pulses and cycle totals on the full
 *       testbench runs (2 188 184 738 / 617 518 104 cycles) and on random NFAs with every row shape;
 *   (2) oracle B (functional set semantics, oracle_b.c) == the executed reference in counters and pulses;
 *   (3) both x-fills of the translator's 2-state model agree on everything the testbench observes;
 *   (4) the committed outputs of the executed reference (tests/golden/ref_vsim.json, made by
 *       tests/golden/make_ref_golden.py) == what both oracles compute, which is also what SURVEY.md
 *       Appendix C lists.  The closed-form cycle model must equal oracle A's cycle count as before.
 * What remains an assumption is stated in tb_driver.c: the block-memory IP the reference does not ship is a
 * synchronous ROM with one cycle of read latency (the only latency under which FPGA.v's own timing works).
 *
 * All citations are file:line relative to /root/reference.
 */
#ifndef RFB_ORACLE_H
#define RFB_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint32_t stream, pos, state; } orc_rec;

/* Block_Mem COE decoder (format: SURVEY Appendix A.1; slot order Design/FPGA.v:881-884).
 * Returns 0 and a malloc'd array of 32-bit entries E[4*line+slot]; caller frees with orc_free. */
int orc_coe_parse(const char *path, uint32_t **entries, size_t *n_entries);
/* Size auto-detect (the COE stores no size; Design/FPGA.v:26 takes it as a port). -1 if ambiguous. */
int64_t orc_detect_size(const uint32_t *E, size_t n_entries);
/* Simulation .mem decoder: $readmemh text, one byte per token (testbench_BLK_Mem.sv:34-35). */
int orc_mem_parse(const char *path, uint8_t **bytes, size_t *n);
void orc_free(void *p);

/* Oracle B: functional restatement.  One stream, start set {0} (Design/FPGA.v:146-147),
 * n_steps symbol steps; for k in [0,n_steps): every zero-out-degree state in S_k is a match
 * at pos k (Design/FPGA.v:210-226), S_{k+1} = successors of S_k on data[k] (FPGA.v:264-268).
 * counts[size] (u64, may be NULL) are incremented; records (may be NULL) receive up to cap
 * entries in (pos,state) ascending order, tagged with stream_id; *n_recs gets the total number
 * of matches (may exceed cap).  sum_active/max_active (may be NULL) report |S_k| statistics. */
int orc_b_scan(const uint32_t *E, size_t n_entries, uint32_t size, const uint8_t *data,
               uint64_t n_steps, uint32_t stream_id, uint64_t *counts, orc_rec *recs,
               uint64_t cap, uint64_t *n_recs, uint64_t *sum_active, uint32_t *max_active);

/* Multi-stream convenience: n_streams streams at data + s*stride, n_steps each, n_threads
 * host threads (contiguous stream ranges per thread).  Records come out in canonical
 * (stream,pos,state) order. */
int orc_b_scan_many(const uint32_t *E, size_t n_entries, uint32_t size, const uint8_t *data,
                    uint64_t n_streams, uint64_t stride, uint64_t n_steps, int n_threads,
                    uint64_t *counts, orc_rec *recs, uint64_t cap, uint64_t *n_recs,
                    uint64_t *sum_active);

/* Oracle A: cycle-level restatement of CSR_traversal (Design/FPGA.v:115-900) + a 1-cycle
 * synchronous ROM for design_1_wrapper (Design/top.v:10-13) + the testbench feeder/counters/
 * termination (testbench_BLK_Mem.sv:26-86) for an M-entry trace pair (the TB hard-codes
 * M = 200000 at :71).  lo feeds input_char / match_count, hi feeds input_char_2 /
 * match_count_2 (TB:56-57,61-69).
 *   mc1/mc2[size]   : 10-bit wrapping counters exactly as TB:21-22 (may be NULL)
 *   cnt1/cnt2[size] : the same counters without the wrap (may be NULL)
 *   recs            : match pulses as (stream 0=lo/1=hi, pos=step index, state=i), emission order
 *   cycles          : the TB's "Total no. cycles" (TB:52,84)
 * addr_bits: width of rd_address (16 in Design/FPGA.v:33; wider = documented deviation for
 * NFAs that exceed the FPGA's own address space).  fast_idle != 0 fast-forwards runs of the
 * one-cycle idle branch (FPGA.v:744-752) arithmetically; cycle count and results are identical. */
int orc_a_run(const uint32_t *E, size_t n_entries, uint32_t size, const uint8_t *lo,
              const uint8_t *hi, uint64_t M, int addr_bits, int fast_idle, uint16_t *mc1,
              uint16_t *mc2, uint64_t *cnt1, uint64_t *cnt2, orc_rec *recs, uint64_t cap,
              uint64_t *n_recs, uint64_t *cycles);

/* Oracle A with a per-edge trace of the module's ports (see oracle_a.c), without fast_idle. */
int orc_a_trace(const uint32_t *E, size_t n_entries, uint32_t size, const uint8_t *lo, const uint8_t *hi,
                uint64_t M, int addr_bits, uint64_t *trace, uint64_t trace_cap, uint64_t *cycles);

/* Closed-form cycle model, computed from oracle-B sets of both streams. */
int orc_cycle_model(const uint32_t *E, size_t n_entries, uint32_t size, const uint8_t *lo,
                    const uint8_t *hi, uint64_t M, uint64_t *cycles);

/* CPU baseline driver: runs `n_pairs` (lo,hi) stream pairs of M entries each through oracle A on
 * n_threads host threads; lo_i = data + (2i)*stride, hi_i = data + (2i+1)*stride.  Returns total
 * simulated cycles and symbols processed (2*(M-1) per pair). */
int orc_a_run_many(const uint32_t *E, size_t n_entries, uint32_t size, const uint8_t *data,
                   uint64_t n_pairs, uint64_t stride, uint64_t M, int n_threads, int fast_idle,
                   uint64_t *counts, uint64_t *total_cycles, uint64_t *total_symbols);
#ifdef __cplusplus
}
#endif
#endif
