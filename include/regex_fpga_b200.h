/*
 * regex_fpga_b200.h -- C ABI of the B200-native CSR-NFA scan engine.
 *
 * Drop-in boundary for the one hot path of linfenghuaster/Regex-FPGA: what Design/top.v +
 * Design/FPGA.v compute when driven by Simulation/testbench_BLK_Mem.sv.  The reference has no
 * software API; its boundary is the port list of `top` (Design/top.v:1-3) plus two file formats.
 * Each entry point below names the reference interface it replaces (file:line relative to the
 * reference repository).
 *
 * Conventions: every function returns 0 (RFB_OK) or a negative rfb_status; no exceptions cross the
 * boundary; handles are opaque; the library owns all device memory it allocates; a context is bound
 * to ONE GPU and may be used by one host thread at a time (one process per GPU -- multi-GPU jobs
 * shard streams across processes).  There is NO CPU fallback: if no CUDA device
 * is usable rfb_ctx_create fails with RFB_E_NODEVICE.
 */
#ifndef REGEX_FPGA_B200_H
#define REGEX_FPGA_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 3: + rfb_nfa_describe, execution-image files, rfb_scan_submit / rfb_scan_wait (structs unchanged since 2)
 * 4: + rfb_nfa_calibrate / rfb_nfa_calibration, rfb_group_* (structs unchanged) */
#define RFB_ABI_VERSION 4

typedef enum rfb_status {
    RFB_OK = 0,
    RFB_E_INVALID = -1,     /* bad argument */
    RFB_E_IO = -2,          /* file could not be read / written */
    RFB_E_FORMAT = -3,      /* malformed .coe / .mem text */
    RFB_E_NFA = -4,         /* CSR image fails validation (row_ptr order, target range, size) */
    RFB_E_CUDA = -5,        /* a CUDA call failed; see rfb_last_error */
    RFB_E_NODEVICE = -6,    /* no usable CUDA device (the library has no CPU path) */
    RFB_E_NOMEM = -7,
    RFB_E_UNSUPPORTED = -8,
    RFB_E_INTERNAL = -9     /* execution image failed self-verification against the CSR */
} rfb_status;

typedef struct rfb_ctx rfb_ctx;   /* one GPU, its streams and scratch memory */
typedef struct rfb_nfa rfb_nfa;   /* one CSR NFA resident on that GPU */

/* One accepting_match_flag / accepting_match_flag_2 pulse (Design/FPGA.v:215-223) as sampled by the
 * testbench (testbench_BLK_Mem.sv:61-69): `state` is the value of port i[19:0] during the pulse,
 * `pos` is the index k of the symbol step in which the accepting state was found in the current
 * set S_k (one step after the symbol that completed the match), `stream` is the stream index. */
typedef struct rfb_match {
    uint32_t stream;
    uint32_t pos;
    uint32_t state;
} rfb_match;

typedef struct rfb_nfa_info {
    uint32_t n_states;        /* `size` port, Design/FPGA.v:26 */
    uint32_t n_transitions;   /* row_ptr[size] */
    uint32_t n_accepting;     /* zero-out-degree states, Design/FPGA.v:210-213 */
    uint32_t n_entries;       /* 32-bit entries of the BRAM image incl. padding */
    /* execution image (load-time re-indexing of the CSR for the lane kernel) */
    uint32_t image_ok;        /* 1: lane kernel usable; 0: only the general warp kernel */
    uint32_t image_bytes;     /* shared-memory bytes of the staged tables */
    uint32_t n_sticky;        /* self-looping states held in the per-stream bit mask */
    uint32_t sticky_words;    /* 64-bit words of that mask */
    uint32_t n_slots;         /* entries of the edge table */
    uint32_t n_class_sets;    /* distinct symbol classes with more than two members */
    uint32_t bucket_bits;     /* log2 of buckets per branching state */
    uint32_t n_parts;         /* > 1: the NFA is scanned as several independent groups of connected components,
                                 each with tables that fit one SM (image_* fields describe the first) */
} rfb_nfa_info;

/* A batch of independent byte streams.  Stream s occupies bytes
 *   data + (offsets ? offsets[s] : s * stride) ... + steps(s),   steps(s) = steps ? steps[s] : n_steps.
 * All pointers are HOST pointers for rfb_scan and DEVICE pointers for rfb_scan_device.  The device
 * buffer must be readable up to the next multiple of 16 bytes past the last stream byte. */
typedef struct rfb_batch {
    const uint8_t *data;
    uint64_t data_bytes;       /* size of the data buffer (bounds checks / staging) */
    uint64_t n_streams;
    uint64_t stride;           /* used when offsets == NULL */
    const uint64_t *offsets;   /* optional, n_streams entries */
    uint32_t n_steps;          /* symbol steps per stream when steps == NULL */
    const uint32_t *steps;     /* optional per-stream step counts (ragged batches) */
    uint32_t stream_id_base;   /* added to the stream field of every record (sharded jobs) */
    uint32_t pos_base;         /* added to the pos field of every record (resumed streams) */
    /* Resumable scans -- the software form of the input_char_flag handshake (Design/FPGA.v:125,740,763): the
     * design is a streaming device whose only carry-over state is the active set.  Per stream, state_stride =
     * 1 + state_cap 32-bit words: word 0 = number of active states, words 1.. = their state ids (any order).
     * state_in  == NULL: every stream starts from the reset state {0} (Design/FPGA.v:146-147).
     * state_out == NULL: the final set S_{n_steps} is discarded, as the testbench does (TB:71-86).
     * A final set with more than state_cap members is reported as count 0xFFFFFFFF (RFB_STATE_OVERFLOW). */
    const uint32_t *state_in;
    uint32_t *state_out;
    uint32_t state_cap;
    uint32_t reserved;
} rfb_batch;
#define RFB_STATE_OVERFLOW 0xFFFFFFFFu

/* flags for rfb_scan / rfb_scan_device */
#define RFB_SCAN_DEFAULT      0u
#define RFB_SCAN_SORT_RECORDS 1u   /* records in canonical (stream,pos,state) ascending order (sorted on the GPU;
                                      rfb_scan_device: not together with RFB_SCAN_ASYNC) */
#define RFB_SCAN_FORCE_WARP   2u   /* use only the general warp-per-stream kernel */
#define RFB_SCAN_NO_COUNTS    4u   /* skip per-state counters */
#define RFB_SCAN_ASYNC        8u   /* rfb_scan_device: enqueue only; call rfb_scan_collect later */
#define RFB_SCAN_ACCUMULATE  16u   /* rfb_scan_device: do not zero counts/records before the scan */

typedef struct rfb_result {
    /* caller-provided buffers (host memory for rfb_scan, device memory for rfb_scan_device) */
    uint64_t *counts;          /* [n_states] matches per state id, or NULL */
    rfb_match *records;        /* [record_capacity] or NULL */
    uint64_t record_capacity;
    /* outputs */
    uint64_t n_matches;        /* all pulses, whether or not a record slot was available */
    uint64_t n_records;        /* min(n_matches, record_capacity) */
    uint64_t n_dropped;        /* n_matches - n_records; overflow is counted, never undefined */
    uint64_t n_symbols;        /* symbol steps executed */
    uint64_t n_rescanned;      /* streams handed from the lane kernel to the warp kernel */
    float gpu_ms;              /* device time of the scan kernels (CUDA events) */
    uint32_t n_launches;       /* kernels launched by this call */
} rfb_result;

/* ---- context ------------------------------------------------------------------------------- */
/* Replaces: powering up the board.  device_id is a CUDA ordinal. */
int rfb_ctx_create(int device_id, rfb_ctx **out);
void rfb_ctx_destroy(rfb_ctx *ctx);
/* Message of the most recent failure on this thread (ctx may be NULL). Never NULL. */
const char *rfb_last_error(const rfb_ctx *ctx);
int rfb_abi_version(void);

/* ---- transition memory --------------------------------------------------------------------- */
/* Replaces: BRAM initialisation of design_1_wrapper from a Xilinx .coe (Design/top.v:10-13,
 * Block_Mem/CSR_BlockMem*.coe) and the `size` port (Design/FPGA.v:26, testbench_BLK_Mem.sv:20).
 * n_states < 0 auto-detects the size from the image (the .coe does not store it). */
int rfb_nfa_load_coe(rfb_ctx *ctx, const char *path, int64_t n_states, rfb_nfa **out);
/* Same, from BRAM contents already in memory: entries[4*line+slot], slot 0 = rd_bus[127:96]
 * (Design/FPGA.v:881-884). */
int rfb_nfa_from_entries(rfb_ctx *ctx, const uint32_t *entries, size_t n_entries, int64_t n_states,
                         rfb_nfa **out);
void rfb_nfa_destroy(rfb_nfa *nfa);
int rfb_nfa_get_info(const rfb_nfa *nfa, rfb_nfa_info *info);
/* Copies the BRAM image back (n_entries from rfb_nfa_get_info). */
int rfb_nfa_get_entries(const rfb_nfa *nfa, uint32_t *entries, size_t capacity);

/* Traffic calibration.  The hot kernel keeps the most visited rows of its start-DFA table in shared memory;
 * which rows those are depends on the traffic.  The library measures it by itself on the first large
 * (>= 8192 streams) uniformly strided batch an NFA scans -- a host-side walk over 2048 sampled streams, a few
 * tens of milliseconds, once -- unless RFB_NO_CALIBRATE is set.  rfb_nfa_calibrate does the same explicitly
 * on a HOST batch of the caller's choice (e.g. before the first timed scan, or again when the traffic has
 * changed); it waits for scans in flight.  Calibration never changes a result, only speed.  No reference
 * counterpart: the FPGA reads every row from the same block RAM (Design/FPGA.v:773-795).
 * rfb_nfa_calibration returns 1 if the NFA is calibrated (0 if not) and reports the sample size in symbols
 * and the share of the sample's start-DFA lookups that land in shared-memory rows (either may be NULL). */
int rfb_nfa_calibrate(rfb_ctx *ctx, rfb_nfa *nfa, const rfb_batch *sample);
int rfb_nfa_calibration(const rfb_nfa *nfa, uint64_t *sample_symbols, double *hot_fraction);

/* ---- host-side format helpers (no GPU involved) --------------------------------------------- */
/* Replaces: $readmemh of Simulation/input_trace_{hi,lo}_*.mem (testbench_BLK_Mem.sv:34-35).
 * *bytes is malloc'd by the library; release with rfb_free. */
int rfb_trace_load_mem(const char *path, uint8_t **bytes, size_t *n);
int rfb_trace_write_mem(const char *path, const uint8_t *bytes, size_t n);
/* .coe text <-> entries.  style 0: one 128-bit word per line, no terminator (as
 * CSR_BlockMem_snort_16.coe); style 1: blank-separated on one line, ';' terminated (as
 * CSR_BlockMem.coe). */
int rfb_coe_parse(const char *path, uint32_t **entries, size_t *n_entries);
int rfb_coe_write(const char *path, const uint32_t *entries, size_t n_entries, int style);
int64_t rfb_coe_detect_size(const uint32_t *entries, size_t n_entries);
void rfb_free(void *p);
/* Host-only self-check of the load-time re-indexing: validates the image, builds the lane kernel's
 * execution tables and proves them equivalent to the CSR for every (state, symbol) pair -- the same
 * verification every rfb_nfa_load_coe / rfb_nfa_from_entries runs before uploading.  sticky_words and
 * bucket_bits <= 0 select the defaults.  Fills *info (image_ok = 0 with RFB_OK means the NFA is valid
 * but needs the general warp kernel, e.g. tables too large for shared memory). */
int rfb_image_check(const uint32_t *entries, size_t n_entries, int64_t n_states, int sticky_words,
                    int bucket_bits, rfb_nfa_info *info);
/* One line of text per part describing how the NFA will be scanned (tables, sticky states, start DFA:
 * states, symbol classes, states beyond the budget, insertion-list entries).  Writes at most cap bytes
 * including the terminator; returns RFB_OK, or RFB_E_INVALID if cap is too small. */
int rfb_nfa_describe(const rfb_nfa *nfa, char *buf, size_t cap);

/* ---- execution-image files ------------------------------------------------
 * The load-time re-indexing as an on-disk artefact: the BRAM image as loaded plus the tables the
 * lane kernel runs on (per part, for an NFA that is cut into parts), layout documented in
 * regex_fpga_b200/csrc/imagefile.cpp and DESIGN.md.  Saving skips nothing on the way back: loading
 * checks the checksum, bounds, the structure of every table, and then proves the tables equivalent
 * to the CSR exactly as a freshly built image is proven (RFB_E_FORMAT otherwise).  What loading
 * saves is the table construction (hash search, start-DFA subset construction).
 *   rfb_nfa_save_image / rfb_nfa_load_image : from / to an NFA resident on a context's GPU
 *   rfb_image_file_build / rfb_image_file_check : host-only (no GPU): build and write, read and verify */
int rfb_nfa_save_image(const rfb_nfa *nfa, const char *path);
int rfb_nfa_load_image(rfb_ctx *ctx, const char *path, rfb_nfa **out);
int rfb_image_file_build(const uint32_t *entries, size_t n_entries, int64_t n_states, const char *path);
int rfb_image_file_check(const char *path, rfb_nfa_info *info);
/* Number of symbol steps the testbench executes on an M-entry trace: M-1 (the last entry is loaded
 * but never processed and the final set is never examined; testbench_BLK_Mem.sv:71-86). */
uint32_t rfb_tb_steps(uint32_t trace_entries);

/* ---- the scan -------------------------------------------------------------------------------
 * Replaces: the clk/reset/input_char/input_char_2/input_char_flag loop of `top`
 * (Design/top.v:1-3, Design/FPGA.v:23-37) and the testbench's match counters
 * (testbench_BLK_Mem.sv:53-69).  Every stream starts from the reset state {0}
 * (Design/FPGA.v:146-147).  A testbench run is n_streams = 2 (stream 0 = lo trace = input_char /
 * match_count, stream 1 = hi trace = input_char_2 / match_count_2), n_steps = rfb_tb_steps(M). */
int rfb_scan(rfb_ctx *ctx, const rfb_nfa *nfa, const rfb_batch *batch, uint32_t flags,
             rfb_result *result);
/* Device-resident variant: batch pointers and result->counts / result->records are device memory
 * on the context's GPU; work is enqueued on `cuda_stream` (a cudaStream_t, 0 = the context's own
 * stream).  Without RFB_SCAN_ASYNC the call synchronises the stream and fills the scalar outputs. */
int rfb_scan_device(rfb_ctx *ctx, const rfb_nfa *nfa, const rfb_batch *batch, uint32_t flags,
                    void *cuda_stream, rfb_result *result);
/* Completes an RFB_SCAN_ASYNC call: synchronises and fills the scalar outputs of *result. */
int rfb_scan_collect(rfb_ctx *ctx, rfb_result *result);

/* Pipelined host path: rfb_scan with up to TWO host batches in flight, so that the next batch's H2D copy
 * overlaps the tail of the current scan, its record sort and the D2H of its results (a lone rfb_scan
 * leaves the PCIe link idle during those).  rfb_scan_submit enqueues the copy and the kernels of a
 * uniformly strided batch (no offsets / steps / state_in / state_out: RFB_E_UNSUPPORTED) and returns;
 * the batch data, result->counts and result->records must stay valid (and should be pinned) until
 * rfb_scan_wait has returned that result.  rfb_scan_wait completes the OLDEST submitted batch, fills
 * its rfb_result exactly as rfb_scan would, and stores its address in *done (nullable). */
int rfb_scan_submit(rfb_ctx *ctx, const rfb_nfa *nfa, const rfb_batch *batch, uint32_t flags,
                    rfb_result *result);
int rfb_scan_wait(rfb_ctx *ctx, rfb_result **done);

/* ---- multi-GPU groups (one process, N GPUs) -------------------------------------------------
 * Replaces: nothing in the reference (one FPGA, one clock domain) -- this is how BASELINE config 4 is
 * driven from a C host.  What makes it legal is the reference's own structure: its two streams share
 * only reads of the transition memory (Design/FPGA.v:54-57,264-268) and every stream starts from the
 * reset state {0} (FPGA.v:146-147), so streams shard freely.  A group is N contexts + N NCCL
 * communicators (ncclCommInitAll; NCCL is loaded with dlopen at the first rfb_group_create, the
 * library has no link-time dependency on it).  rfb_group_scan cuts a HOST batch into N contiguous
 * stream shards, scans them concurrently (one host thread per GPU), SUM-all-reduces the per-state
 * counts over NCCL and writes the shards' records, in shard order = canonical order when
 * RFB_SCAN_SORT_RECORDS is set, into the caller's buffer.  Flags: RFB_SCAN_SORT_RECORDS,
 * RFB_SCAN_FORCE_WARP, RFB_SCAN_NO_COUNTS.  gpu_ms is the maximum over the GPUs.
 * A multi-process job (one process per GPU, as bench.py runs) uses rfb_ctx / rfb_scan per rank with
 * stream_id_base and its own communicator instead; see INTEGRATION.md. */
typedef struct rfb_group rfb_group;
typedef struct rfb_group_nfa rfb_group_nfa;
int rfb_group_create(const int *device_ids, int n, rfb_group **out);
void rfb_group_destroy(rfb_group *group);
int rfb_group_size(const rfb_group *group);
rfb_ctx *rfb_group_ctx(rfb_group *group, int i);             /* member context (owned by the group) */
const char *rfb_group_last_error(const rfb_group *group);
int rfb_group_nfa_load_coe(rfb_group *group, const char *path, int64_t n_states, rfb_group_nfa **out);
int rfb_group_nfa_from_entries(rfb_group *group, const uint32_t *entries, size_t n_entries,
                               int64_t n_states, rfb_group_nfa **out);
void rfb_group_nfa_destroy(rfb_group_nfa *nfa);
rfb_nfa *rfb_group_nfa_member(rfb_group_nfa *nfa, int i);    /* member NFA on GPU i (owned by the handle) */
int rfb_group_scan(rfb_group *group, const rfb_group_nfa *nfa, const rfb_batch *batch, uint32_t flags,
                   rfb_result *result);

/* Informational: the testbench's "Total no. cycles" (testbench_BLK_Mem.sv:52,84) for an M-entry
 * (lo,hi) trace pair, from the closed-form cycle model of Design/FPGA.v, evaluated on
 * the GPU from the per-step active sets of both streams.  Host pointers. */
int rfb_fpga_cycles(rfb_ctx *ctx, const rfb_nfa *nfa, const uint8_t *lo, const uint8_t *hi,
                    uint32_t trace_entries, uint64_t *cycles);

#ifdef __cplusplus
}
#endif
#endif /* REGEX_FPGA_B200_H */
