// device.h -- structures shared between the C ABI (api.cu) and the kernels (scan_*.cu).
#pragma once
#include "../../include/regex_fpga_b200.h"
#include "host.h"
#include <cuda_runtime.h>

namespace rfb {

// Zeroed before every scan; read back (32 bytes) after it.
struct ScanGlobals {
    unsigned long long n_matches;   // every accept pulse, recorded or not
    unsigned long long n_symbols;   // symbol steps executed (ragged batches only; else host-computed)
    unsigned int next_stream;       // lane kernel: dynamic stream fetch
    unsigned int n_rescan;          // streams queued for the warp kernel
    unsigned int next_item;         // warp kernel: dynamic work fetch
    unsigned int n_rescan_total;    // handed-over streams (accumulates when the three above are reset between launches)
    unsigned int chunks_ready;      // rfb_scan: input chunks whose H2D copy has completed (written by the copy stream)
    unsigned int pad[3];
    // parts of a cut NFA scanned in one launch (scan_lane_multi_kernel): per-part stream fetch, hand-over count, hand-over fetch
    unsigned int part_next[16];
    unsigned int part_rescan[16];
    unsigned int part_item[16];
};
constexpr uint32_t MAX_MULTI_PARTS = 16;

struct BatchDev {
    const uint8_t *data;
    unsigned long long n_streams;
    unsigned long long stride;
    const unsigned long long *offsets;  // nullable
    const unsigned int *steps;          // nullable
    unsigned int n_steps;
    unsigned int stream_id_base;
    // rfb_scan only: streams [c*chunk_streams, (c+1)*chunk_streams) become readable when *ready > c (0: no gating)
    unsigned int chunk_streams;
    const unsigned int *ready;
    // resumable scans: per-stream active sets in / out (original state ids), stride 1 + state_cap words
    unsigned int count_symbols;  // ragged batches: add the stream lengths to n_symbols (first part of a multi-part NFA only)
    unsigned int pos_base;
    unsigned int state_cap;
    const unsigned int *state_in;
    unsigned int *state_out;
    unsigned int state_append;   // parts after the first add their members to the set the earlier parts wrote
};

struct OutDev {
    unsigned long long *counts;  // nullable
    rfb_match *records;          // nullable
    unsigned long long capacity;
    ScanGlobals *g;
    uint2 *rescan;               // (stream, first pos to report) pairs, capacity n_streams (x parts in a multi-part launch)
    const unsigned int *q_rescan_n;  // general kernel: hand-over queue to drain instead of g->n_rescan / g->next_item (NULL: those)
    unsigned int *q_next_item;
};

struct NfaDev {
    uint32_t n_states;
    uint32_t n_ref_states;       // states of the NFA as loaded (ids in state_in / state_out rows, indices of id_of_orig / sub_of_ref)
    const uint32_t *row_ptr;     // [n_states + 1]
    const uint32_t *trans;       // [nnz]  {symbol[31:24], target[23:0]}  Design/FPGA.v:888-898
    // edge-grouped CSR (general kernel)
    const uint32_t *eptr;        // [n_states + 1]
    const unsigned long long *erec;   // [n_edges]
    const uint32_t *emembs;      // [n_sets * 8]
    const uint32_t *state_map;   // general kernel: state id of this (sub-)NFA -> reference state id; NULL = identity
    const uint32_t *sub_of_ref;  // reference state id -> state id of this sub-NFA (0xFFFFFFFF: other part); NULL = identity
    // execution image
    const uint8_t *blob;         // ImageHeader::blob_bytes bytes, 16-byte aligned
    const uint32_t *orig_of_id;  // [n_slots]
    const uint32_t *id_of_orig;  // [n_states]
    // start DFA (image.cpp): tables in global memory
    const uint16_t *dfa_dt;      // [dfa_states * dfa_ncls]: next state | 0x8000 if the transition has an insertion list
    const uint32_t *dfa_dta;     // [dfa_states * dfa_ncls]: index of that list in dfa_act
    const uint16_t *dfa_act;     // insertion lists: internal id | 0x8000 if another entry follows
    const uint32_t *dfa_mem_ptr; // [dfa_states + 1]: never-materialised members of each DFA state ...
    const uint16_t *dfa_mem_ids; // ... as internal ids
    // A part of a cut NFA whose share of state 0's row is empty would see state 0 as a zero-out-degree (accepting) state
    // although the full NFA's state 0 has edges (into other parts); a genuinely accepting state 0 is reported by the first
    // part only.  no_report_lane / no_report_sub: the id (internal / sub-NFA) whose pulses this part must not report.
    uint32_t no_report_lane, no_report_sub;
    uint32_t hot_rows, hot_bytes; // set per launch: rows of dfa_dt staged into shared memory / bytes copied for them (16-byte multiple)
    ImageHeader h;
};

// lane kernel geometry
constexpr int LANE_THREADS = 1024;   // one CTA per SM
// warp kernel geometry
constexpr int WARP_THREADS = 256;
constexpr int WARP_LCAP = 512;       // sparse list entries per warp before it switches to bitmap scans

size_t lane_smem_bytes(const ImageHeader &h);
uint32_t lane_hot_rows(const ImageHeader &h, uint32_t *copy_bytes);
int lane_ring_cap(const ImageHeader &h);
size_t warp_smem_bytes(uint32_t n_states, int warps_per_cta);

// Enqueue the lane kernel (one thread per stream, tables in shared memory).
cudaError_t launch_scan_lane(const NfaDev &nfa, const BatchDev &batch, const OutDev &out, int n_sms,
                             cudaStream_t stream);
cudaError_t launch_scan_lane_multi(const NfaDev *d_parts, uint32_t n_parts, uint32_t sticky_words, int cap, size_t smem,
                                   const BatchDev &batch, const OutDev &out, int n_sms, cudaStream_t stream);
// Enqueue the general kernel (one warp per stream, CSR in global memory / L2).
//   from_rescan = false: all streams of the batch;  true: the (stream, first_pos) pairs queued by the
//   lane kernel -- the count is read on the device, so no host round trip is needed in between.
cudaError_t launch_scan_warp(const NfaDev &nfa, const BatchDev &batch, const OutDev &out, bool from_rescan,
                             int n_sms, cudaStream_t stream);
cudaError_t launch_tb_cycles(const NfaDev &nfa, const uint32_t *cost, const uint8_t *lo, const uint8_t *hi, uint32_t n_steps,
                             unsigned long long *total, cudaStream_t stream);
// canonical (stream, pos, state) order of records[0..n) on the device (sort.cu)
cudaError_t launch_sort_records(rfb_match *records, rfb_match *tmp, unsigned long long n, uint32_t key_bytes, uint32_t *hist,
                                cudaStream_t stream);
size_t sort_hist_words(unsigned long long n);
cudaError_t configure_kernels();
constexpr size_t MAX_DYN_SMEM = 227 * 1024;   // per-CTA opt-in limit on sm_100

}  // namespace rfb
