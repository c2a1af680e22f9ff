// host.h -- host-side model of the reference's data: file formats, CSR NFA, execution image.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

namespace rfb {

// ---- formats.cpp ------------------------------------------------------------------------------
// Xilinx .coe BRAM image <-> 32-bit entries (Block_Mem/CSR_BlockMem*.coe; entry 4*line+slot,
// slot 0 = rd_bus[127:96], Design/FPGA.v:881-884).  Throw-free: return 0 or a negative rfb_status
// and fill `err`.
int coe_parse_file(const std::string &path, std::vector<uint32_t> &entries, std::string &err);
int coe_write_file(const std::string &path, const uint32_t *entries, size_t n, int style,
                   std::string &err);
// $readmemh byte traces (Simulation/input_trace_*.mem, testbench_BLK_Mem.sv:34-35).
int mem_parse_file(const std::string &path, std::vector<uint8_t> &bytes, std::string &err);
int mem_write_file(const std::string &path, const uint8_t *bytes, size_t n, std::string &err);
// The image does not store `size` (Design/FPGA.v:26 takes it as a port); -1 if not unique.
int64_t detect_size(const uint32_t *entries, size_t n);

// ---- nfa.cpp ----------------------------------------------------------------------------------
struct Nfa {
    uint32_t n_states = 0;
    uint32_t nnz = 0;
    uint32_t n_accepting = 0;
    std::vector<uint32_t> entries;  // the BRAM image as loaded (row_ptr | transitions | padding)
    const uint32_t *row_ptr() const { return entries.data(); }
    const uint32_t *trans() const { return entries.data() + n_states + 1; }  // Design/FPGA.v:773,793
    uint32_t degree(uint32_t s) const { return row_ptr()[s + 1] - row_ptr()[s]; }
};
// Validates and adopts an image.  n_states < 0: auto-detect.
int nfa_from_entries(const uint32_t *entries, size_t n, int64_t n_states, Nfa &out, std::string &err);

// ---- parts.cpp --------------------------------------------------------------------------------
// An NFA whose tables do not fit one SM's shared memory is cut along its connected components (the start state 0,
// which nothing targets, is shared): the NFA step is a union over active states and no transition crosses
// components, so scanning a stream against every part and merging the reports is exact.  BASELINE config 5
// (7 x snort_16 behind one start state, 66 592 states) becomes 7 parts of 9 514 states.
// groups[g] = sorted original ids (without state 0) of part g; empty result = cannot be split.
void nfa_components(const Nfa &nfa, uint32_t max_states_per_part, std::vector<std::vector<uint32_t>> &groups);
// sub-NFA of state 0 (row restricted to the part) + `states`; to_orig[i] = original id of sub state i
int nfa_extract(const Nfa &nfa, const std::vector<uint32_t> &states, Nfa &sub, std::vector<uint32_t> &to_orig, std::string &err);

// ---- ecsr.cpp ---------------------------------------------------------------------------------
// Edge-grouped CSR for the general kernel: a state's transitions grouped by target, one 64-bit record per
// (symbol-set -> target) edge instead of one entry per (symbol, target) pair.  A state that self-loops on all
// 256 symbols is 1 record instead of 256 CSR entries (the FPGA streams all 256 through its 4-lane compare,
// Design/FPGA.v:227-407).  record = a[7:0] | b[15:8] | is_class[16] | set_id[31:17] | target[55:32]
//   pair  : edge taken iff c == a or c == b;   class : iff bit c of memb[set_id] (256-bit bitmap) is set
struct Ecsr {
    std::vector<uint32_t> eptr;    // [n_states + 1]; eptr[s] == eptr[s+1]  <=>  accepting (zero out-degree)
    std::vector<uint64_t> erec;    // [n_edges]
    std::vector<uint32_t> memb;    // [n_sets * 8]
    uint32_t n_sets = 0;
};
int ecsr_build(const Nfa &nfa, Ecsr &out, std::string &err);   // includes an exhaustive check against the CSR

// ---- image.cpp --------------------------------------------------------------------------------
// Execution image: the CSR re-indexed at load time for the lane kernel (one thread per stream).
// See DESIGN.md "Execution image" for the layout; image_successors() is its executable definition
// and image_verify() proves it equivalent to the CSR for every (state, symbol).
struct ImageOptions {
    int sticky_words = 0;       // 0 = auto (1 or 2)
    int sticky_min_self = 16;   // a state is mask-resident if it self-loops on >= this many symbols
    int bucket_bits = -1;       // -1 = auto; buckets per branching state = 1 << bucket_bits
    int accel = 1;              // build the start DFA for the always-active sticky state
    uint32_t dfa_max_states = 32766;   // <= 32766 (15-bit ids)
    int dfa_absorb = 12;        // up to this many sticky states that the start DFA enters are tracked BY the DFA (as ordinary
                                // members with a self loop) instead of by the mask, while the DFA stays complete: see image_build
    std::vector<uint32_t> not_sticky;   // internal (image_build): states kept out of the sticky mask
    bool verify = true;         // internal: trial builds skip the equivalence proof, the final build runs it
    uint32_t fixed_hash_mul = 0, fixed_hash_shift = 0;   // internal: reuse a bucket hash instead of searching (mul != 0)
    bool probe_dfa_only = false;   // internal: stop after the start DFA (its size decides whether a sticky state may move into it)
    uint32_t max_bytes = 200 * 1024;
};

struct ImageHeader {  // mirrored on the device (passed by value to the kernels)
    uint32_t n_slots;       // entries in tab
    uint32_t gbase;         // first slot of the hashed rows (branching states, then sticky rows)
    uint32_t nsb;           // sticky bits = 64 * sticky_words; ids < nsb are mask-resident
    uint32_t sticky_words;
    uint32_t bucket_bits;
    uint32_t hash_mul;      // h(c) = ((c * hash_mul) >> hash_shift) & 0xFF; bucket = h(c) & row mask
    uint32_t hash_shift;
    uint32_t start_id;      // internal id of state 0
    uint32_t n_sets;
    uint32_t acc_base;      // accepting states own the contiguous id range [acc_base, acc_base + n_acc)
    uint32_t n_acc;
    uint32_t srow_base;     // first slot of the sticky rows (sized per state, see off_sdesc)
    // byte offsets of the sections inside the blob (all 16-byte aligned)
    uint32_t off_tab, off_mask, off_memb, off_sdesc;   // sdesc[b] = row base | (row mask << 16) of sticky bit b
    uint32_t off_cmap;      // cmap[c] = start-DFA symbol class | h(c) << 16
    uint32_t off_look;      // look[c'][sticky_words] (64-bit words): sticky states whose firing can matter when the NEXT symbol is c
    // start DFA of the always-active sticky state (bit 0), tables in global memory; accel == 0: absent
    uint32_t accel, dfa_ncls, dfa_states;
    uint32_t blob_bytes;
};

struct Image {
    bool ok = false;
    std::string why_not;               // reason when !ok
    ImageHeader h{};
    std::vector<uint8_t> blob;         // tab | mask | memb, staged verbatim into shared memory
    std::vector<uint32_t> orig_of_id;  // internal id -> original state id (0xFFFFFFFF: not a state)
    std::vector<uint32_t> id_of_orig;  // original state id -> internal id
    uint32_t n_sticky = 0;
    uint32_t n_absorbed = 0;                            // sticky states tracked by the start DFA instead of the mask
    uint32_t n_sticky_dropped = 0;                      // self-looping states that did not fit the mask (run as ordinary states)
    // start DFA (see image.cpp): the successors of the always-active state A (bit 0) that are neither sticky nor
    // accepting are never materialised; a per-stream DFA state d stands for the set of them that is active.
    uint32_t accel_state = 0xFFFFFFFFu;   // original id of A
    struct Dfa {
        uint32_t ncls = 1, n = 1;          // symbol classes, DFA states (0 = A not active yet, 1 = A alone)
        std::vector<uint16_t> dt;          // [n * ncls]: next state | 0x8000 if the transition has an insertion list
        std::vector<uint32_t> dta;         // [n * ncls]: index of that list in act
        std::vector<uint16_t> act;         // insertion lists: internal id | 0x8000 if another entry follows
        std::vector<uint32_t> mem_ptr;     // [n + 1]: members of each DFA state ...
        std::vector<uint16_t> mem_ids;     // ... as internal ids (never sticky, never accepting)
        uint32_t n_frontier = 0;           // states whose rows fall back to a shorter history
    } dfa;
    std::vector<std::pair<int, uint32_t>> dfa_sticky_targets;   // sticky states the DFA enters: (-symbols they fire on, original id)
};

// tab entry encoding
constexpr uint32_t TAB_MORE = 0x80000000u;
inline uint32_t tab_pack(uint32_t a, uint32_t b, uint32_t tgt, bool more) {
    return (a & 0xFF) | ((b & 0xFF) << 8) | ((tgt & 0x7FFF) << 16) | (more ? TAB_MORE : 0u);
}
// a > b encodes a special: code = (0xFF - a) * 256 + b
constexpr uint32_t CODE_INDIRECT = 2;
inline uint32_t tab_special(uint32_t code, uint32_t tgt, bool more) {
    return tab_pack(0xFF - (code >> 8), code & 0xFF, tgt, more);
}

int image_build(const Nfa &nfa, const ImageOptions &opt, Image &img, std::string &err);
// Successors of original state s on symbol c as computed THROUGH the image (sorted, original ids);
// *accepting reports whether the image treats s as accepting.
void image_successors(const Image &img, uint32_t s, uint32_t c, std::vector<uint32_t> &out,
                      bool *accepting);
// Exhaustive equivalence check image == CSR.  Returns 0 or RFB_E_INTERNAL with a message.
int image_verify(const Nfa &nfa, const Image &img, std::string &err);

// ---- imagefile.cpp ----------------------------------------------------------------------------
// The scan plan of an NFA: the NFA as loaded and its parts (one, or several when the whole NFA's tables do not fit
// one SM), each with its execution image.  plan_write / plan_read are the on-disk form (layout in imagefile.cpp);
// plan_read trusts nothing: checksum, bounds, image_validate_structure() and image_verify() for every part.
struct PlanPart {
    Nfa sub;                           // the part as an NFA of its own (state 0 + the part's states)
    Image img;
    std::vector<uint32_t> to_orig;     // sub state id -> reference state id (empty: identity)
};
struct Plan {
    Nfa host;
    std::vector<PlanPart> parts;       // >= 1
};
int plan_build(const uint32_t *entries, size_t n_entries, int64_t n_states, const ImageOptions &opt, bool allow_split,
               Plan &plan, std::string &err);
int plan_write(const Plan &plan, const std::string &path, std::string &err);
int plan_read(const std::string &path, Plan &plan, std::string &err);
// every index an interpreter of the tables can follow stays inside them (run before image_verify on untrusted input)
int image_validate_structure(const Image &img, uint32_t n_states, std::string &err);

}  // namespace rfb
