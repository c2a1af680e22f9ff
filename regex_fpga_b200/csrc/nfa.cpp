// nfa.cpp -- validation of a CSR BRAM image (Design/FPGA.v:773,782-795,881-898).
#include "host.h"
#include "../../include/regex_fpga_b200.h"

namespace rfb {

int nfa_from_entries(const uint32_t *E, size_t n, int64_t n_states, Nfa &out, std::string &err) {
    if (!E || n < 2) { err = "empty image"; return RFB_E_INVALID; }
    if (n_states < 0) {
        n_states = detect_size(E, n);
        if (n_states < 0) { err = "cannot auto-detect the state count from the image; pass n_states"; return RFB_E_NFA; }
    }
    if (n_states == 0 || n_states > 0xFFFFFF) { err = "n_states out of range (1..2^24-1; targets are 24-bit, Design/FPGA.v:895)"; return RFB_E_NFA; }
    const uint64_t size = (uint64_t)n_states;
    if (size + 1 > n) { err = "image shorter than row_ptr"; return RFB_E_NFA; }
    if (E[0] != 0) { err = "row_ptr[0] != 0"; return RFB_E_NFA; }
    for (uint64_t s = 0; s < size; s++)
        if (E[s + 1] < E[s]) { err = "row_ptr not non-decreasing at state " + std::to_string(s); return RFB_E_NFA; }
    const uint64_t nnz = E[size];
    if (size + 1 + nnz > n) { err = "image shorter than row_ptr[size] transitions"; return RFB_E_NFA; }
    const uint32_t *tr = E + size + 1;
    for (uint64_t j = 0; j < nnz; j++)
        if ((tr[j] & 0xFFFFFFu) >= size) { err = "transition " + std::to_string(j) + " targets a state >= size"; return RFB_E_NFA; }
    out.n_states = (uint32_t)size;
    out.nnz = (uint32_t)nnz;
    out.entries.assign(E, E + n);
    out.n_accepting = 0;
    for (uint64_t s = 0; s < size; s++) out.n_accepting += (E[s + 1] == E[s]);
    return RFB_OK;
}

}  // namespace rfb
