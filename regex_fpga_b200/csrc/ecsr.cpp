// ecsr.cpp -- edge-grouped CSR for the general (warp-per-stream) kernel; see host.h.
#include "host.h"
#include "../../include/regex_fpga_b200.h"
#include <algorithm>
#include <array>
#include <map>

namespace rfb {

int ecsr_build(const Nfa &nfa, Ecsr &out, std::string &err) {
    const uint32_t N = nfa.n_states;
    const uint32_t *rp = nfa.row_ptr();
    const uint32_t *tr = nfa.trans();
    out = Ecsr();
    out.eptr.assign(N + 1, 0);
    std::map<std::array<uint64_t, 4>, uint32_t> set_id;
    for (uint32_t s = 0; s < N; s++) {
        std::map<uint32_t, std::array<uint64_t, 4>> by_tgt;     // target -> symbol set
        for (uint32_t j = rp[s]; j < rp[s + 1]; j++) {
            const uint32_t c = tr[j] >> 24;
            auto &w = by_tgt.emplace(tr[j] & 0xFFFFFFu, std::array<uint64_t, 4>{{0, 0, 0, 0}}).first->second;
            w[c >> 6] |= 1ull << (c & 63);
        }
        for (auto &kv : by_tgt) {
            std::vector<uint32_t> m;
            for (uint32_t c = 0; c < 256; c++) if ((kv.second[c >> 6] >> (c & 63)) & 1) m.push_back(c);
            uint64_t rec;
            if (m.size() <= 2) rec = (uint64_t)m[0] | ((uint64_t)m.back() << 8);
            else {
                auto it = set_id.find(kv.second);
                if (it == set_id.end()) it = set_id.emplace(kv.second, (uint32_t)set_id.size()).first;
                if (it->second >= (1u << 15)) { err = "more than 32768 distinct symbol classes"; return RFB_E_UNSUPPORTED; }
                rec = (1ull << 16) | ((uint64_t)it->second << 17);
            }
            out.erec.push_back(rec | ((uint64_t)kv.first << 32));
        }
        out.eptr[s + 1] = (uint32_t)out.erec.size();
    }
    out.n_sets = (uint32_t)set_id.size();
    out.memb.assign(std::max<size_t>(1, set_id.size()) * 8, 0);
    for (auto &kv : set_id)
        for (uint32_t c = 0; c < 256; c++)
            if ((kv.first[c >> 6] >> (c & 63)) & 1) out.memb[kv.second * 8 + (c >> 5)] |= 1u << (c & 31);

    // exhaustive check: successors through the records == successors through the CSR, for every (state, symbol)
    std::vector<uint32_t> got, want;
    for (uint32_t s = 0; s < N; s++) {
        if ((out.eptr[s] == out.eptr[s + 1]) != (rp[s] == rp[s + 1])) { err = "edge-grouped CSR: accept mismatch at state " + std::to_string(s); return RFB_E_INTERNAL; }
        for (uint32_t c = 0; c < 256; c++) {
            want.clear(); got.clear();
            for (uint32_t j = rp[s]; j < rp[s + 1]; j++) if ((tr[j] >> 24) == c) want.push_back(tr[j] & 0xFFFFFFu);
            std::sort(want.begin(), want.end());
            want.erase(std::unique(want.begin(), want.end()), want.end());
            for (uint32_t j = out.eptr[s]; j < out.eptr[s + 1]; j++) {
                const uint64_t r = out.erec[j];
                const uint32_t lo = (uint32_t)r;
                const bool hit = (lo & 0x10000u) ? ((out.memb[(lo >> 17) * 8 + (c >> 5)] >> (c & 31)) & 1u) != 0
                                                 : (c == (lo & 0xFF) || c == ((lo >> 8) & 0xFF));
                if (hit) got.push_back((uint32_t)(r >> 32));
            }
            std::sort(got.begin(), got.end());
            if (got != want) { err = "edge-grouped CSR disagrees with the CSR at state " + std::to_string(s) + " symbol " + std::to_string(c); return RFB_E_INTERNAL; }
        }
    }
    return RFB_OK;
}

}  // namespace rfb
