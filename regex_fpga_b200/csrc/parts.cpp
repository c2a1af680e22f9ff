// parts.cpp -- cutting a large NFA into independently scannable parts; see host.h.
#include "host.h"
#include "../../include/regex_fpga_b200.h"
#include <algorithm>
#include <map>
#include <numeric>

namespace rfb {

void nfa_components(const Nfa &nfa, uint32_t max_states, std::vector<std::vector<uint32_t>> &groups) {
    groups.clear();
    const uint32_t N = nfa.n_states;
    const uint32_t *rp = nfa.row_ptr();
    const uint32_t *tr = nfa.trans();
    for (uint32_t j = 0; j < nfa.nnz; j++)
        if ((tr[j] & 0xFFFFFFu) == 0) return;            // something targets the start state: it is not a pure fan-out root
    std::vector<uint32_t> parent(N);
    std::iota(parent.begin(), parent.end(), 0u);
    auto find = [&](uint32_t x) { while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; } return x; };
    for (uint32_t s = 1; s < N; s++)
        for (uint32_t j = rp[s]; j < rp[s + 1]; j++) {
            const uint32_t a = find(s), b = find(tr[j] & 0xFFFFFFu);
            if (a != b) parent[std::max(a, b)] = std::min(a, b);
        }
    std::map<uint32_t, std::vector<uint32_t>> comp;      // keyed by smallest member: deterministic order
    for (uint32_t s = 1; s < N; s++) comp[find(s)].push_back(s);
    if (comp.size() < 2) return;
    std::vector<uint32_t> cur;
    for (auto &kv : comp) {
        if (!cur.empty() && cur.size() + kv.second.size() > max_states) { groups.push_back(cur); cur.clear(); }
        cur.insert(cur.end(), kv.second.begin(), kv.second.end());
    }
    if (!cur.empty()) groups.push_back(cur);
    for (auto &g : groups) std::sort(g.begin(), g.end());
    if (groups.size() < 2) groups.clear();
}

int nfa_extract(const Nfa &nfa, const std::vector<uint32_t> &states, Nfa &sub, std::vector<uint32_t> &to_orig, std::string &err) {
    const uint32_t *rp = nfa.row_ptr();
    const uint32_t *tr = nfa.trans();
    to_orig.assign(1, 0u);
    to_orig.insert(to_orig.end(), states.begin(), states.end());
    std::vector<uint32_t> sub_of(nfa.n_states, 0xFFFFFFFFu);
    for (uint32_t i = 0; i < to_orig.size(); i++) sub_of[to_orig[i]] = i;
    std::vector<uint32_t> nrp(1, 0u), ntr;
    for (uint32_t i = 0; i < to_orig.size(); i++) {
        const uint32_t s = to_orig[i];
        for (uint32_t j = rp[s]; j < rp[s + 1]; j++) {
            const uint32_t t = sub_of[tr[j] & 0xFFFFFFu];
            if (t == 0xFFFFFFFFu) {
                if (s == 0) continue;                     // the start state's edges into other parts
                err = "state " + std::to_string(s) + " has a transition that leaves its component";
                return RFB_E_INTERNAL;
            }
            ntr.push_back((tr[j] & 0xFF000000u) | t);
        }
        nrp.push_back((uint32_t)ntr.size());
    }
    std::vector<uint32_t> e(nrp);
    e.insert(e.end(), ntr.begin(), ntr.end());
    while (e.size() % 4) e.push_back(0);
    return nfa_from_entries(e.data(), e.size(), (int64_t)to_orig.size(), sub, err);
}

}  // namespace rfb
