// image.cpp -- load-time re-indexing of the CSR NFA into the lane kernel's execution image.
//
// The FPGA scans every state index per symbol (Design/FPGA.v:744-752) and streams whole CSR rows of
// the active ones (FPGA.v:227-407).  On the shipped rulesets 87% of the cycles are that idle scan and
// most of the rest is re-reading the 256+ entry rows of states that stay active for ever
// (SURVEY.md 3.2).  The image removes both without changing what is computed:
//
//   * "sticky" states (self-loop on >= sticky_min_self symbols) live in a per-stream bit mask P.
//     Per symbol c:  P' = (P & K[c]) | newly-entered;  the states in P & M[c] additionally fire their
//     non-self edges.  K[c] bit b = sticky state b self-loops on c;  M[c] bit b = it has a non-self
//     edge on c;  A[c] = ~K[c] | M[c] lets the kernel skip both when nothing happens.
//     LOOK-AHEAD: most firings are wasted -- the state entered dies on the very next symbol (on the shipped hi trace
//     86 % of them).  look[c'] bit b = "sticky state b has a non-self edge to a target that is sticky, accepting, or
//     has ANY edge on symbol c'".  A firing of b whose next symbol c' has the bit clear can only enter ordinary
//     states that have no successor on c' and are not accepting: they would be dropped one step later without a
//     report, so the kernel skips the firing (never for the last symbol of a stream: S_{n_steps} is observable).
//   * every state's (non-self, for sticky states) edges live in ONE table of 32-bit records `tab`:
//       - a state with a single (symbol-set -> target) edge is ONE record at tab[id];
//       - a branching state, and every sticky state, is a row of 2^bucket_bits records indexed by a
//         hash h(c) of the symbol: a branching state's row has 2^bucket_bits records at
//         tab[id + (h(c) & (2^bucket_bits - 1))]; the row of sticky bit b is sized to that state
//         (1..256 records, sdesc[b] = base | mask << 16) so that its firing costs one lookup;
//         a bucket holding several edges redirects to a contiguous chain; an empty bucket holds a
//         pair that cannot match in that bucket.
//     record = a[7:0] | b[15:8] | target_id[30:16] | more[31]
//       a <= b            : edge taken iff c == a or c == b
//       a == 0xFF > b     : indirect, target_id = index of the chain
//       a in {0xFE,0xFD}  : edge taken iff c is in class set (0xFE - a) * 253 + b   (b < 253)
//   * the always-active sticky state A (in an unanchored ruleset: the global ".*" state, which self-loops on all
//     256 symbols) gets bit 0 and a START DFA.  The states reachable from A through ordinary (neither sticky nor
//     accepting) states are never materialised while the DFA can follow them: a per-stream DFA state d stands for
//     the set members(d) of such states that is active, and one lookup DT[d][cls(c)] replaces the expansion of all
//     of them (the NFA step is a union over active states, so tracking part of the set as a subset-construction
//     state is exact).  Sticky, accepting and -- for states beyond the size budget -- ordinary successors that the
//     next DFA state does not contain are listed per transition (an "insertion list") and enter the per-stream
//     structures like any other target.  States beyond the budget resolve to the state of their longest tracked
//     suffix history (Aho-Corasick failure links over the subset states), so no transition is ever missing.
//     cmap[c] = class of c for the DFA | h(c) << 16.  The DFA tables live in global memory (L1/L2 resident).
//   * accepting (zero-out-degree) states own the contiguous id range [acc_base, acc_base + n_acc):
//     the kernel reports them when it pops them, without a table lookup.
//   * state ids are internal (sticky ids < 64 * sticky_words); orig_of_id restores the reference's
//     state numbers in every match record.
//
// image_successors() interprets these tables exactly as the kernel does and image_verify() compares
// it with the CSR for all n_states x 256 (state, symbol) pairs at load time.
#include "host.h"
#include "../../include/regex_fpga_b200.h"
#include <algorithm>
#include <array>
#include <cstring>
#include <map>
#include <unordered_map>

namespace rfb {
namespace {

struct SymSet {
    std::array<uint64_t, 4> w{{0, 0, 0, 0}};
    void set(uint32_t c) { w[c >> 6] |= 1ull << (c & 63); }
    bool has(uint32_t c) const { return (w[c >> 6] >> (c & 63)) & 1; }
    int count() const { return __builtin_popcountll(w[0]) + __builtin_popcountll(w[1]) + __builtin_popcountll(w[2]) + __builtin_popcountll(w[3]); }
    bool operator<(const SymSet &o) const { return w < o.w; }
    bool operator==(const SymSet &o) const { return w == o.w; }
    bool any() const { return w[0] | w[1] | w[2] | w[3]; }
    SymSet operator&(const SymSet &o) const { SymSet r; for (int i = 0; i < 4; i++) r.w[i] = w[i] & o.w[i]; return r; }
    std::vector<uint32_t> members() const { std::vector<uint32_t> v; for (uint32_t c = 0; c < 256; c++) if (has(c)) v.push_back(c); return v; }
};

struct Edge { uint32_t tgt; SymSet syms; };  // tgt = ORIGINAL state id

inline uint32_t align16(uint32_t x) { return (x + 15u) & ~15u; }

struct VecHash {
    size_t operator()(const std::vector<uint32_t> &v) const {
        uint64_t h = 0xcbf29ce484222325ull;
        for (uint32_t x : v) { h ^= x; h *= 0x100000001b3ull; }
        return (size_t)h;
    }
};

int image_build_one(const Nfa &nfa, const ImageOptions &opt, Image &img, std::string &err) {
    img = Image();
    const uint32_t N = nfa.n_states;
    const uint32_t *rp = nfa.row_ptr();
    const uint32_t *tr = nfa.trans();

    // ---- per-state edges grouped by target -----------------------------------------------------
    std::vector<std::vector<Edge>> edges(N);
    std::vector<SymSet> selfset(N);
    for (uint32_t s = 0; s < N; s++) {
        std::map<uint32_t, SymSet> by_tgt;
        for (uint32_t j = rp[s]; j < rp[s + 1]; j++) by_tgt[tr[j] & 0xFFFFFFu].set(tr[j] >> 24);
        for (auto &kv : by_tgt) {
            if (kv.first == s) selfset[s] = kv.second;
            edges[s].push_back(Edge{kv.first, kv.second});
        }
    }

    // ---- sticky selection ----------------------------------------------------------------------
    std::vector<uint32_t> cand;
    for (uint32_t s = 0; s < N; s++)
        if (selfset[s].count() >= opt.sticky_min_self && !std::binary_search(opt.not_sticky.begin(), opt.not_sticky.end(), s)) cand.push_back(s);
    std::stable_sort(cand.begin(), cand.end(), [&](uint32_t x, uint32_t y) { return selfset[x].count() > selfset[y].count(); });
    int W = opt.sticky_words;
    if (W != 1 && W != 2) W = cand.size() <= 64 ? 1 : 2;
    const uint32_t nsb = 64u * (uint32_t)W;
    if (cand.size() > nsb) { img.n_sticky_dropped = (uint32_t)(cand.size() - nsb); cand.resize(nsb); }
    std::vector<int32_t> sticky_bit(N, -1);
    for (size_t b = 0; b < cand.size(); b++) sticky_bit[cand[b]] = (int32_t)b;
    img.n_sticky = (uint32_t)cand.size();
    // a sticky state keeps its self loop in K; only its other edges go to its row
    for (uint32_t p : cand) {
        std::vector<Edge> keep;
        for (const Edge &e : edges[p]) if (e.tgt != p) keep.push_back(e);
        edges[p].swap(keep);
    }

    // ---- always-active sticky state ---------------------------------------------------------------------
    auto is_accept = [&](uint32_t s) { return rp[s] == rp[s + 1]; };
    std::vector<Edge> a_edges;       // A's non-self edges: followed by the start DFA, not by A's row
    bool accel = false;
    if (opt.accel && !cand.empty()) {
        // A must self-loop on all 256 symbols (once entered it never leaves the set).  Prefer the state entered
        // from state 0 on every symbol (the global ".*" of an unanchored ruleset), then the most edges.
        SymSet all256;
        for (uint32_t c = 0; c < 256; c++) all256.set(c);
        size_t best = 0; int best_n = 0;
        for (size_t b = 0; b < cand.size(); b++) {
            if (!(selfset[cand[b]] == all256) || edges[cand[b]].empty()) continue;
            int n = 1 + (int)edges[cand[b]].size();
            for (const Edge &e : edges[0]) if (e.tgt == cand[b] && e.syms == all256) n += 1000000;
            if (n > best_n) { best_n = n; best = b; }
        }
        if (best_n > 0) {
            std::swap(cand[0], cand[best]);
            for (size_t b = 0; b < cand.size(); b++) sticky_bit[cand[b]] = (int32_t)b;
            a_edges.swap(edges[cand[0]]);
            accel = true; img.accel_state = cand[0];
        }
    }

    // ---- class sets (more than two symbols) ------------------------------------------------------
    std::map<SymSet, uint32_t> set_id;
    auto class_of = [&](const SymSet &s) -> uint32_t {
        auto it = set_id.find(s);
        if (it != set_id.end()) return it->second;
        uint32_t id = (uint32_t)set_id.size();
        set_id.emplace(s, id);
        return id;
    };

    // ---- id assignment: [0,nsb) sticky | accepting | single-edge | hashed rows | sticky rows | chains ----
    auto is_single = [&](uint32_t s) { return edges[s].size() == 1; };
    img.id_of_orig.assign(N, 0xFFFFFFFFu);
    for (size_t b = 0; b < cand.size(); b++) img.id_of_orig[cand[b]] = (uint32_t)b;
    uint32_t next_id = nsb;
    const uint32_t acc_base = next_id;
    for (uint32_t s = 0; s < N; s++) if (sticky_bit[s] < 0 && is_accept(s)) img.id_of_orig[s] = next_id++;
    const uint32_t n_acc = next_id - acc_base;
    for (uint32_t s = 0; s < N; s++) if (sticky_bit[s] < 0 && !is_accept(s) && is_single(s)) img.id_of_orig[s] = next_id++;
    const uint32_t gbase = next_id;
    int bb = opt.bucket_bits;
    if (bb < 1) bb = 4;
    if (bb > 6) bb = 6;
    const uint32_t NB = 1u << bb;
    std::vector<uint32_t> branchers;
    for (uint32_t s = 0; s < N; s++)
        if (sticky_bit[s] < 0 && !is_accept(s) && !is_single(s)) { img.id_of_orig[s] = next_id; next_id += NB; branchers.push_back(s); }
    const uint32_t srow_base = next_id;   // sticky rows are sized once the hash is known

    // ---- start DFA --------------------------------------------------------------------------------------
    std::vector<uint32_t> cmap(256, 0);                       // per symbol: DFA class | h(c) << 16
    Image::Dfa &D = img.dfa;
    D = Image::Dfa();
    D.dt.assign(1, 0); D.dta.assign(1, 0); D.mem_ptr.assign(2, 0);
    if (accel) {
        auto ordinary = [&](uint32_t s) { return sticky_bit[s] < 0 && !is_accept(s); };
        // states the DFA can hold: reachable from A through ordinary states
        std::vector<char> in_r(N, 0);
        std::vector<uint32_t> stack;
        for (const Edge &e : a_edges) if (ordinary(e.tgt) && !in_r[e.tgt]) { in_r[e.tgt] = 1; stack.push_back(e.tgt); }
        while (!stack.empty()) {
            const uint32_t s = stack.back(); stack.pop_back();
            for (const Edge &e : edges[s]) if (ordinary(e.tgt) && !in_r[e.tgt]) { in_r[e.tgt] = 1; stack.push_back(e.tgt); }
        }
        // symbol classes: two symbols are equivalent iff no edge of A or of those states tells them apart
        std::vector<uint32_t> cls(256, 0);
        uint32_t ncls = 1;
        {
            std::map<SymSet, int> distinct;
            for (const Edge &e : a_edges) distinct.emplace(e.syms, 0);
            for (uint32_t s = 0; s < N; s++) if (in_r[s]) for (const Edge &e : edges[s]) distinct.emplace(e.syms, 0);
            for (const auto &kv : distinct) {
                std::map<std::pair<uint32_t, bool>, uint32_t> split;
                for (uint32_t c = 0; c < 256; c++) {
                    auto it = split.emplace(std::make_pair(cls[c], kv.first.has(c)), (uint32_t)split.size()).first;
                    cls[c] = it->second;
                }
                ncls = (uint32_t)split.size();
            }
        }
        std::vector<uint32_t> rep(ncls, 0xFFFFFFFFu);
        for (uint32_t c = 0; c < 256; c++) if (rep[cls[c]] == 0xFFFFFFFFu) rep[cls[c]] = c;
        for (uint32_t c = 0; c < 256; c++) cmap[c] = cls[c];

        // successors per (state the DFA can hold, class) and of A per class: the expansion below only concatenates
        std::vector<int32_t> r_index(N, -1);
        uint32_t n_r = 0;
        for (uint32_t s = 0; s < N; s++) if (in_r[s]) r_index[s] = (int32_t)n_r++;
        std::vector<std::vector<uint32_t>> succ_cls((size_t)n_r * ncls), a_cls(ncls);
        for (uint32_t q = 0; q < ncls; q++) for (const Edge &e : a_edges) if (e.syms.has(rep[q])) a_cls[q].push_back(e.tgt);
        for (uint32_t s = 0; s < N; s++) if (in_r[s])
            for (const Edge &e : edges[s]) for (uint32_t q = 0; q < ncls; q++) if (e.syms.has(rep[q])) succ_cls[(size_t)r_index[s] * ncls + q].push_back(e.tgt);

        std::vector<char> entered_sticky(N, 0);               // sticky states that appear in an insertion list
        const size_t ACT_CAP = 6u << 20;                      // insertion-list entries (12 MB)
        uint32_t budget = std::min<uint32_t>(std::max<uint32_t>(opt.dfa_max_states, ncls + 2), 32766);
        for (;; budget = std::max<uint32_t>(ncls + 2, budget / 2)) {
            // breadth-first subset construction; states keyed by their ordinary members (original ids, sorted)
            std::unordered_map<std::vector<uint32_t>, uint32_t, VecHash> id_of;
            id_of.reserve(budget * 2);
            std::vector<std::vector<uint32_t>> members(2);
            std::vector<uint32_t> fail(2, 1);
            std::map<std::vector<uint16_t>, uint32_t> act_of;
            id_of[{}] = 1;
            D.dt.assign((size_t)2 * ncls, 0); D.dta.assign((size_t)2 * ncls, 0);
            D.act.assign(1, 0);                                // index 0: unused
            D.n_frontier = 0;
            bool act_overflow = false;
            uint64_t work = 0;                                 // edge-list visits so far: bounds the load time on NFAs whose subsets explode
            const uint64_t WORK_CAP = 60ull * 1000 * 1000;
            std::vector<uint32_t> T, ord;
            std::vector<uint16_t> lst;
            bool probe_abort = false;                          // probe: the first subset beyond the budget settles the answer
            for (uint32_t d = 1; d < members.size() && !act_overflow && !probe_abort; d++) {
                if (D.dt.size() < (size_t)(d + 1) * ncls) { D.dt.resize((size_t)(d + 1) * ncls, 0); D.dta.resize((size_t)(d + 1) * ncls, 0); }
                bool fell_back = false;
                work += (uint64_t)ncls * (members[d].size() + 1);
                for (uint32_t q = 0; q < ncls; q++) {
                    T = a_cls[q];
                    for (uint32_t m : members[d]) { const auto &v = succ_cls[(size_t)r_index[m] * ncls + q]; T.insert(T.end(), v.begin(), v.end()); }
                    std::sort(T.begin(), T.end());
                    T.erase(std::unique(T.begin(), T.end()), T.end());
                    ord.clear(); lst.clear();
                    for (uint32_t t : T) {
                        if (ordinary(t)) ord.push_back(t);
                        else { lst.push_back((uint16_t)img.id_of_orig[t]); if (sticky_bit[t] > 0) entered_sticky[t] = 1; }
                    }
                    uint32_t nd;
                    auto it = id_of.find(ord);
                    if (it != id_of.end()) nd = it->second;
                    else if (members.size() < budget && work < WORK_CAP) {
                        nd = (uint32_t)members.size();
                        id_of.emplace(ord, nd);
                        members.push_back(ord);
                        fail.push_back(d == 1 ? 1u : (uint32_t)(D.dt[(size_t)fail[d] * ncls + q] & 0x7FFFu));
                    } else {
                        // beyond the budget: continue from the longest tracked suffix history (its state holds a
                        // subset of ord) and insert the rest explicitly
                        fell_back = true;
                        if (opt.probe_dfa_only) { probe_abort = true; break; }
                        nd = d == 1 ? 1u : (uint32_t)(D.dt[(size_t)fail[d] * ncls + q] & 0x7FFFu);
                        if (!std::includes(ord.begin(), ord.end(), members[nd].begin(), members[nd].end())) nd = 1;
                        for (uint32_t t : ord) if (!std::binary_search(members[nd].begin(), members[nd].end(), t)) lst.push_back((uint16_t)img.id_of_orig[t]);
                    }
                    uint32_t at = 0;
                    if (!lst.empty()) {
                        std::sort(lst.begin(), lst.end());
                        auto jt = act_of.find(lst);
                        if (jt != act_of.end()) at = jt->second;
                        else {
                            at = (uint32_t)D.act.size();
                            act_of.emplace(lst, at);
                            for (size_t k = 0; k < lst.size(); k++) D.act.push_back((uint16_t)(lst[k] | (k + 1 < lst.size() ? 0x8000u : 0u)));
                            if (D.act.size() > ACT_CAP) act_overflow = true;
                        }
                    }
                    D.dt[(size_t)d * ncls + q] = (uint16_t)(nd | (at ? 0x8000u : 0u));
                    D.dta[(size_t)d * ncls + q] = at;
                }
                D.n_frontier += fell_back;
            }
            if (act_overflow && budget > ncls + 2) continue;
            if (act_overflow) { img.why_not = "start DFA insertion lists too large"; }
            D.ncls = ncls; D.n = (uint32_t)members.size();
            D.mem_ptr.assign(1, 0); D.mem_ids.clear();
            for (const auto &m : members) {
                for (uint32_t t : m) D.mem_ids.push_back((uint16_t)img.id_of_orig[t]);
                D.mem_ptr.push_back((uint32_t)D.mem_ids.size());
            }
            break;
        }
        for (uint32_t t = 0; t < N; t++) if (entered_sticky[t]) {
            SymSet fire;
            for (const Edge &e : edges[t]) for (int w = 0; w < 4; w++) fire.w[w] |= e.syms.w[w];
            img.dfa_sticky_targets.emplace_back(-fire.count(), t);
        }
        std::sort(img.dfa_sticky_targets.begin(), img.dfa_sticky_targets.end());
    }

    if (opt.probe_dfa_only) { img.n_absorbed = (uint32_t)opt.not_sticky.size(); img.h.accel = accel ? 1u : 0u; img.ok = img.why_not.empty(); return RFB_OK; }

    // ---- bucket hash: pick (mul, shift) minimising the expected table lookups per visit -------------
    // A visit with symbol c costs 1 lookup when bucket(c) holds <= 1 edge, 1 + n when it holds n >= 2
    // (indirection + chain).  Symbols that appear on some edge of the state are what the traffic that
    // activated the state tends to continue with, so they carry half the weight; all others the rest.
    auto hfull = [&](uint32_t c, uint32_t mul, uint32_t sh) { return ((c * mul) >> sh) & 0xFFu; };
    auto bucket_of = [&](uint32_t c, uint32_t mul, uint32_t sh) { return hfull(c, mul, sh) & (NB - 1); };
    uint32_t best_mul = 1, best_sh = 0;
    double best_cost = 1e300;
    if (opt.fixed_hash_mul) { best_mul = opt.fixed_hash_mul; best_sh = opt.fixed_hash_shift; }
    else for (uint32_t mul = 1; mul < 64; mul += 2)
        for (uint32_t sh = 0; sh < 8; sh++) {
            bool seen8[256] = {false}, bij = true;      // h must be a bijection so that 256-slot rows are direct
            for (uint32_t c = 0; c < 256 && bij; c++) { uint32_t v = hfull(c, mul, sh); bij = !seen8[v]; seen8[v] = true; }
            if (!bij) continue;
            double cost = 0;
            for (uint32_t s : branchers) {
                if (cost >= best_cost) break;
                if (edges[s].empty()) continue;
                uint32_t per_bucket[64] = {0};
                SymSet used;
                for (const Edge &e : edges[s]) {
                    uint64_t seen = 0;
                    for (uint32_t c : e.syms.members()) { seen |= 1ull << bucket_of(c, mul, sh); used.set(c); }
                    for (uint32_t q = 0; q < NB; q++) per_bucket[q] += (seen >> q) & 1;
                }
                const int n_used = used.count();
                const double w_used = 1.0 / n_used, w_other = n_used < 256 ? 1.0 / (256 - n_used) : 0.0;
                for (uint32_t c = 0; c < 256; c++) {
                    const uint32_t n = per_bucket[bucket_of(c, mul, sh)];
                    cost += (used.has(c) ? w_used : w_other) * (n <= 1 ? 1.0 : 1.0 + n);
                }
            }
            if (cost < best_cost) { best_cost = cost; best_mul = mul; best_sh = sh; }
        }

    // ---- sticky rows: the smallest power-of-two row in which no two edges share a bucket (<= 256) ----
    std::vector<uint32_t> srow_bits(nsb, 0), srow_off(nsb, 0);
    {
        uint32_t off = srow_base;
        for (uint32_t b = 0; b < nsb; b++) {
            uint32_t bits = 0;
            if (b < cand.size()) {
                for (; bits < 8; bits++) {
                    bool clash = false;
                    std::vector<int> owner(1u << bits, -1);
                    for (size_t ei = 0; ei < edges[cand[b]].size() && !clash; ei++)
                        for (uint32_t c : edges[cand[b]][ei].syms.members()) {
                            int &o = owner[hfull(c, best_mul, best_sh) & ((1u << bits) - 1)];
                            if (o >= 0 && o != (int)ei) { clash = true; break; }
                            o = (int)ei;
                        }
                    if (!clash) break;
                }
            }
            srow_bits[b] = bits; srow_off[b] = off; off += 1u << bits;
        }
        next_id = off;
    }
    const uint32_t chain_base = next_id;

    // ---- fill tab ----------------------------------------------------------------------------------
    bool ids_ok = true;
    std::vector<uint32_t> tab(chain_base, tab_pack(0, 0, 0, false));
    auto record_for = [&](const Edge &e, const SymSet &visible, bool more) -> uint32_t {
        // `visible` = the symbols that can reach this record; inside it the edge must fire iff c in e.syms
        SymSet eff = e.syms & visible;
        uint32_t tid = img.id_of_orig[e.tgt];
        if (tid > 0x7FFF) ids_ok = false;
        auto m = eff.members();
        if (m.size() <= 2 && !m.empty()) return tab_pack(m[0], m.back(), tid, more);
        uint32_t n = class_of(e.syms);
        uint32_t a = 0xFE - n / 253, b = n % 253;
        return tab_pack(a, b, tid, more);
    };
    SymSet all;
    for (uint32_t c = 0; c < 256; c++) all.set(c);
    auto fill_row = [&](uint32_t row, uint32_t bits, const std::vector<Edge> &ed) {
        const uint32_t nb = 1u << bits;
        for (uint32_t q = 0; q < nb; q++) {
            SymSet vis;                       // symbols that reach bucket q of this row
            int never = -1;                   // a symbol that cannot (filler for an empty bucket)
            for (uint32_t c = 0; c < 256; c++) { if ((hfull(c, best_mul, best_sh) & (nb - 1)) == q) vis.set(c); else if (never < 0) never = (int)c; }
            std::vector<const Edge *> in;
            for (const Edge &e : ed) if ((e.syms & vis).any()) in.push_back(&e);
            if (in.empty()) {
                if (never >= 0) tab[row + q] = tab_pack((uint32_t)never, (uint32_t)never, 0, false);
                else { uint32_t n = class_of(SymSet()); tab[row + q] = tab_pack(0xFE - n / 253, n % 253, 0, false); }  // 1-slot row: empty class
                continue;
            }
            if (in.size() == 1) { tab[row + q] = record_for(*in[0], vis, false); continue; }
            uint32_t start = (uint32_t)tab.size();
            if (start > 0x7FFF) ids_ok = false;
            tab[row + q] = tab_special(CODE_INDIRECT, start, false);
            for (size_t k = 0; k < in.size(); k++) tab.push_back(record_for(*in[k], vis, k + 1 < in.size()));
        }
    };
    for (uint32_t s = 0; s < N; s++) {
        if (sticky_bit[s] >= 0) { fill_row(srow_off[sticky_bit[s]], srow_bits[sticky_bit[s]], edges[s]); continue; }
        const uint32_t id = img.id_of_orig[s];
        if (is_accept(s)) continue;                     // reported by id range, never looked up
        if (is_single(s)) tab[id] = record_for(edges[s][0], all, false);
        else fill_row(id, (uint32_t)bb, edges[s]);
    }
    for (uint32_t b = (uint32_t)cand.size(); b < nsb; b++) fill_row(srow_off[b], 0, {});
    std::vector<uint32_t> sdesc(nsb);
    for (uint32_t b = 0; b < nsb; b++) {
        if (srow_off[b] > 0xFFFF) ids_ok = false;
        sdesc[b] = srow_off[b] | (((1u << srow_bits[b]) - 1) << 16);
    }

    // ---- sticky masks ----------------------------------------------------------------------------------
    const uint32_t mstride = 32u * (uint32_t)W;  // bytes per symbol: A[W] (padded to 16) | K[W] | M[W]
    std::vector<uint8_t> mask(256 * mstride, 0);
    for (size_t b = 0; b < nsb; b++) {
        for (uint32_t c = 0; c < 256; c++) {
            uint64_t *A = reinterpret_cast<uint64_t *>(&mask[c * mstride]);
            uint64_t *K = reinterpret_cast<uint64_t *>(&mask[c * mstride + 16]);
            uint64_t *M = K + W;
            const uint64_t bit = 1ull << (b & 63);
            if (b >= cand.size()) { K[b >> 6] |= bit; continue; }   // absent slot: inert
            const uint32_t p = cand[b];
            bool fires = false;
            for (const Edge &e : edges[p]) fires = fires || e.syms.has(c);
            if (selfset[p].has(c)) K[b >> 6] |= bit; else A[b >> 6] |= bit;
            if (fires) { M[b >> 6] |= bit; A[b >> 6] |= bit; }
        }
    }

    // ---- look-ahead masks (see the header comment) -----------------------------------------------------
    std::vector<uint64_t> look(256 * (size_t)W, 0);
    {
        std::vector<SymSet> out_syms(N);                       // symbols on which a state has any CSR transition
        for (uint32_t s2 = 0; s2 < N; s2++) for (uint32_t j = rp[s2]; j < rp[s2 + 1]; j++) out_syms[s2].set(tr[j] >> 24);
        for (size_t b = 0; b < cand.size(); b++) {
            SymSet alive;                                      // next symbols under which some target of b's edges matters
            bool always = false;
            for (const Edge &e : edges[cand[b]]) {
                if (sticky_bit[e.tgt] >= 0 || is_accept(e.tgt)) { always = true; break; }
                for (int w = 0; w < 4; w++) alive.w[w] |= out_syms[e.tgt].w[w];
            }
            for (uint32_t c = 0; c < 256; c++) if (always || alive.has(c)) look[c * W + (b >> 6)] |= 1ull << (b & 63);
        }
    }

    for (uint32_t c = 0; c < 256; c++) cmap[c] = (cmap[c] & 0xFFFFu) | (hfull(c, best_mul, best_sh) << 16);
    // the lane kernel reads the attention mask and the descriptor of a symbol with ONE 16-byte load (W == 1) / from one
    // row (W == 2): {class, hash} of cmap[c] as two 32-bit words in the padding of symbol c's mask row
    for (uint32_t c = 0; c < 256; c++) {
        const uint32_t cls_hf[2] = {cmap[c] & 0xFFFFu, cmap[c] >> 16};
        std::memcpy(&mask[c * mstride + (W == 1 ? 8u : 48u)], cls_hf, 8);
    }

    // ---- class membership bitmaps ---------------------------------------------------------------------
    std::vector<uint32_t> memb(std::max<size_t>(1, set_id.size()) * 8, 0);
    for (auto &kv : set_id)
        for (uint32_t c = 0; c < 256; c++)
            if (kv.first.has(c)) memb[kv.second * 8 + (c >> 5)] |= 1u << (c & 31);

    // ---- feasibility -------------------------------------------------------------------------------------
    if (tab.size() > 0x8000) ids_ok = false;
    if (set_id.size() > 506) img.why_not = "more than 506 distinct symbol classes";
    if (!ids_ok && img.why_not.empty()) img.why_not = "edge table exceeds the 15-bit id space (" + std::to_string(tab.size()) + " slots)";

    // ---- blob ------------------------------------------------------------------------------------------------
    ImageHeader &h = img.h;
    h.n_slots = (uint32_t)tab.size();
    h.gbase = gbase;
    h.nsb = nsb;
    h.sticky_words = (uint32_t)W;
    h.bucket_bits = (uint32_t)bb;
    h.hash_mul = best_mul;
    h.hash_shift = best_sh;
    h.start_id = img.id_of_orig[0];
    h.n_sets = (uint32_t)set_id.size();
    h.acc_base = acc_base;
    h.n_acc = n_acc;
    h.srow_base = srow_base;
    // Layout: the fixed-size tables first, at offsets that depend only on the sticky word count, so that the kernel
    // addresses them with immediates: mask | cmap | sdesc | tab | memb
    uint32_t off = 0;
    h.off_mask = off;  off += 256u * 32u * (uint32_t)W;
    h.off_cmap = off;  off += 1024;
    h.off_sdesc = off; off += nsb * 4;
    h.off_tab = off;   off = align16(off + (uint32_t)tab.size() * 4);
    h.accel = accel ? 1u : 0u;
    h.dfa_ncls = D.ncls;
    h.dfa_states = D.n;
    h.off_memb = off;  off = align16(off + (uint32_t)memb.size() * 4);
    h.off_look = off;  off = align16(off + (uint32_t)look.size() * 8);
    h.blob_bytes = off;
    if (img.why_not.empty() && off > opt.max_bytes) img.why_not = "tables need " + std::to_string(off) + " bytes of shared memory (limit " + std::to_string(opt.max_bytes) + ")";
    img.blob.assign(off, 0);
    std::memcpy(&img.blob[h.off_tab], tab.data(), tab.size() * 4);
    std::memcpy(&img.blob[h.off_mask], mask.data(), mask.size());
    std::memcpy(&img.blob[h.off_memb], memb.data(), memb.size() * 4);
    std::memcpy(&img.blob[h.off_sdesc], sdesc.data(), sdesc.size() * 4);
    std::memcpy(&img.blob[h.off_cmap], cmap.data(), 1024);
    std::memcpy(&img.blob[h.off_look], look.data(), look.size() * 8);

    img.orig_of_id.assign(tab.size(), 0xFFFFFFFFu);
    for (uint32_t s = 0; s < N; s++) if (img.id_of_orig[s] < img.orig_of_id.size()) img.orig_of_id[img.id_of_orig[s]] = s;
    img.n_absorbed = (uint32_t)opt.not_sticky.size();
    img.ok = img.why_not.empty();
    if (img.ok && opt.verify) {
        int rc = image_verify(nfa, img, err);
        if (rc) { img.ok = false; return rc; }
    }
    return RFB_OK;
}

}  // namespace

// bucket_bits < 1: the most buckets (up to 16 per hashed row) whose tables still fit.
static int image_build_bits(const Nfa &nfa, const ImageOptions &opt, Image &img, std::string &err) {
    if (opt.bucket_bits >= 1) return image_build_one(nfa, opt, img, err);
    ImageOptions o = opt;
    int rc = RFB_OK;
    for (int bb = 4; bb >= 1; bb--) {
        o.bucket_bits = bb;
        rc = image_build_one(nfa, o, img, err);
        if (rc != RFB_OK || img.ok) return rc;
    }
    return rc;
}

// Which self-looping states live in the mask and which in the start DFA.  A sticky state costs the lane kernel an
// explicit row lookup every time it fires and an explicit ring entry for every successor, while a state that the start
// DFA tracks costs nothing.  A self-looping state that the DFA enters directly (it appears in an insertion list) can
// just as well be an ordinary member of the DFA's subsets: its self loop keeps it in the subset, its other edges are
// followed by the DFA.  That multiplies the reachable subsets, so states are moved greedily -- those firing on the
// most symbols first -- and a move is kept only while the DFA stays COMPLETE within its budget (no failure-link rows)
// and the tables still fit.  On snort_16 three states move (the DFA grows from 8 495 to 16 434 states), the explicit
// lookups per symbol of the hi trace windows drop from 0.56 to 0.17, and the mask shrinks to one 64-bit word.
int image_build(const Nfa &nfa, const ImageOptions &opt_in, Image &img, std::string &err) {
    ImageOptions opt = opt_in;
    opt.not_sticky.clear();
    opt.verify = true;
    int rc = image_build_bits(nfa, opt, img, err);
    if (rc != RFB_OK || !img.ok || !img.h.accel || opt.dfa_absorb <= 0 || img.dfa.n_frontier != 0) return rc;
    ImageOptions trial_opt = opt;
    trial_opt.verify = false;
    trial_opt.probe_dfa_only = true;
    trial_opt.bucket_bits = (int)img.h.bucket_bits;
    trial_opt.fixed_hash_mul = img.h.hash_mul; trial_opt.fixed_hash_shift = img.h.hash_shift;   // the search is the slow part of a build
    std::vector<uint32_t> rejected;
    int trials = 0, kept = 0;
    Image best = img;                                   // verified
    bool best_verified = true;
    const int max_trials = opt.dfa_absorb + 6;
    for (bool progress = true; progress && trials < max_trials && kept < opt.dfa_absorb;) {
        progress = false;
        // sticky states the current DFA enters, by the number of symbols on which they fire
        std::vector<std::pair<int, uint32_t>> cands;
        for (const auto &cd : best.dfa_sticky_targets)
            if (-cd.first >= 4 && std::find(rejected.begin(), rejected.end(), cd.second) == rejected.end()) cands.push_back(cd);
        std::sort(cands.begin(), cands.end());
        for (const auto &cd : cands) {
            if (trials >= max_trials || kept >= opt.dfa_absorb) break;
            trials++;
            ImageOptions t = trial_opt;
            t.not_sticky = opt.not_sticky;
            t.not_sticky.push_back(cd.second);
            std::sort(t.not_sticky.begin(), t.not_sticky.end());
            Image cand_img;
            std::string e2;
            const int r2 = image_build_one(nfa, t, cand_img, e2);
            if (r2 == RFB_OK && cand_img.ok && cand_img.h.accel && cand_img.dfa.n_frontier == 0) {
                opt.not_sticky = t.not_sticky;
                best = std::move(cand_img);
                best_verified = false;
                kept++;
                progress = true;
                break;                                  // the candidate list changes with the DFA: recompute it
            }
            rejected.push_back(cd.second);
        }
    }
    if (!best_verified) {   // the final choice goes through the full proof like any image
        opt.bucket_bits = (int)img.h.bucket_bits;
        opt.fixed_hash_mul = img.h.hash_mul; opt.fixed_hash_shift = img.h.hash_shift;
        Image final_img;
        rc = image_build_one(nfa, opt, final_img, err);
        if (rc != RFB_OK || !final_img.ok) { err.clear(); return RFB_OK; }   // keep the verified baseline in img
        img = std::move(final_img);
    }
    return RFB_OK;
}

// The kernel's semantics, on the host.  Keep in lock-step with scan_lane_kernel (scan.cu).
void image_successors(const Image &img, uint32_t s, uint32_t c, std::vector<uint32_t> &out, bool *accepting) {
    const ImageHeader &h = img.h;
    const uint32_t *tab = reinterpret_cast<const uint32_t *>(&img.blob[h.off_tab]);
    const uint32_t *memb = reinterpret_cast<const uint32_t *>(&img.blob[h.off_memb]);
    const uint32_t mstride = 32u * h.sticky_words;
    const uint32_t hf = ((c * h.hash_mul) >> h.hash_shift) & 0xFFu;
    const uint32_t hc = hf & ((1u << h.bucket_bits) - 1);
    const uint32_t *sdesc = reinterpret_cast<const uint32_t *>(&img.blob[h.off_sdesc]);
    out.clear();
    if (accepting) *accepting = false;
    const uint32_t id = img.id_of_orig[s];
    std::vector<uint32_t> ids;
    uint32_t idx = 0;
    bool walk = false;
    if (id < h.nsb) {
        const uint64_t *K = reinterpret_cast<const uint64_t *>(&img.blob[h.off_mask + c * mstride + 16]);
        const uint64_t *M = K + h.sticky_words;
        const uint64_t bit = 1ull << (id & 63);
        if (K[id >> 6] & bit) ids.push_back(id);
        if (M[id >> 6] & bit) { idx = (sdesc[id] & 0xFFFFu) + (hf & (sdesc[id] >> 16)); walk = true; }
        if (h.accel && id == 0) {   // A's edges are followed by the start DFA: row of DFA state 1 (A alone)
            const uint32_t *cmap = reinterpret_cast<const uint32_t *>(&img.blob[h.off_cmap]);
            const Image::Dfa &D = img.dfa;
            const size_t at = (size_t)1 * D.ncls + (cmap[c] & 0xFF);
            const uint32_t nd = D.dt[at] & 0x7FFFu;
            for (uint32_t j = D.mem_ptr[nd]; j < D.mem_ptr[nd + 1]; j++) ids.push_back(D.mem_ids[j]);
            if (D.dt[at] & 0x8000u) for (uint32_t q = D.dta[at];; q++) { ids.push_back(D.act[q] & 0x7FFFu); if (!(D.act[q] & 0x8000u)) break; }
        }
    } else if (id - h.acc_base < h.n_acc) {
        if (accepting) *accepting = true;
    } else {
        idx = id + (id >= h.gbase ? hc : 0u);
        walk = true;
    }
    while (walk) {
        const uint32_t e = tab[idx];
        const uint32_t a = e & 0xFF, b = (e >> 8) & 0xFF, t = (e >> 16) & 0x7FFF;
        if (a <= b) { if (c == a || c == b) ids.push_back(t); }
        else if (a == 0xFF) { idx = t; continue; }
        else {
            const uint32_t n = (0xFE - a) * 253 + b;
            if ((memb[n * 8 + (c >> 5)] >> (c & 31)) & 1) ids.push_back(t);
        }
        if (!(e & TAB_MORE)) break;
        idx++;
    }
    for (uint32_t i : ids) out.push_back(i < img.orig_of_id.size() ? img.orig_of_id[i] : 0xFFFFFFFFu);
    std::sort(out.begin(), out.end());
    out.erase(std::unique(out.begin(), out.end()), out.end());
}

int image_verify(const Nfa &nfa, const Image &img, std::string &err) {
    const uint32_t N = nfa.n_states;
    const uint32_t *rp = nfa.row_ptr();
    const uint32_t *tr = nfa.trans();
    std::vector<uint32_t> got, want;
    std::vector<std::vector<uint32_t>> by_sym(256);
    for (uint32_t s = 0; s < N; s++) {
        for (auto &v : by_sym) v.clear();
        for (uint32_t j = rp[s]; j < rp[s + 1]; j++) by_sym[tr[j] >> 24].push_back(tr[j] & 0xFFFFFFu);
        for (uint32_t c = 0; c < 256; c++) {
            want = by_sym[c];
            std::sort(want.begin(), want.end());
            want.erase(std::unique(want.begin(), want.end()), want.end());
            bool acc = false;
            image_successors(img, s, c, got, &acc);
            if (got != want || acc != (rp[s] == rp[s + 1])) {
                err = "execution image disagrees with the CSR at state " + std::to_string(s) + " symbol " + std::to_string(c);
                return RFB_E_INTERNAL;
            }
        }
    }
    {   // what the kernel reads but image_successors() does not: the "attention" masks and the per-symbol hash
        const ImageHeader &h = img.h;
        const uint32_t W = h.sticky_words, mstride = 32u * W;
        const uint32_t *cmap = reinterpret_cast<const uint32_t *>(&img.blob[h.off_cmap]);
        for (uint32_t c = 0; c < 256; c++) {
            const uint64_t *A = reinterpret_cast<const uint64_t *>(&img.blob[h.off_mask + c * mstride]);
            const uint64_t *K = reinterpret_cast<const uint64_t *>(&img.blob[h.off_mask + c * mstride + 16]);
            const uint64_t *M = K + W;
            for (uint32_t w = 0; w < W; w++)
                if (A[w] != (~K[w] | M[w])) { err = "sticky attention mask disagrees with K and M at symbol " + std::to_string(c); return RFB_E_INTERNAL; }
            if ((cmap[c] >> 16) != (((c * h.hash_mul) >> h.hash_shift) & 0xFFu)) { err = "per-symbol hash disagrees with the header at symbol " + std::to_string(c); return RFB_E_INTERNAL; }
            uint32_t cm_copy[2];
            std::memcpy(cm_copy, &img.blob[h.off_mask + c * mstride + (W == 1 ? 8u : 48u)], 8);
            if (cm_copy[0] != (cmap[c] & 0xFFFFu) || cm_copy[1] != (cmap[c] >> 16)) { err = "symbol descriptor copy in the mask row disagrees with cmap at symbol " + std::to_string(c); return RFB_E_INTERNAL; }
            if ((cmap[c] & 0xFFFFu) >= std::max<uint32_t>(1u, img.dfa.ncls)) { err = "symbol class out of range at symbol " + std::to_string(c); return RFB_E_INTERNAL; }
        }
    }
    {   // look-ahead masks: a clear bit (b, c') promises that no non-self edge of sticky state b leads to a state that is
        // sticky, accepting, or has any transition on c' (the kernel then skips b's firing when c' is the next symbol)
        const ImageHeader &h = img.h;
        const uint32_t W = h.sticky_words;
        const uint64_t *look = reinterpret_cast<const uint64_t *>(&img.blob[h.off_look]);
        std::vector<std::array<uint64_t, 4>> out_syms(N, std::array<uint64_t, 4>{{0, 0, 0, 0}});
        for (uint32_t s = 0; s < N; s++) for (uint32_t j = rp[s]; j < rp[s + 1]; j++) out_syms[s][(tr[j] >> 24) >> 6] |= 1ull << ((tr[j] >> 24) & 63);
        for (uint32_t b = 0; b < h.nsb; b++) {
            if (b >= img.orig_of_id.size() || img.orig_of_id[b] == 0xFFFFFFFFu) continue;     // unused bit: never set in P
            if (h.accel && b == 0) continue;                                                   // A's edges are followed by the start DFA, A never fires
            const uint32_t p = img.orig_of_id[b];
            for (uint32_t j = rp[p]; j < rp[p + 1]; j++) {
                const uint32_t t = tr[j] & 0xFFFFFFu;
                if (t == p) continue;
                const bool matters_always = img.id_of_orig[t] < h.nsb || rp[t] == rp[t + 1];
                for (uint32_t c = 0; c < 256; c++) {
                    if ((look[c * W + (b >> 6)] >> (b & 63)) & 1) continue;
                    if (matters_always || ((out_syms[t][c >> 6] >> (c & 63)) & 1)) {
                        err = "look-ahead mask would drop a live successor of state " + std::to_string(p) + " (next symbol " + std::to_string(c) + ")";
                        return RFB_E_INTERNAL;
                    }
                }
            }
        }
    }
    if (img.h.accel) {
        // Start DFA: for every DFA state d >= 1 and every symbol c,
        //   members(next) + insertion list  ==  successors on c of members(d) and of A (A's self loop aside),
        // members are ordinary states, and state 0 (A not active yet) is inert.
        const ImageHeader &h = img.h;
        const Image::Dfa &D = img.dfa;
        const uint32_t *cmap = reinterpret_cast<const uint32_t *>(&img.blob[h.off_cmap]);
        const uint32_t A = img.accel_state;
        if (img.id_of_orig[A] != 0 || D.n < 2 || D.dt.size() != (size_t)D.n * D.ncls || D.mem_ptr.size() != D.n + 1) { err = "start DFA is malformed"; return RFB_E_INTERNAL; }
        for (uint32_t q = 0; q < D.ncls; q++) if (D.dt[q] != 0) { err = "start DFA state 0 is not inert"; return RFB_E_INTERNAL; }
        if (D.mem_ptr[1] != D.mem_ptr[2]) { err = "start DFA state 1 is not empty"; return RFB_E_INTERNAL; }
        for (uint16_t m : D.mem_ids)
            if (m < h.nsb || m - h.acc_base < h.n_acc) { err = "start DFA hides a sticky or accepting state"; return RFB_E_INTERNAL; }
        // successor lists per (state, symbol) of the states the DFA can hold, from the CSR
        std::vector<std::vector<uint32_t>> rows;              // rows[k * 256 + c]
        std::vector<int32_t> row_of(N, -1);
        auto row = [&](uint32_t s) -> size_t {
            if (row_of[s] < 0) {
                row_of[s] = (int32_t)(rows.size() / 256);
                rows.resize(rows.size() + 256);
                for (uint32_t j = rp[s]; j < rp[s + 1]; j++) {
                    const uint32_t t = tr[j] & 0xFFFFFFu;
                    if (!(s == A && t == A)) rows[(size_t)row_of[s] * 256 + (tr[j] >> 24)].push_back(t);
                }
            }
            return (size_t)row_of[s] * 256;
        };
        for (uint32_t d = 1; d < D.n; d++) {
            for (uint32_t c = 0; c < 256; c++) {
                want.clear();
                { const size_t r = row(A) + c; want.insert(want.end(), rows[r].begin(), rows[r].end()); }
                for (uint32_t j = D.mem_ptr[d]; j < D.mem_ptr[d + 1]; j++) {
                    const size_t r = row(img.orig_of_id[D.mem_ids[j]]) + c;
                    want.insert(want.end(), rows[r].begin(), rows[r].end());
                }
                std::sort(want.begin(), want.end());
                want.erase(std::unique(want.begin(), want.end()), want.end());
                got.clear();
                const size_t at = (size_t)d * D.ncls + (cmap[c] & 0xFF);
                const uint32_t nd = D.dt[at] & 0x7FFFu;
                if (nd == 0 || nd >= D.n) { err = "start DFA transition leaves the table"; return RFB_E_INTERNAL; }
                for (uint32_t j = D.mem_ptr[nd]; j < D.mem_ptr[nd + 1]; j++) got.push_back(img.orig_of_id[D.mem_ids[j]]);
                if (D.dt[at] & 0x8000u)
                    for (uint32_t q = D.dta[at];; q++) {
                        if (q == 0 || q >= D.act.size()) { err = "start DFA insertion list out of range"; return RFB_E_INTERNAL; }
                        got.push_back(img.orig_of_id[D.act[q] & 0x7FFFu]);
                        if (!(D.act[q] & 0x8000u)) break;
                    }
                std::sort(got.begin(), got.end());
                if (got != want) { err = "start DFA disagrees with the CSR at DFA state " + std::to_string(d) + " symbol " + std::to_string(c); return RFB_E_INTERNAL; }
            }
        }
    }
    return RFB_OK;
}

}  // namespace rfb
