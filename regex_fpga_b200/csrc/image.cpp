// image.cpp -- load-time re-indexing of the CSR NFA into the lane kernel's execution image.
//
// The FPGA scans every state index per symbol (Design/FPGA.v:744-752) and streams whole CSR rows of
// the active ones (FPGA.v:227-407).  On the shipped rulesets 87% of the cycles are that idle scan and
// most of the rest is re-reading the 256+ entry rows of states that stay active for ever
// (SURVEY.md 3.2).  The image removes both without changing what is computed:
//
//   * "sticky" states (self-loop on >= sticky_min_self symbols) live in a per-stream bit mask P.
//     Per symbol c:  P' = (P & K[c]) | newly-entered;  injections = targets of (P & M[c]).
//     K[c] bit b = sticky state b self-loops on c;  M[c] bit b = it has a non-self edge on c;
//     A[c] = ~K[c] | M[c] lets the kernel skip both when nothing happens.
//   * every other state is an id into one table of 32-bit edge records `tab`:
//       - a state with a single (symbol-set -> target) edge or no edge is ONE record;
//       - a branching state is a row of 2^bucket_bits records indexed by a hash of the symbol;
//         a bucket holding several edges redirects to a contiguous chain.
//     record = a[7:0] | b[15:8] | target_id[30:16] | more[31]
//       a <= b : edge taken iff c == a or c == b
//       a == 0xFF > b : b = 0 empty, 1 accepting state, 2 indirect (target_id = chain start)
//       a in {0xFE,0xFD} > b : edge taken iff c is in class set (0xFE - a) * 253 + b
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

}  // namespace

static int image_build_one(const Nfa &nfa, const ImageOptions &opt, Image &img, std::string &err) {
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
        if (selfset[s].count() >= opt.sticky_min_self) cand.push_back(s);
    std::stable_sort(cand.begin(), cand.end(), [&](uint32_t x, uint32_t y) { return selfset[x].count() > selfset[y].count(); });
    int W = opt.sticky_words;
    if (W != 1 && W != 2) W = cand.size() <= 64 ? 1 : 2;
    const uint32_t nsb = 64u * (uint32_t)W;
    if (cand.size() > nsb) cand.resize(nsb);
    std::vector<int32_t> sticky_bit(N, -1);
    for (size_t b = 0; b < cand.size(); b++) sticky_bit[cand[b]] = (int32_t)b;
    img.n_sticky = (uint32_t)cand.size();

    // ---- class sets (more than two symbols) ------------------------------------------------------
    std::map<SymSet, uint32_t> set_id;
    auto class_of = [&](const SymSet &s) -> uint32_t {
        auto it = set_id.find(s);
        if (it != set_id.end()) return it->second;
        uint32_t id = (uint32_t)set_id.size();
        set_id.emplace(s, id);
        return id;
    };

    // ---- id assignment ---------------------------------------------------------------------------
    // non-sticky edges of a non-sticky state (the self loop is an ordinary edge there)
    auto is_single = [&](uint32_t s) { return edges[s].size() <= 1; };
    img.id_of_orig.assign(N, 0xFFFFFFFFu);
    uint32_t next_id = nsb;
    for (size_t b = 0; b < cand.size(); b++) img.id_of_orig[cand[b]] = (uint32_t)b;
    for (uint32_t s = 0; s < N; s++)
        if (sticky_bit[s] < 0 && is_single(s)) img.id_of_orig[s] = next_id++;
    const uint32_t gbase = next_id;
    int bb = opt.bucket_bits;
    uint32_t n_branch = 0;
    for (uint32_t s = 0; s < N; s++) n_branch += (sticky_bit[s] < 0 && !is_single(s));
    if (bb < 0) bb = 4;
    if (bb > 6) bb = 6;
    const uint32_t NB = 1u << bb;
    for (uint32_t s = 0; s < N; s++)
        if (sticky_bit[s] < 0 && !is_single(s)) { img.id_of_orig[s] = next_id; next_id += NB; }
    const uint32_t chain_base = next_id;

    // ---- bucket hash: pick (mul, shift) minimising the expected table lookups per visit -------------
    // A visit with symbol c costs 1 lookup when bucket(c) holds <= 1 edge, 1 + n when it holds n >= 2
    // (indirection + chain).  Symbols that appear on some edge of the state are what the traffic that
    // activated the state tends to continue with, so they carry the weight; all others share weight 1.
    auto bucket_of = [&](uint32_t c, uint32_t mul, uint32_t sh) { return ((c * mul) >> sh) & (NB - 1); };
    uint32_t best_mul = 1, best_sh = 0;
    double best_cost = 1e300;
    std::vector<uint32_t> branchers;
    for (uint32_t s = 0; s < N; s++) if (sticky_bit[s] < 0 && !is_single(s)) branchers.push_back(s);
    for (uint32_t mul = 1; mul < 64; mul += 2)
        for (uint32_t sh = 0; sh < 8; sh++) {
            double cost = 0;
            for (uint32_t s : branchers) {
                if (cost >= best_cost) break;
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

    // ---- fill tab ----------------------------------------------------------------------------------
    std::vector<uint32_t> tab(chain_base, tab_special(CODE_EMPTY, 0, false));
    bool ids_ok = true;
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
    for (uint32_t s = 0; s < N; s++) {
        if (sticky_bit[s] >= 0) continue;
        const uint32_t id = img.id_of_orig[s];
        if (is_single(s)) {
            if (edges[s].empty()) tab[id] = tab_special(CODE_ACCEPT, 0, false);  // Design/FPGA.v:210-213
            else tab[id] = record_for(edges[s][0], all, false);
            continue;
        }
        for (uint32_t q = 0; q < NB; q++) {
            SymSet vis;
            for (uint32_t c = 0; c < 256; c++) if (bucket_of(c, best_mul, best_sh) == q) vis.set(c);
            std::vector<const Edge *> in;
            for (const Edge &e : edges[s]) if ((e.syms & vis).any()) in.push_back(&e);
            if (in.empty()) continue;  // stays EMPTY
            if (in.size() == 1) { tab[id + q] = record_for(*in[0], vis, false); continue; }
            uint32_t start = (uint32_t)tab.size();
            if (start > 0x7FFF) ids_ok = false;
            tab[id + q] = tab_special(CODE_INDIRECT, start, false);
            for (size_t k = 0; k < in.size(); k++) tab.push_back(record_for(*in[k], vis, k + 1 < in.size()));
        }
    }

    // ---- sticky tables -------------------------------------------------------------------------------
    const uint32_t mstride = 32u * (uint32_t)W;  // bytes per symbol: A[W] (pad to 16) K[W] M[W]
    std::vector<uint8_t> mask(256 * mstride, 0);
    std::vector<uint16_t> inj((size_t)nsb * 256, 0xFFFF);
    std::vector<uint16_t> tlist;
    for (size_t b = 0; b < cand.size(); b++) {
        const uint32_t p = cand[b];
        for (uint32_t c = 0; c < 256; c++) {
            std::vector<uint32_t> tg;
            for (const Edge &e : edges[p]) if (e.tgt != p && e.syms.has(c)) tg.push_back(img.id_of_orig[e.tgt]);
            uint64_t *A = reinterpret_cast<uint64_t *>(&mask[c * mstride]);
            uint64_t *K = reinterpret_cast<uint64_t *>(&mask[c * mstride + 16]);
            uint64_t *M = K + W;
            const uint64_t bit = 1ull << (b & 63);
            if (selfset[p].has(c)) K[b >> 6] |= bit; else A[b >> 6] |= bit;
            if (!tg.empty()) {
                M[b >> 6] |= bit; A[b >> 6] |= bit;
                for (uint32_t t : tg) if (t > 0x7FFF) ids_ok = false;
                if (tg.size() == 1) inj[b * 256 + c] = (uint16_t)tg[0];
                else {
                    if (tlist.size() + tg.size() > 0x7FFE) ids_ok = false;
                    inj[b * 256 + c] = (uint16_t)(0x8000u | tlist.size());
                    for (size_t k = 0; k < tg.size(); k++) tlist.push_back((uint16_t)(tg[k] | (k + 1 < tg.size() ? 0x8000u : 0u)));
                }
            }
        }
    }
    // bits of absent sticky slots: K = 1 (harmless), A = 0
    for (uint32_t c = 0; c < 256; c++) {
        uint64_t *K = reinterpret_cast<uint64_t *>(&mask[c * mstride + 16]);
        for (uint32_t b = (uint32_t)cand.size(); b < nsb; b++) K[b >> 6] |= 1ull << (b & 63);
    }

    // ---- class membership bitmaps ---------------------------------------------------------------------
    std::vector<uint32_t> memb(std::max<size_t>(1, set_id.size()) * 8, 0);
    for (auto &kv : set_id)
        for (uint32_t c = 0; c < 256; c++)
            if (kv.first.has(c)) memb[kv.second * 8 + (c >> 5)] |= 1u << (c & 31);

    // ---- feasibility -------------------------------------------------------------------------------------
    if (tab.size() > 0x8000) ids_ok = false;
    if (set_id.size() > 506) { img.why_not = "more than 506 distinct symbol classes"; }
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
    uint32_t off = 0;
    h.off_tab = off;   off = align16(off + (uint32_t)tab.size() * 4);
    h.off_inj = off;   off = align16(off + (uint32_t)inj.size() * 2);
    h.off_mask = off;  off = align16(off + (uint32_t)mask.size());
    h.off_memb = off;  off = align16(off + (uint32_t)memb.size() * 4);
    h.off_tlist = off; off = align16(off + (uint32_t)std::max<size_t>(8, tlist.size()) * 2);
    h.blob_bytes = off;
    if (img.why_not.empty() && off > opt.max_bytes) img.why_not = "tables need " + std::to_string(off) + " bytes of shared memory (limit " + std::to_string(opt.max_bytes) + ")";
    img.blob.assign(off, 0);
    std::memcpy(&img.blob[h.off_tab], tab.data(), tab.size() * 4);
    std::memcpy(&img.blob[h.off_inj], inj.data(), inj.size() * 2);
    std::memcpy(&img.blob[h.off_mask], mask.data(), mask.size());
    std::memcpy(&img.blob[h.off_memb], memb.data(), memb.size() * 4);
    if (!tlist.empty()) std::memcpy(&img.blob[h.off_tlist], tlist.data(), tlist.size() * 2);

    img.orig_of_id.assign(tab.size(), 0xFFFFFFFFu);
    for (uint32_t s = 0; s < N; s++) if (img.id_of_orig[s] < img.orig_of_id.size()) img.orig_of_id[img.id_of_orig[s]] = s;
    img.ok = img.why_not.empty();
    if (img.ok) {
        int rc = image_verify(nfa, img, err);
        if (rc) { img.ok = false; return rc; }
    }
    return RFB_OK;
}

// bucket_bits < 0: the most buckets (up to 16 per branching state) whose tables still fit.
int image_build(const Nfa &nfa, const ImageOptions &opt, Image &img, std::string &err) {
    if (opt.bucket_bits >= 0) return image_build_one(nfa, opt, img, err);
    ImageOptions o = opt;
    int rc = RFB_OK;
    for (int bb = 4; bb >= 1; bb--) {
        o.bucket_bits = bb;
        rc = image_build_one(nfa, o, img, err);
        if (rc != RFB_OK || img.ok) return rc;
    }
    return rc;
}

// The kernel's semantics, on the host.  Keep in lock-step with scan_lane.cu.
void image_successors(const Image &img, uint32_t s, uint32_t c, std::vector<uint32_t> &out, bool *accepting) {
    const ImageHeader &h = img.h;
    const uint32_t *tab = reinterpret_cast<const uint32_t *>(&img.blob[h.off_tab]);
    const uint16_t *inj = reinterpret_cast<const uint16_t *>(&img.blob[h.off_inj]);
    const uint32_t *memb = reinterpret_cast<const uint32_t *>(&img.blob[h.off_memb]);
    const uint16_t *tlist = reinterpret_cast<const uint16_t *>(&img.blob[h.off_tlist]);
    const uint32_t mstride = 32u * h.sticky_words;
    out.clear();
    if (accepting) *accepting = false;
    const uint32_t id = img.id_of_orig[s];
    std::vector<uint32_t> ids;
    if (id < h.nsb) {
        const uint64_t *K = reinterpret_cast<const uint64_t *>(&img.blob[h.off_mask + c * mstride + 16]);
        const uint64_t *M = K + h.sticky_words;
        const uint64_t bit = 1ull << (id & 63);
        if (K[id >> 6] & bit) ids.push_back(id);
        if (M[id >> 6] & bit) {
            uint32_t x = inj[id * 256 + c];
            if (x != 0xFFFF) {
                if (x < 0x8000) ids.push_back(x);
                else for (uint32_t q = x & 0x7FFF;; q++) { ids.push_back(tlist[q] & 0x7FFF); if (!(tlist[q] & 0x8000)) break; }
            }
        }
    } else {
        uint32_t idx = id;
        if (id >= h.gbase) idx += ((c * h.hash_mul) >> h.hash_shift) & ((1u << h.bucket_bits) - 1);
        for (;;) {
            const uint32_t e = tab[idx];
            const uint32_t a = e & 0xFF, b = (e >> 8) & 0xFF, t = (e >> 16) & 0x7FFF;
            if (a <= b) { if (c == a || c == b) ids.push_back(t); }
            else if (a == 0xFF) {
                if (b == CODE_ACCEPT) { if (accepting) *accepting = true; }
                else if (b == CODE_INDIRECT) { idx = t; continue; }
            } else {
                const uint32_t n = (0xFE - a) * 253 + b;
                if ((memb[n * 8 + (c >> 5)] >> (c & 31)) & 1) ids.push_back(t);
            }
            if (!(e & TAB_MORE)) break;
            idx++;
        }
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
    return RFB_OK;
}

}  // namespace rfb
