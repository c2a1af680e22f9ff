// image_replay.cpp -- TEST INFRASTRUCTURE (built and run by tests/test_image_replay.py, never part of the library):
// interprets the execution image on the host exactly as scan_lane_kernel does -- sticky masks, hashed edge table,
// start DFA with insertion lists -- and prints the match records, so that the dynamic semantics of the tables
// (not only the per-(state, symbol) equivalence that image_verify proves) are checked against the oracle without a GPU.
//   image_replay <coe> <data.bin> <n_streams> <stride> <n_steps> <dfa_budget>
#include "host.h"
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <set>
#include <tuple>
using namespace rfb;

int main(int argc, char **argv) {
    if (argc < 7) return 2;
    std::vector<uint32_t> E; std::string err;
    if (coe_parse_file(argv[1], E, err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
    std::vector<uint8_t> data;
    { FILE *f = fopen(argv[2], "rb"); if (!f) return 1; uint8_t buf[65536]; size_t k; while ((k = fread(buf, 1, sizeof buf, f)) > 0) data.insert(data.end(), buf, buf + k); fclose(f); }
    const uint64_t n_streams = strtoull(argv[3], nullptr, 10), stride = strtoull(argv[4], nullptr, 10);
    const uint32_t n_steps = (uint32_t)atoi(argv[5]);
    ImageOptions opt; opt.dfa_max_states = (uint32_t)atoi(argv[6]); opt.accel = opt.dfa_max_states ? 1 : 0; opt.bucket_bits = -1;
    Nfa nfa; Image img;
    if (nfa_from_entries(E.data(), E.size(), -1, nfa, err) || image_build(nfa, opt, img, err) || !img.ok) { fprintf(stderr, "image: %s %s\n", err.c_str(), img.why_not.c_str()); return 1; }
    const ImageHeader &h = img.h;
    const uint32_t *tab = (const uint32_t *)&img.blob[h.off_tab], *memb = (const uint32_t *)&img.blob[h.off_memb];
    const uint32_t *sdesc = (const uint32_t *)&img.blob[h.off_sdesc], *cmap = (const uint32_t *)&img.blob[h.off_cmap];
    const uint32_t W = h.sticky_words, ms = 32 * W;
    const Image::Dfa &D = img.dfa;
    std::vector<std::tuple<uint32_t, uint32_t, uint32_t>> recs;
    for (uint64_t sid = 0; sid < n_streams; sid++) {
        const uint8_t *sp = &data[sid * stride];
        uint64_t P[2] = {0, 0};
        uint32_t d = 0;
        std::vector<uint32_t> cur, nxt;
        if (h.start_id < h.nsb) P[h.start_id >> 6] |= 1ull << (h.start_id & 63); else cur.push_back(h.start_id);
        for (uint32_t k = 0; k < n_steps; k++) {
            const uint32_t c = sp[k], hf = cmap[c] >> 16, hc = hf & ((1u << h.bucket_bits) - 1);
            std::set<uint32_t> seen;
            auto insert = [&](uint32_t t) { if (t < h.nsb) P[t >> 6] |= 1ull << (t & 63); else if (seen.insert(t).second) nxt.push_back(t); };
            auto walk = [&](uint32_t idx) {
                for (;;) {
                    const uint32_t e = tab[idx], a = e & 0xFF, b = (e >> 8) & 0xFF, t = (e >> 16) & 0x7FFF;
                    if (a <= b) { if (c == a || c == b) insert(t); }
                    else if (a == 0xFF) { idx = t; continue; }
                    else { const uint32_t n = (0xFE - a) * 253 + b; if ((memb[n * 8 + (c >> 5)] >> (c & 31)) & 1) insert(t); }
                    if (!(e & TAB_MORE)) break;
                    idx++;
                }
            };
            // open: start DFA, then the sticky masks (the insertions of this step land in P only after the masks were read)
            std::vector<uint32_t> pending;                      // insertion list of the DFA transition
            if (h.accel) {
                d = std::max<uint32_t>(d, (uint32_t)(P[0] & 1));
                const size_t at = (size_t)d * D.ncls + (cmap[c] & 0xFF);
                d = D.dt[at] & 0x7FFF;
                if (D.dt[at] & 0x8000) for (uint32_t q = D.dta[at];; q++) { pending.push_back(D.act[q] & 0x7FFF); if (!(D.act[q] & 0x8000)) break; }
            }
            const uint64_t *A = (const uint64_t *)&img.blob[h.off_mask + c * ms];
            const uint64_t *K = (const uint64_t *)&img.blob[h.off_mask + c * ms + 16];
            const uint64_t *M = K + W;
            uint64_t fire[2] = {0, 0};
            bool attn = false;
            for (uint32_t w = 0; w < W; w++) attn |= (P[w] & A[w]) != 0;
            if (attn) for (uint32_t w = 0; w < W; w++) { fire[w] = P[w] & M[w]; P[w] &= K[w]; }
            if (attn && k + 1 < n_steps) {   // look-ahead: a firing whose targets cannot outlive the next symbol is skipped
                const uint64_t *LK = (const uint64_t *)&img.blob[h.off_look + sp[k + 1] * 8 * W];
                for (uint32_t w = 0; w < W; w++) fire[w] &= LK[w];
            }
            // drain
            for (uint32_t t : pending) insert(t);
            for (uint32_t u : cur) {
                if (u - h.acc_base < h.n_acc) { recs.emplace_back((uint32_t)sid, k, img.orig_of_id[u]); continue; }
                walk(u + (u >= h.gbase ? hc : 0));
            }
            for (uint32_t w = 0; w < W; w++)
                while (fire[w]) { const uint32_t b = (uint32_t)__builtin_ctzll(fire[w]) + 64 * w; fire[w] &= fire[w] - 1; walk((sdesc[b] & 0xFFFF) + (hf & (sdesc[b] >> 16))); }
            cur.swap(nxt); nxt.clear();
        }
    }
    std::sort(recs.begin(), recs.end());
    for (auto &r : recs) printf("%u %u %u\n", std::get<0>(r), std::get<1>(r), std::get<2>(r));
    return 0;
}
