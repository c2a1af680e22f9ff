// image_stats.cpp -- developer diagnostic (NOT part of the product library or of any test): replays
// streams through the execution-image tables on the host and counts what the lane kernel would do per
// symbol (list entries, table lookups, indirections, class tests, pushes, sticky "attention" events),
// including the per-32-stream maximum that bounds a lock-step warp.  Used to size kernel design choices.
//   g++ -O2 -std=c++17 -I regex_fpga_b200/csrc tools/image_stats.cpp regex_fpga_b200/csrc/{image,nfa,formats}.cpp -o /tmp/image_stats
//   /tmp/image_stats <coe> <lo.mem> <hi.mem> [n_streams] [bucket_bits] [sticky_words] [dfa_max_states]
#include "host.h"
#include <algorithm>
#include <cstdio>
#include <functional>
#include <cstdlib>
#include <set>
#include <vector>
using namespace rfb;

static uint64_t splitmix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

struct Stats { double g_t2 = 0, g_attn = 0, g_ring2 = 0, g_nonsimple = 0, g_nextbad = 0, gen = 0, gen_simple = 0, surv = 0, inj_useful = 0, t2hits = 0, entries = 0, lookups = 0, indirect = 0, cls = 0, pushes = 0, attn = 0, inj = 0, sticky = 0, symbols = 0, maxlist = 0; };

int main(int argc, char **argv) {
    if (argc < 4) return 1;
    std::vector<uint32_t> E; std::vector<uint8_t> lo, hi; std::string err;
    if (coe_parse_file(argv[1], E, err) || mem_parse_file(argv[2], lo, err) || mem_parse_file(argv[3], hi, err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
    int n_streams = argc > 4 ? atoi(argv[4]) : 1024;
    ImageOptions opt; if (argc > 5) opt.bucket_bits = atoi(argv[5]); if (argc > 6) opt.sticky_words = atoi(argv[6]); if (argc > 7) opt.dfa_max_states = atoi(argv[7]); if (argc > 8) opt.dfa_absorb = atoi(argv[8]);
    Nfa nfa; Image img;
    if (nfa_from_entries(E.data(), E.size(), -1, nfa, err) || image_build(nfa, opt, img, err) || !img.ok) { fprintf(stderr, "image: %s %s\n", err.c_str(), img.why_not.c_str()); return 1; }
    const ImageHeader &h = img.h;
    const uint32_t *tab = (const uint32_t *)&img.blob[h.off_tab];
    const uint32_t *memb = (const uint32_t *)&img.blob[h.off_memb];
    const uint32_t *sdesc = (const uint32_t *)&img.blob[h.off_sdesc];
    const uint32_t W = h.sticky_words, ms = 32 * W, L = 1500;
    printf("start DFA: %u states (%u beyond the budget), %u classes, %zu insertion-list entries, %u sticky states absorbed\n", img.dfa.n, img.dfa.n_frontier, img.dfa.ncls, img.dfa.act.size(), img.n_absorbed);
    printf("image: slots %u gbase %u nsb %u W %u bucket_bits %u bytes %u sets %u sticky %u hash mul %u sh %u\n", h.n_slots, h.gbase, h.nsb, W, h.bucket_bits, h.blob_bytes, h.n_sets, img.n_sticky, h.hash_mul, h.hash_shift);
    Stats st[2];
    std::vector<uint64_t> dhist;
    std::vector<std::vector<uint32_t>> per_sym_lookups(n_streams, std::vector<uint32_t>(L));
    std::vector<std::vector<uint32_t>> per_sym_entries(n_streams, std::vector<uint32_t>(L));
    for (int j = 0; j < n_streams; j++) {
        const std::vector<uint8_t> &src = (j & 1) ? hi : lo;
        uint64_t off = splitmix64(0x5EED0001ull ^ (uint64_t)j) % (std::min(lo.size(), hi.size()) - L + 1);
        Stats &s = st[j & 1];
        uint64_t P[2] = {0, 0};
        uint32_t d = 0;
        const uint32_t *cmap = (const uint32_t *)&img.blob[h.off_cmap]; const Image::Dfa &D = img.dfa;
        std::vector<uint32_t> cur, nxt;
        if (h.start_id < h.nsb) P[h.start_id >> 6] |= 1ull << (h.start_id & 63); else cur.push_back(h.start_id);
        for (uint32_t k = 0; k < L; k++) {
            uint32_t c = src[off + k], hf = ((c * h.hash_mul) >> h.hash_shift) & 0xFF, hc = hf & ((1u << h.bucket_bits) - 1);
            uint64_t Pn[2] = {0, 0};
            std::set<uint32_t> seen;
            uint32_t lk = 0;   // loop iterations of this lane in this step (pops of accept states count too)
            const uint32_t cnext = k + 1 < L ? src[off + k + 1] : 0x100u;
            // would target t have any effect beyond this step? (sticky / accepting / has an edge on the NEXT symbol)
            auto survives = [&](uint32_t t) -> bool {
                if (cnext > 0xFF) return true;
                if (t < h.nsb || t - h.acc_base < h.n_acc) return true;
                const uint32_t hf2 = ((cnext * h.hash_mul) >> h.hash_shift) & 0xFF, hc2 = hf2 & ((1u << h.bucket_bits) - 1);
                uint32_t idx = t + (t >= h.gbase ? hc2 : 0);
                for (;;) {
                    uint32_t e = tab[idx], a = e & 0xFF, b = (e >> 8) & 0xFF, tt = (e >> 16) & 0x7FFF;
                    if (a <= b) { if (cnext == a || cnext == b) return true; }
                    else if (a == 0xFF) { idx = tt; continue; }
                    else { uint32_t n = (0xFE - a) * 253 + b; if ((memb[n * 8 + (cnext >> 5)] >> (cnext & 31)) & 1) return true; }
                    if (!(e & TAB_MORE)) break;
                    idx++;
                }
                return false;
            };
            bool any_surv = false;
            auto push = [&](uint32_t t) { s.pushes++; if (survives(t)) { s.surv++; any_surv = true; } if (t < h.nsb) Pn[t >> 6] |= 1ull << (t & 63); else if (seen.insert(t).second) nxt.push_back(t); };
            auto walk = [&](uint32_t idx) {
                for (;;) {
                    uint32_t e = tab[idx]; lk++;
                    uint32_t a = e & 0xFF, b = (e >> 8) & 0xFF, t = (e >> 16) & 0x7FFF;
                    if (a <= b) { if (c == a || c == b) push(t); }
                    else if (a == 0xFF) { s.indirect++; idx = t; continue; }
                    else { s.cls++; uint32_t n = (0xFE - a) * 253 + b; if ((memb[n * 8 + (c >> 5)] >> (c & 31)) & 1) push(t); }
                    if (!(e & TAB_MORE)) break;
                    idx++;
                }
            };
            bool t2_flag = false, attn_real_flag = false;
            if (h.accel) {
                d = std::max<uint32_t>(d, P[0] & 1);
                const size_t at = (size_t)d * D.ncls + (cmap[c] & 0xFF);
                d = D.dt[at] & 0x7FFF;
                if (dhist.size() < D.n) dhist.resize(D.n, 0);
                dhist[d]++;
                if (D.dt[at] & 0x8000) { t2_flag = true; s.t2hits++; for (uint32_t q = D.dta[at];; q++) { push(D.act[q] & 0x7FFF); lk++; if (!(D.act[q] & 0x8000)) break; } }
            }
            s.entries += cur.size(); per_sym_entries[j][k] = cur.size();
            s.maxlist = std::max<double>(s.maxlist, cur.size());
            const uint64_t *A = (const uint64_t *)&img.blob[h.off_mask + c * ms];
            const uint64_t *K = (const uint64_t *)&img.blob[h.off_mask + c * ms + 16];
            const uint64_t *M = K + W;
            bool attn = false;
            uint64_t im[2] = {0, 0};
            for (uint32_t w = 0; w < W; w++) { attn |= (P[w] & A[w]) != 0; s.sticky += __builtin_popcountll(P[w]); }
            if (attn) {
                bool real = false;
                for (uint32_t w = 0; w < W; w++) { im[w] = P[w] & M[w]; if (P[w] & ~K[w]) real = true; P[w] &= K[w]; }
                if (k + 1 < L && !getenv("NO_LOOKAHEAD")) { const uint64_t *LK = (const uint64_t *)&img.blob[h.off_look + src[off + k + 1] * 8 * W]; for (uint32_t w = 0; w < W; w++) im[w] &= LK[w]; }
                for (uint32_t w = 0; w < W; w++) if (im[w]) real = true;
                if (real) { s.attn++; attn_real_flag = true; }
            }
            for (uint32_t u : cur) {
                if (u - h.acc_base < h.n_acc) { lk++; continue; }
                walk(u + (u >= h.gbase ? hc : 0));
            }
            for (uint32_t w = 0; w < W; w++)
                while (im[w]) { uint32_t b = __builtin_ctzll(im[w]) + 64 * w; im[w] &= im[w] - 1; s.inj++; any_surv = false; walk((sdesc[b] & 0xFFFF) + (hf & (sdesc[b] >> 16))); if (any_surv) s.inj_useful++; }
            s.lookups += lk; per_sym_lookups[j][k] = lk;
            {   // classification of this symbol for the kernel's blocks
                const uint32_t single_base = h.acc_base + h.n_acc;
                auto simple = [&](uint32_t u) { return u >= single_base && u < h.gbase && (tab[u] & 0xFF) <= ((tab[u] >> 8) & 0xFF); };
                const bool flagged = h.accel && false;
                bool dfa_flag = false;
                (void)flagged;
                if (h.accel) { /* recompute: was the DFA transition of this symbol flagged? */ }
                const bool ev_attn = attn_real_flag;
                const bool ring_empty = cur.empty();
                bool one_simple = cur.size() == 1 && simple(cur[0]);
                bool next_ok = nxt.empty() || (nxt.size() == 1 && simple(nxt[0]));
                const bool general = !ring_empty || ev_attn || t2_flag;
                if (general) s.gen++;
                if (general) {
                    if (t2_flag) s.g_t2++;
                    else if (ev_attn) s.g_attn++;
                    else if (cur.size() >= 2) s.g_ring2++;
                    else if (!one_simple) s.g_nonsimple++;
                    else if (!next_ok) s.g_nextbad++;
                }
                // a symbol the extended quiet run could take: no flagged transition, no real attention, at most one simple transient
                // state before and after
                if (general && !ev_attn && !t2_flag && one_simple && next_ok) s.gen_simple++;
                (void)dfa_flag;
            }
            for (uint32_t w = 0; w < W; w++) P[w] |= Pn[w];
            cur.swap(nxt); nxt.clear(); s.symbols++;
        }
    }
    if (getenv("DUMP_STREAMS")) {   // per stream: class, explicit lookups, static firing-weight score (full / first 128 bytes)
        for (int j = 0; j < n_streams; j++) {
            const std::vector<uint8_t> &src = (j & 1) ? hi : lo;
            uint64_t off = splitmix64(0x5EED0001ull ^ (uint64_t)j) % (std::min(lo.size(), hi.size()) - L + 1);
            uint64_t look = 0, sc = 0, sc128 = 0;
            uint32_t fl64 = 0, fl128 = 0, fl256 = 0, st64 = 0, st128 = 0, st256 = 0, dd = 1;   // DFA-only sniff: flagged transitions / sticky insertions in a prefix
            { const uint32_t *cmap2 = (const uint32_t *)&img.blob[h.off_cmap]; const Image::Dfa &D = img.dfa;
              for (uint32_t k = 0; k < 256 && h.accel; k++) {
                const size_t at = (size_t)dd * D.ncls + (cmap2[src[off + k]] & 0xFF);
                dd = D.dt[at] & 0x7FFF; if (dd == 0) dd = 1;
                uint32_t nst = 0;
                if (D.dt[at] & 0x8000) for (uint32_t q = D.dta[at];; q++) { if ((D.act[q] & 0x7FFFu) < h.nsb) nst++; if (!(D.act[q] & 0x8000)) break; }
                const uint32_t f = (D.dt[at] >> 15) & 1;
                if (k < 64) { fl64 += f; st64 += nst; } if (k < 128) { fl128 += f; st128 += nst; } fl256 += f; st256 += nst;
              } }
            for (uint32_t k = 0; k < L; k++) {
                look += per_sym_lookups[j][k];
                const uint32_t c = src[off + k];
                const uint64_t *K = (const uint64_t *)&img.blob[h.off_mask + c * ms + 16];
                const uint64_t *M = K + W;
                uint32_t wgt = 0; for (uint32_t w = 0; w < W; w++) wgt += __builtin_popcountll(M[w]);
                sc += wgt; if (k < 128) sc128 += wgt;
            }
            printf("S %d %d %llu %llu %llu %u %u %u %u %u %u\n", j, j & 1, (unsigned long long)look, (unsigned long long)sc, (unsigned long long)sc128, fl64, fl128, fl256, st64, st128, st256);
        }
    }
    for (int t = 0; t < 2; t++) {
        Stats &s = st[t];
        printf("%s: per symbol: pushes surviving the next symbol %.4f of %.4f; firings with a surviving target %.4f of %.4f\n", t ? "hi" : "lo", s.surv / s.symbols, s.pushes / s.symbols, s.inj_useful / s.symbols, s.inj / s.symbols);
        printf("%s: general-step symbols per stream %.1f, of which a one-simple-transient-state quiet run could take %.1f\n", t ? "hi" : "lo", s.gen / s.symbols * L, s.gen_simple / s.symbols * L);
        printf("%s:   first reason a general step is needed, per stream: flagged DFA transition %.1f, attention %.1f, >= 2 transient states %.1f, one non-simple state %.1f, successor not simple %.1f\n", t ? "hi" : "lo", s.g_t2 / s.symbols * L, s.g_attn / s.symbols * L, s.g_ring2 / s.symbols * L, s.g_nonsimple / s.symbols * L, s.g_nextbad / s.symbols * L);
        printf("%s: per symbol: t2hits %.3f entries %.3f lookups %.3f indirect %.3f class %.3f pushes %.3f attn %.3f inj %.3f sticky %.2f maxlist %.0f\n", t ? "hi" : "lo", s.t2hits / s.symbols,
               s.entries / s.symbols, s.lookups / s.symbols, s.indirect / s.symbols, s.cls / s.symbols, s.pushes / s.symbols, s.attn / s.symbols, s.inj / s.symbols, s.sticky / s.symbols, s.maxlist);
    }
    { std::vector<uint64_t> sh = dhist; std::sort(sh.begin(), sh.end(), std::greater<uint64_t>()); uint64_t tot = 0, cum = 0; for (auto v : sh) tot += v; size_t i = 0;
      for (size_t H : {16, 64, 128, 256, 325, 512, 700, 1024, 2048, 4096}) { for (; i < H && i < sh.size(); i++) cum += sh[i]; printf("DFA lookups landing in the %zu most visited states: %.4f\n", H, tot ? (double)cum / tot : 0.0); } }
    {   // static prior: visits under i.i.d. uniform bytes; how much of the REAL traffic do its top-H states cover?
        const uint32_t *cmap = (const uint32_t *)&img.blob[h.off_cmap]; const Image::Dfa &D = img.dfa;
        std::vector<uint64_t> prior(D.n, 0); uint32_t d = 1; uint64_t x = 12345;
        for (int it = 0; it < 4000000; it++) { x = splitmix64(x); const uint32_t c = x & 0xFF; d = D.dt[(size_t)d * D.ncls + (cmap[c] & 0xFF)] & 0x7FFF; if (!d) d = 1; prior[d]++; }
        std::vector<uint32_t> ord(D.n); for (uint32_t i = 0; i < D.n; i++) ord[i] = i;
        std::stable_sort(ord.begin(), ord.end(), [&](uint32_t a, uint32_t b) { return prior[a] > prior[b]; });
        uint64_t tot = 0, cum = 0; for (auto v : dhist) tot += v; size_t i = 0;
        for (size_t H : {64, 325, 700, 1024, 4096}) { for (; i < H && i < ord.size(); i++) cum += ord[i] < dhist.size() ? dhist[ord[i]] : 0; printf("real lookups covered by the top %zu states of the uniform-byte prior: %.4f\n", H, (double)cum / tot); }
    }
    { uint64_t tot = 0, cum = 0; for (auto v : dhist) tot += v; size_t i = 0; for (size_t H : {16, 64, 128, 256, 325, 512, 1024, 2048, 4096}) { for (; i < H && i < dhist.size(); i++) cum += dhist[i]; printf("DFA lookups landing in states < %zu (breadth-first order): %.4f\n", H, tot ? (double)cum / tot : 0.0); } }
    // lock-step warp bound: mean over (warp, symbol) of max over its 32 lanes
    double sum_max = 0, sum_tot = 0, sum_maxe = 0; uint64_t cnt = 0;
    for (int w0 = 0; w0 + 32 <= n_streams; w0 += 32)
        for (uint32_t k = 0; k < L; k++) {
            uint32_t m = 0, tot = 0, me = 0;
            for (int l = 0; l < 32; l++) { m = std::max(m, per_sym_lookups[w0 + l][k]); tot += per_sym_lookups[w0 + l][k]; me = std::max(me, per_sym_entries[w0 + l][k]); }
            sum_max += m; sum_tot += tot; sum_maxe += me; cnt++;
        }
    printf("per warp-symbol: max-lane lookups %.2f  max-lane entries %.2f  total lookups %.2f (=> %.2f balanced iterations)\n", sum_max / cnt, sum_maxe / cnt, sum_tot / cnt, sum_tot / cnt / 32);
    return 0;
}
