// imagefile.cpp -- the scan plan of an NFA (its parts and their execution images) and its on-disk form.
//
// the load-time re-indexing as a documented, verifiable, cacheable artefact.  A plan file holds
// the BRAM image as loaded plus, per part, the tables the lane kernel runs on.  Nothing in it is trusted: plan_read()
// checks the checksum, bounds every size against the file, validates the structure of every table (every index
// an interpreter of the tables can follow stays inside them) and then runs image_verify(), which proves the tables
// equivalent to the CSR for every (state, symbol) and every start-DFA transition -- exactly what a freshly built
// image goes through.  The sub-NFAs of a cut NFA are not stored: they are re-derived from the stored state groups.
//
// Layout (little endian; all counts are element counts):
//   char magic[8] = "RFBIMG\0\1";  u32 version = 2;  u32 n_parts;  u64 n_entries;  u32 n_states;  u32 reserved = 0
//   u32 entries[n_entries]                                       the .coe contents (row_ptr | transitions | pad)
//   per part:
//     u32 n_group;  u32 group[n_group]      reference ids of the part's states, ascending, without state 0
//                                           (n_group = 0 and n_parts = 1: the whole NFA)
//     u32 image_ok                          0: this part runs on the general kernel, nothing else is stored for it
//     u32 header_bytes; ImageHeader         (host.h; header_bytes must equal sizeof(ImageHeader))
//     u64 n; u8  blob[n]                    mask | cmap | sdesc | tab | memb | look, staged verbatim into shared memory
//     u64 n; u32 orig_of_id[n]              internal id -> state id of the part
//     u64 n; u32 id_of_orig[n]              state id of the part -> internal id
//     u32 n_sticky; u32 n_sticky_dropped; u32 accel_state; u32 n_absorbed
//     u32 dfa_ncls; u32 dfa_n; u32 dfa_n_frontier
//     u64 n; u16 dt[n];  u64 n; u32 dta[n];  u64 n; u16 act[n];  u64 n; u32 mem_ptr[n];  u64 n; u16 mem_ids[n]
//   u64 fnv1a64 of every byte before it
#include "host.h"
#include "../../include/regex_fpga_b200.h"
#include <cstdio>
#include <cstring>
#include <map>

namespace rfb {
namespace {

const char MAGIC[8] = {'R', 'F', 'B', 'I', 'M', 'G', 0, 1};
const uint32_t FILE_VERSION = 2;       // 2: + look-ahead masks (ImageHeader::off_look)

uint64_t fnv1a64(const uint8_t *p, size_t n) {
    uint64_t h = 0xcbf29ce484222325ull;
    for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 0x100000001b3ull; }
    return h;
}

struct Writer {
    std::vector<uint8_t> b;
    void raw(const void *p, size_t n) { const uint8_t *q = static_cast<const uint8_t *>(p); b.insert(b.end(), q, q + n); }
    void u32(uint32_t v) { raw(&v, 4); }
    void u64(uint64_t v) { raw(&v, 8); }
    template <typename T> void vec(const std::vector<T> &v) { u64(v.size()); if (!v.empty()) raw(v.data(), v.size() * sizeof(T)); }
};

struct Reader {
    const uint8_t *p; size_t n, at = 0; bool ok = true;
    bool raw(void *dst, size_t k) { if (!ok || k > n - at) { ok = false; return false; } std::memcpy(dst, p + at, k); at += k; return true; }
    uint32_t u32() { uint32_t v = 0; raw(&v, 4); return v; }
    uint64_t u64() { uint64_t v = 0; raw(&v, 8); return v; }
    template <typename T> bool vec(std::vector<T> &v, uint64_t max_elems) {
        const uint64_t k = u64();
        if (!ok || k > max_elems || k * sizeof(T) > n - at) { ok = false; return false; }
        v.resize((size_t)k);
        return k == 0 || raw(v.data(), (size_t)k * sizeof(T));
    }
};

}  // namespace

// Every index that image_successors(), image_verify() and the lane kernel can follow stays inside the tables.
int image_validate_structure(const Image &img, uint32_t n_states, std::string &err) {
    const ImageHeader &h = img.h;
    auto bad = [&](const char *what) { err = std::string("execution image is malformed: ") + what; return RFB_E_FORMAT; };
    const size_t B = img.blob.size();
    if (B != h.blob_bytes || B > (1u << 20)) return bad("blob size");
    if (B + 16 * 1024 * 2 + 64 + 256 + 4096 > 227 * 1024) return bad("tables do not fit one SM's shared memory beside the smallest rings");
    if (h.sticky_words != 1 && h.sticky_words != 2) return bad("sticky words");
    const uint32_t W = h.sticky_words;
    if (h.nsb != 64 * W || h.bucket_bits < 1 || h.bucket_bits > 6 || h.hash_shift > 7) return bad("header fields");
    if (h.n_slots == 0 || h.n_slots > 0x8000) return bad("slot count");
    auto fits = [&](uint64_t off, uint64_t bytes) { return off % 4 == 0 && off + bytes <= B; };
    if (h.off_mask != 0 || !fits(h.off_mask, 256ull * 32 * W)) return bad("mask section");
    if (h.off_cmap != 256u * 32 * W || !fits(h.off_cmap, 1024)) return bad("cmap section");
    if (h.off_sdesc != h.off_cmap + 1024 || !fits(h.off_sdesc, 4ull * h.nsb)) return bad("sdesc section");
    if (h.off_tab != h.off_sdesc + 4 * h.nsb || !fits(h.off_tab, 4ull * h.n_slots)) return bad("tab section");
    if (!fits(h.off_memb, 32ull * (h.n_sets ? h.n_sets : 1)) || h.n_sets > 506) return bad("memb section");
    if (h.off_look % 8 != 0 || !fits(h.off_look, 256ull * 8 * W)) return bad("look-ahead section");
    if (h.gbase > h.n_slots || h.acc_base > h.n_slots || h.n_acc > h.n_slots - h.acc_base || h.acc_base < h.nsb) return bad("id ranges");
    if (img.orig_of_id.size() != h.n_slots || img.id_of_orig.size() != n_states) return bad("id maps");
    for (uint32_t s = 0; s < n_states; s++) {
        const uint32_t id = img.id_of_orig[s];
        if (id >= h.n_slots || img.orig_of_id[id] != s) return bad("id maps disagree");
    }
    // orig_of_id must be the EXACT inverse of id_of_orig: a second slot naming the same state would let an edge, an
    // insertion-list entry or a DFA member point at an alias whose row is arbitrary while image_verify(), which maps
    // ids back through orig_of_id, still sees the right successor
    for (uint32_t id = 0; id < h.n_slots; id++) {
        const uint32_t o = img.orig_of_id[id];
        if (o == 0xFFFFFFFFu) continue;
        if (o >= n_states || img.id_of_orig[o] != id) return bad("orig_of_id is not the inverse of id_of_orig");
    }
    if (h.start_id != img.id_of_orig[0]) return bad("start id");
    const uint32_t *tab = reinterpret_cast<const uint32_t *>(&img.blob[h.off_tab]);
    const uint32_t *sdesc = reinterpret_cast<const uint32_t *>(&img.blob[h.off_sdesc]);
    const uint32_t *cmap = reinterpret_cast<const uint32_t *>(&img.blob[h.off_cmap]);
    for (uint32_t i = 0; i < h.n_slots; i++) {
        const uint32_t e = tab[i], a = e & 0xFF, b = (e >> 8) & 0xFF, t = (e >> 16) & 0x7FFF;
        if ((e & TAB_MORE) && i + 1 >= h.n_slots) return bad("chain runs past the table");
        if (a <= b) { if (t >= h.n_slots) return bad("edge target"); }
        else if (a == 0xFF) { if (t >= h.n_slots || t <= i) return bad("chain pointer"); }   // chains lie behind their rows: no cycles
        else if ((0xFE - a) * 253 + b >= (h.n_sets ? h.n_sets : 1)) return bad("class set id");
        else if (t >= h.n_slots) return bad("edge target");
    }
    for (uint32_t b = 0; b < h.nsb; b++)
        if ((uint64_t)(sdesc[b] & 0xFFFFu) + (sdesc[b] >> 16) >= h.n_slots || (sdesc[b] >> 16) > 0xFF) return bad("sticky row");
    // hashed rows: id + (hash & mask) must stay inside the table
    for (uint32_t s = 0; s < n_states; s++)
        if (img.id_of_orig[s] >= h.gbase && img.id_of_orig[s] + (1u << h.bucket_bits) > h.n_slots) return bad("hashed row");
    const Image::Dfa &D = img.dfa;
    if (h.accel > 1) return bad("accel flag");
    if (h.accel) {
        if (img.accel_state >= n_states || img.id_of_orig[img.accel_state] != 0) return bad("always-active state");
        if (D.ncls == 0 || D.ncls > 256 || D.n < 2 || D.n > 32766 || h.dfa_ncls != D.ncls || h.dfa_states != D.n) return bad("start DFA size");
        if (D.dt.size() != (size_t)D.n * D.ncls || D.dta.size() != D.dt.size() || D.mem_ptr.size() != (size_t)D.n + 1) return bad("start DFA tables");
        for (uint32_t c = 0; c < 256; c++) if ((cmap[c] & 0xFF) >= D.ncls) return bad("symbol class");
        if (D.act.empty() || (D.act.back() & 0x8000u)) return bad("insertion lists");
        for (uint16_t v : D.act) if ((uint32_t)(v & 0x7FFF) >= h.n_slots) return bad("insertion list entry");
        for (size_t i = 0; i < D.dt.size(); i++) {
            if ((uint32_t)(D.dt[i] & 0x7FFF) >= D.n) return bad("start DFA target");
            if ((D.dt[i] & 0x8000u) ? (D.dta[i] == 0 || D.dta[i] >= D.act.size()) : D.dta[i] != 0) return bad("insertion list pointer");
        }
        if (D.mem_ptr[0] != 0 || D.mem_ptr.back() != D.mem_ids.size()) return bad("start DFA members");
        for (uint32_t i = 0; i < D.n; i++) if (D.mem_ptr[i] > D.mem_ptr[i + 1]) return bad("start DFA members");
        for (uint16_t v : D.mem_ids) if (v >= h.n_slots) return bad("start DFA member id");
    }
    return RFB_OK;
}

int plan_build(const uint32_t *entries, size_t n_entries, int64_t n_states, const ImageOptions &opt, bool allow_split,
               Plan &plan, std::string &err) {
    plan = Plan();
    int rc = nfa_from_entries(entries, n_entries, n_states, plan.host, err);
    if (rc) return rc;
    // the whole NFA as one part if its tables fit one SM; otherwise groups of connected components, each with
    // tables that fit; if it cannot be split, one part served by the general kernel
    plan.parts.resize(1);
    plan.parts[0].sub = plan.host;
    rc = image_build(plan.host, opt, plan.parts[0].img, err);
    if (rc) return rc;
    if (plan.parts[0].img.ok || !allow_split) return RFB_OK;
    // Prefer the coarsest cut whose parts all get full-quality tables (>= 8 buckets per branching state, every
    // self-looping state in the mask): such a part runs at the speed of a small NFA, and a part that misses
    // either costs far more than one extra pass over the batch.  Otherwise the coarsest cut that fits at all.
    // The cuts are explored with plain images (no sticky states moved into the start DFA: that search is the slow part
    // of a build and does not change whether a part fits); the parts of the chosen cut are then built in full, once
    // per distinct sub-NFA (the replicas of config 5 are identical).
    ImageOptions fast = opt;
    fast.dfa_absorb = 0;
    std::vector<PlanPart> fallback;
    for (uint32_t limit = 24000; limit >= 1500; limit /= 2) {
        std::vector<std::vector<uint32_t>> groups;
        nfa_components(plan.host, limit, groups);
        if (groups.empty()) break;
        std::vector<PlanPart> parts(groups.size());
        std::map<std::vector<uint32_t>, size_t> first_with;     // sub-NFA image -> first part that has it
        bool all_ok = true, all_good = true;
        for (size_t g = 0; g < groups.size(); g++) {
            rc = nfa_extract(plan.host, groups[g], parts[g].sub, parts[g].to_orig, err);
            if (rc) return rc;
            auto it = first_with.find(parts[g].sub.entries);
            if (it != first_with.end()) parts[g].img = parts[it->second].img;
            else {
                rc = image_build(parts[g].sub, fast, parts[g].img, err);
                if (rc) return rc;
                first_with.emplace(parts[g].sub.entries, g);
            }
            all_ok = all_ok && parts[g].img.ok;
            all_good = all_good && parts[g].img.ok && parts[g].img.h.bucket_bits >= 3 && parts[g].img.n_sticky_dropped == 0;
        }
        if (all_good) { plan.parts.swap(parts); fallback.clear(); break; }
        if (all_ok && fallback.empty()) fallback.swap(parts);
    }
    if (!fallback.empty()) plan.parts.swap(fallback);
    if (plan.parts.size() > 1 && opt.dfa_absorb > 0) {
        std::map<std::vector<uint32_t>, size_t> first_with;
        for (size_t g = 0; g < plan.parts.size(); g++) {
            PlanPart &p = plan.parts[g];
            if (!p.img.ok) continue;
            auto it = first_with.find(p.sub.entries);
            if (it != first_with.end()) { p.img = plan.parts[it->second].img; continue; }
            Image full;
            std::string e2;
            if (image_build(p.sub, opt, full, e2) == RFB_OK && full.ok) p.img = std::move(full);
            first_with.emplace(p.sub.entries, g);
        }
    }
    return RFB_OK;
}

int plan_write(const Plan &plan, const std::string &path, std::string &err) {
    Writer w;
    w.raw(MAGIC, 8);
    w.u32(FILE_VERSION); w.u32((uint32_t)plan.parts.size());
    w.u64(plan.host.entries.size()); w.u32(plan.host.n_states); w.u32(0);
    w.raw(plan.host.entries.data(), plan.host.entries.size() * 4);
    for (const PlanPart &p : plan.parts) {
        const uint32_t ng = p.to_orig.empty() ? 0u : (uint32_t)p.to_orig.size() - 1;
        w.u32(ng);
        if (ng) w.raw(p.to_orig.data() + 1, (size_t)ng * 4);
        w.u32(p.img.ok ? 1u : 0u);
        if (!p.img.ok) continue;
        w.u32((uint32_t)sizeof(ImageHeader)); w.raw(&p.img.h, sizeof(ImageHeader));
        w.vec(p.img.blob); w.vec(p.img.orig_of_id); w.vec(p.img.id_of_orig);
        w.u32(p.img.n_sticky); w.u32(p.img.n_sticky_dropped); w.u32(p.img.accel_state); w.u32(p.img.n_absorbed);
        const Image::Dfa &D = p.img.dfa;
        w.u32(D.ncls); w.u32(D.n); w.u32(D.n_frontier);
        w.vec(D.dt); w.vec(D.dta); w.vec(D.act); w.vec(D.mem_ptr); w.vec(D.mem_ids);
    }
    w.u64(fnv1a64(w.b.data(), w.b.size()));
    FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) { err = "cannot create " + path; return RFB_E_IO; }
    const bool ok = std::fwrite(w.b.data(), 1, w.b.size(), f) == w.b.size();
    if (std::fclose(f) != 0 || !ok) { err = "write error on " + path; return RFB_E_IO; }
    return RFB_OK;
}

int plan_read(const std::string &path, Plan &plan, std::string &err) {
    plan = Plan();
    std::vector<uint8_t> buf;
    {
        FILE *f = std::fopen(path.c_str(), "rb");
        if (!f) { err = "cannot open " + path; return RFB_E_IO; }
        std::fseek(f, 0, SEEK_END);
        const long sz = std::ftell(f);
        std::fseek(f, 0, SEEK_SET);
        if (sz < 40 || sz > (1l << 30)) { std::fclose(f); err = path + ": not an execution image file"; return RFB_E_FORMAT; }
        buf.resize((size_t)sz);
        const bool ok = std::fread(buf.data(), 1, buf.size(), f) == buf.size();
        std::fclose(f);
        if (!ok) { err = "read error on " + path; return RFB_E_IO; }
    }
    auto bad = [&](const std::string &what) { err = path + ": " + what; return RFB_E_FORMAT; };
    if (std::memcmp(buf.data(), MAGIC, 8) != 0) return bad("not an execution image file");
    uint64_t sum = 0;
    std::memcpy(&sum, &buf[buf.size() - 8], 8);
    if (sum != fnv1a64(buf.data(), buf.size() - 8)) return bad("checksum mismatch");
    Reader r{buf.data(), buf.size() - 8};
    r.at = 8;
    const uint32_t version = r.u32(), n_parts = r.u32();
    const uint64_t n_entries = r.u64();
    const uint32_t n_states = r.u32();
    r.u32();
    if (!r.ok || version != FILE_VERSION) return bad("unsupported version");
    if (n_parts == 0 || n_parts > 4096 || n_entries > (1ull << 26) || n_entries * 4 > r.n - r.at) return bad("truncated");
    std::vector<uint32_t> entries((size_t)n_entries);
    r.raw(entries.data(), entries.size() * 4);
    int rc = nfa_from_entries(entries.data(), entries.size(), (int64_t)n_states, plan.host, err);
    if (rc) return rc;
    plan.parts.resize(n_parts);
    std::vector<uint8_t> owner(plan.host.n_states, 0);
    for (uint32_t g = 0; g < n_parts; g++) {
        PlanPart &p = plan.parts[g];
        const uint32_t ng = r.u32();
        if (!r.ok || ng >= plan.host.n_states || (ng == 0) != (n_parts == 1)) return bad("bad part table");
        if (ng == 0) p.sub = plan.host;
        else {
            std::vector<uint32_t> group(ng);
            if (!r.raw(group.data(), (size_t)ng * 4)) return bad("truncated");
            for (uint32_t i = 0; i < ng; i++) {
                if (group[i] == 0 || group[i] >= plan.host.n_states || (i && group[i] <= group[i - 1]) || owner[group[i]]) return bad("bad part table");
                owner[group[i]] = 1;
            }
            rc = nfa_extract(plan.host, group, p.sub, p.to_orig, err);   // fails if a transition leaves the group
            if (rc) return rc;
        }
        const uint32_t ok = r.u32();
        if (!r.ok || ok > 1) return bad("truncated");
        if (!ok) { p.img.ok = false; p.img.why_not = "stored without tables"; continue; }
        if (r.u32() != sizeof(ImageHeader) || !r.raw(&p.img.h, sizeof(ImageHeader))) return bad("image header size");
        if (!r.vec(p.img.blob, 1u << 20) || !r.vec(p.img.orig_of_id, 0x8000) || !r.vec(p.img.id_of_orig, 1u << 24)) return bad("truncated");
        p.img.n_sticky = r.u32(); p.img.n_sticky_dropped = r.u32(); p.img.accel_state = r.u32(); p.img.n_absorbed = r.u32();
        Image::Dfa &D = p.img.dfa;
        D.ncls = r.u32(); D.n = r.u32(); D.n_frontier = r.u32();
        if (!r.vec(D.dt, 1u << 24) || !r.vec(D.dta, 1u << 24) || !r.vec(D.act, 1u << 24) || !r.vec(D.mem_ptr, 1u << 16) || !r.vec(D.mem_ids, 1u << 24))
            return bad("truncated");
        rc = image_validate_structure(p.img, p.sub.n_states, err);
        if (rc) { err = path + ": " + err; return rc; }
        rc = image_verify(p.sub, p.img, err);
        if (rc) { err = path + ": " + err; return RFB_E_FORMAT; }
        p.img.ok = true;
    }
    if (!r.ok || r.at != r.n) return bad("trailing or missing bytes");
    if (n_parts > 1)
        for (uint32_t s = 1; s < plan.host.n_states; s++) if (!owner[s]) return bad("a state belongs to no part");
    return RFB_OK;
}

}  // namespace rfb
