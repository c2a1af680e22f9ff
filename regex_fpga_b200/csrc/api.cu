// api.cu -- the C ABI (include/regex_fpga_b200.h) over the host model and the kernels.
#include "device.h"
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <mutex>
#include <thread>
#include <new>
#include <string>
#include <vector>

using namespace rfb;

static thread_local std::string g_err = "";

static constexpr unsigned MAX_CHUNKS = 256;
struct rfb_ctx {
    int device = 0;
    int n_sms = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;     // rfb_scan: H2D of chunk i+1 overlaps the kernels of chunk i
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t ev_chunk[16] = {};          // chunk i resident
    ScanGlobals *g = nullptr;          // device
    ScanGlobals *g_host = nullptr;     // pinned
    unsigned int *chunk_vals = nullptr; // pinned {1..MAX_CHUNKS}: sources of the chunks_ready updates
    uint2 *rescan = nullptr;
    size_t rescan_cap = 0;
    // grow-only staging for rfb_scan (host-pointer variant)
    uint8_t *d_data = nullptr; size_t d_data_cap = 0;
    unsigned long long *d_offsets = nullptr; size_t d_offsets_cap = 0;
    unsigned int *d_steps = nullptr; size_t d_steps_cap = 0;
    unsigned long long *d_counts = nullptr; size_t d_counts_cap = 0;
    rfb_match *d_records = nullptr; size_t d_records_cap = 0;
    rfb_match *d_sort_tmp = nullptr; size_t d_sort_tmp_cap = 0;       // canonical ordering: ping-pong buffer + histograms
    uint32_t *d_sort_hist = nullptr; size_t d_sort_hist_cap = 0;
    unsigned int *d_state_in = nullptr; size_t d_state_in_cap = 0;     // resumable scans: per-stream sets in / out
    unsigned int *d_state_out = nullptr; size_t d_state_out_cap = 0;
    // pipelined host path (rfb_scan_submit / rfb_scan_wait): two batches in flight, each with its own buffers
    struct Slot {
        uint8_t *d_data = nullptr; size_t d_data_cap = 0;
        ScanGlobals *g = nullptr, *g_host = nullptr;
        unsigned long long *d_counts = nullptr; size_t d_counts_cap = 0;
        rfb_match *d_records = nullptr; size_t d_records_cap = 0;
        uint2 *rescan = nullptr; size_t rescan_cap = 0;
        cudaEvent_t ev_reset = nullptr, ev0 = nullptr, ev1 = nullptr;
        bool busy = false;
        // what rfb_scan_wait needs
        const rfb_nfa *nfa = nullptr; rfb_batch batch{}; uint32_t flags = 0; rfb_result *res = nullptr;
        uint32_t launches = 0; bool want_counts = false;
    } slot[2];
    int slot_head = 0, slots_busy = 0;     // oldest busy slot, number in flight
    cudaStream_t post_stream = nullptr;    // sort + D2H of a finished batch beside the next batch's kernel
    // state of the last enqueued scan (for rfb_scan_collect)
    cudaStream_t last_stream = nullptr;
    unsigned long long last_symbols = 0;
    bool last_ragged = false;
    uint32_t last_launches = 0;
    std::string err;
};

// One independently scannable piece of an NFA: the whole NFA, or one group of its connected components.
struct Part {
    Nfa sub;
    Image img;
    Ecsr ecsr;
    std::vector<uint32_t> to_orig;     // sub state id -> reference state id (empty: identity)
    NfaDev dev{};
    uint32_t *d_entries = nullptr;
    uint32_t *d_eptr = nullptr; unsigned long long *d_erec = nullptr; uint32_t *d_emembs = nullptr;
    uint8_t *d_blob = nullptr;
    uint32_t *d_orig = nullptr, *d_map = nullptr, *d_subof = nullptr;
    uint32_t *d_idof = nullptr, *d_dta = nullptr, *d_mem_ptr = nullptr;
    uint16_t *d_dt = nullptr, *d_act = nullptr, *d_mem_ids = nullptr;
    bool calibrated = false;           // the device copy of the start-DFA tables is ordered by measured visit frequency
    void release() {
        cudaFree(d_entries); cudaFree(d_eptr); cudaFree(d_erec); cudaFree(d_emembs); cudaFree(d_blob);
        cudaFree(d_orig); cudaFree(d_map); cudaFree(d_subof); cudaFree(d_idof); cudaFree(d_dta); cudaFree(d_mem_ptr); cudaFree(d_dt); cudaFree(d_act); cudaFree(d_mem_ids);
    }
};

struct rfb_nfa {
    bool calibrated = false;           // see calibrate_nfa()
    uint64_t calib_symbols = 0;        // symbols of the sample it was measured on
    double calib_hot_fraction = 0.0;   // share of the sample's start-DFA lookups that land in rows held in shared memory
    rfb_ctx *ctx = nullptr;
    int device = 0;                    // copy of ctx->device: the context may be destroyed first
    Nfa host;                          // the NFA as loaded
    std::vector<Part> parts;           // >= 1; several when the tables of the whole NFA do not fit one SM
    NfaDev full{};                     // raw CSR of the whole NFA on the device (cycle model)
    uint32_t *d_full = nullptr;
    // parts of a cut NFA that can share ONE lane-kernel launch (all with tables, same mask width and ring capacity)
    NfaDev *d_parts = nullptr;
    uint32_t multi_words = 0; int multi_cap = 0; size_t multi_smem = 0;
    bool multi_ok = false;
};

static int fail(rfb_ctx *ctx, int code, const std::string &msg) {
    g_err = msg;
    if (ctx) ctx->err = msg;
    return code;
}
static int cuda_fail(rfb_ctx *ctx, cudaError_t e, const char *what) {
    return fail(ctx, RFB_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CU(ctx, call)                                             \
    do {                                                          \
        cudaError_t e_ = (call);                                  \
        if (e_ != cudaSuccess) return cuda_fail(ctx, e_, #call);  \
    } while (0)

template <class T>
static cudaError_t ensure(T *&p, size_t &cap, size_t want) {
    if (want <= cap && p) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t n = std::max<size_t>(want, 16);
    cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&p), n * sizeof(T));
    if (e == cudaSuccess) cap = n;
    return e;
}

static void fill_info(const Nfa &host, const Image &img, rfb_nfa_info *info) {
    std::memset(info, 0, sizeof *info);
    info->n_states = host.n_states;
    info->n_transitions = host.nnz;
    info->n_accepting = host.n_accepting;
    info->n_entries = (uint32_t)host.entries.size();
    info->image_ok = img.ok ? 1u : 0u;
    info->image_bytes = img.h.blob_bytes;
    info->n_sticky = img.n_sticky;
    info->sticky_words = img.h.sticky_words;
    info->n_slots = img.h.n_slots;
    info->n_class_sets = img.h.n_sets;
    info->bucket_bits = img.h.bucket_bits;
    info->n_parts = 1;
}

static ImageOptions default_image_options() {
    ImageOptions opt;
    if (const char *s = std::getenv("RFB_STICKY_WORDS")) opt.sticky_words = std::atoi(s);
    if (const char *s = std::getenv("RFB_BUCKET_BITS")) opt.bucket_bits = std::atoi(s);
    if (const char *s = std::getenv("RFB_STICKY_MIN_SELF")) opt.sticky_min_self = std::atoi(s);
    if (const char *s = std::getenv("RFB_DFA_ABSORB")) opt.dfa_absorb = std::atoi(s);
    if (const char *s = std::getenv("RFB_DFA_STATES")) { const int v = std::atoi(s); if (v <= 0) opt.accel = 0; else opt.dfa_max_states = (uint32_t)v; }
    // the per-stream rings (16 entries x 1024 streams x 2 bytes) share the SM's shared memory with the tables
    opt.max_bytes = (uint32_t)(MAX_DYN_SMEM - 16 * LANE_THREADS * 2 - 64 - 256 - 4096);   // - barrier, the quiet run's class and attention tables
    return opt;
}

extern "C" {

int rfb_abi_version(void) { return RFB_ABI_VERSION; }

const char *rfb_last_error(const rfb_ctx *ctx) { return ctx ? ctx->err.c_str() : g_err.c_str(); }

int rfb_ctx_create(int device_id, rfb_ctx **out) {
    if (!out) return fail(nullptr, RFB_E_INVALID, "out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(nullptr, RFB_E_NODEVICE, std::string("no usable CUDA device (this library has no CPU path): ") +
                                                 (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    if (device_id < 0 || device_id >= n) return fail(nullptr, RFB_E_INVALID, "device_id out of range");
    rfb_ctx *ctx = new (std::nothrow) rfb_ctx();
    if (!ctx) return fail(nullptr, RFB_E_NOMEM, "out of host memory");
    ctx->device = device_id;
    cudaDeviceProp prop;
    if ((e = cudaSetDevice(device_id)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, device_id)) != cudaSuccess) {
        delete ctx;
        return cuda_fail(nullptr, e, "cudaSetDevice");
    }
    if (prop.major < 10) {
        delete ctx;
        return fail(nullptr, RFB_E_NODEVICE, std::string("device ") + prop.name + " is sm_" + std::to_string(prop.major) +
                                                 std::to_string(prop.minor) + "; this library is built for sm_100a only");
    }
    ctx->n_sms = prop.multiProcessorCount;
    for (int i = 0; i < 16; i++)
        if ((e = cudaEventCreateWithFlags(&ctx->ev_chunk[i], cudaEventDisableTiming)) != cudaSuccess) { int rc = cuda_fail(nullptr, e, "event"); rfb_ctx_destroy(ctx); return rc; }
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaEventCreate(&ctx->ev0)) != cudaSuccess || (e = cudaEventCreate(&ctx->ev1)) != cudaSuccess ||
        (e = cudaMalloc(reinterpret_cast<void **>(&ctx->g), sizeof(ScanGlobals))) != cudaSuccess ||
        (e = cudaMallocHost(reinterpret_cast<void **>(&ctx->g_host), sizeof(ScanGlobals))) != cudaSuccess ||
        (e = cudaMallocHost(reinterpret_cast<void **>(&ctx->chunk_vals), MAX_CHUNKS * sizeof(unsigned int))) != cudaSuccess ||
        (e = configure_kernels()) != cudaSuccess) {
        int rc = cuda_fail(nullptr, e, "context setup");
        rfb_ctx_destroy(ctx);
        return rc;
    }
    for (unsigned i = 0; i < MAX_CHUNKS; i++) ctx->chunk_vals[i] = i + 1;
    *out = ctx;
    return RFB_OK;
}

void rfb_ctx_destroy(rfb_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) { cudaStreamSynchronize(ctx->stream); cudaStreamDestroy(ctx->stream); }
    if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
    for (int i = 0; i < 16; i++) if (ctx->ev_chunk[i]) cudaEventDestroy(ctx->ev_chunk[i]);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    cudaFree(ctx->g);
    if (ctx->g_host) cudaFreeHost(ctx->g_host);
    if (ctx->chunk_vals) cudaFreeHost(ctx->chunk_vals);
    cudaFree(ctx->rescan);
    cudaFree(ctx->d_data); cudaFree(ctx->d_offsets); cudaFree(ctx->d_steps);
    cudaFree(ctx->d_counts); cudaFree(ctx->d_records); cudaFree(ctx->d_state_in); cudaFree(ctx->d_state_out);
    cudaFree(ctx->d_sort_tmp); cudaFree(ctx->d_sort_hist);
    if (ctx->post_stream) { cudaStreamSynchronize(ctx->post_stream); cudaStreamDestroy(ctx->post_stream); }
    for (auto &s : ctx->slot) {
        cudaFree(s.d_data); cudaFree(s.g); cudaFree(s.d_counts); cudaFree(s.d_records); cudaFree(s.rescan);
        if (s.g_host) cudaFreeHost(s.g_host);
        if (s.ev_reset) cudaEventDestroy(s.ev_reset);
        if (s.ev0) cudaEventDestroy(s.ev0);
        if (s.ev1) cudaEventDestroy(s.ev1);
    }
    delete ctx;
}

// ---- transition memory ---------------------------------------------------------------------------
// builds the device copy of one part: edge-grouped CSR always, execution image when it fits
static int upload_part(rfb_ctx *ctx, Part &p, uint32_t n_states_full, bool suppress_start_report, std::string &err) {
    cudaError_t e;
    int rc = ecsr_build(p.sub, p.ecsr, err);
    if (rc) return rc;
    // the general kernel keeps two bit vectors and two lists per warp in shared memory: beyond ~400 K states it cannot run
    if (warp_smem_bytes(p.sub.n_states, 1) > MAX_DYN_SMEM) {
        err = "an NFA (part) of " + std::to_string(p.sub.n_states) + " states is beyond what the general kernel holds in shared memory";
        return RFB_E_UNSUPPORTED;
    }
    p.dev.no_report_lane = 0xFFFFFFFFu; p.dev.no_report_sub = 0xFFFFFFFFu;
    if (suppress_start_report) { p.dev.no_report_sub = 0; if (p.img.ok) p.dev.no_report_lane = p.img.id_of_orig[0]; }
    const Nfa &h = p.sub;
    const Ecsr &ec = p.ecsr;
#define UP(ptr, vec, T)                                                                                             \
    if ((e = cudaMalloc(reinterpret_cast<void **>(&ptr), std::max<size_t>(1, (vec).size()) * sizeof(T))) != cudaSuccess || \
        (e = cudaMemcpy(ptr, (vec).data(), (vec).size() * sizeof(T), cudaMemcpyHostToDevice)) != cudaSuccess)         \
        return cuda_fail(ctx, e, "upload NFA tables")
    UP(p.d_entries, h.entries, uint32_t);
    UP(p.d_eptr, ec.eptr, uint32_t);
    UP(p.d_erec, ec.erec, unsigned long long);
    UP(p.d_emembs, ec.memb, uint32_t);
    p.dev.n_states = h.n_states;
    p.dev.n_ref_states = n_states_full;
    p.dev.row_ptr = p.d_entries;
    p.dev.trans = p.d_entries + h.n_states + 1;
    p.dev.eptr = p.d_eptr; p.dev.erec = p.d_erec; p.dev.emembs = p.d_emembs;
    std::vector<uint32_t> sub_of_ref;                                   // reference id -> sub id (parts of a cut NFA only)
    if (!p.to_orig.empty()) {
        UP(p.d_map, p.to_orig, uint32_t); p.dev.state_map = p.d_map;
        sub_of_ref.assign(n_states_full, 0xFFFFFFFFu);
        for (uint32_t i = 0; i < p.to_orig.size(); i++) sub_of_ref[p.to_orig[i]] = i;
        UP(p.d_subof, sub_of_ref, uint32_t); p.dev.sub_of_ref = p.d_subof;
    }
    if (p.img.ok) {
        const Image &im = p.img;
        std::vector<uint32_t> orig = im.orig_of_id;                     // internal id -> REFERENCE state id
        if (!p.to_orig.empty()) for (auto &o : orig) if (o != 0xFFFFFFFFu) o = p.to_orig[o];
        std::vector<uint32_t> idof = im.id_of_orig;                     // REFERENCE state id -> internal id
        if (!p.to_orig.empty()) {
            idof.assign(n_states_full, 0xFFFFFFFFu);
            for (uint32_t i = 0; i < p.to_orig.size(); i++) idof[p.to_orig[i]] = im.id_of_orig[i];
        }
        UP(p.d_blob, im.blob, uint8_t);
        UP(p.d_orig, orig, uint32_t);
        UP(p.d_idof, idof, uint32_t);
        std::vector<uint16_t> mem_ids = im.dfa.mem_ids;
        if (mem_ids.empty()) mem_ids.push_back(0);
        std::vector<uint16_t> dt = im.dfa.dt;                          // + 16 bytes: the lane kernel stages whole 16-byte units of the first rows
        dt.resize(dt.size() + 8, 0);
        UP(p.d_dt, dt, uint16_t);
        UP(p.d_dta, im.dfa.dta, uint32_t);
        std::vector<uint16_t> act = im.dfa.act;
        if (act.empty()) act.push_back(0);
        UP(p.d_act, act, uint16_t);
        UP(p.d_mem_ptr, im.dfa.mem_ptr, uint32_t);
        UP(p.d_mem_ids, mem_ids, uint16_t);
        p.dev.blob = p.d_blob; p.dev.orig_of_id = p.d_orig;
        p.dev.id_of_orig = p.d_idof;
        p.dev.dfa_dt = p.d_dt; p.dev.dfa_dta = p.d_dta; p.dev.dfa_act = p.d_act;
        p.dev.dfa_mem_ptr = p.d_mem_ptr; p.dev.dfa_mem_ids = p.d_mem_ids;
        p.dev.h = im.h;
    }
#undef UP
    return RFB_OK;
}

// Building a plan (bucket hash search, start-DFA construction with its greedy move of sticky states, equivalence proof)
// takes seconds for a large ruleset; a process that loads the same NFA again -- another context, another GPU, a test
// suite -- gets the plan from a small cache keyed by the image and the options.
static int plan_build_cached(const uint32_t *entries, size_t n_entries, int64_t n_states, const ImageOptions &opt, bool split,
                             Plan &plan, std::string &err) {
    struct Entry { std::vector<uint32_t> entries; int64_t n_states; std::vector<uint64_t> key; Plan plan; };
    static std::mutex mu;
    static std::vector<Entry> cache;
    const std::vector<uint64_t> key = {(uint64_t)opt.sticky_words, (uint64_t)opt.sticky_min_self, (uint64_t)(int64_t)opt.bucket_bits, (uint64_t)opt.accel,
                                       opt.dfa_max_states, (uint64_t)opt.dfa_absorb, opt.max_bytes, (uint64_t)split};
    {
        std::lock_guard<std::mutex> lk(mu);
        for (const Entry &e : cache)
            if (e.n_states == n_states && e.key == key && e.entries.size() == n_entries && std::memcmp(e.entries.data(), entries, n_entries * 4) == 0) {
                plan = e.plan;
                return RFB_OK;
            }
    }
    const int rc = plan_build(entries, n_entries, n_states, opt, split, plan, err);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(mu);
    if (cache.size() >= 4) cache.erase(cache.begin());
    cache.push_back(Entry{std::vector<uint32_t>(entries, entries + n_entries), n_states, key, plan});
    return RFB_OK;
}

// Uploads a scan plan (imagefile.cpp) to the context's GPU.
static int nfa_from_plan(rfb_ctx *ctx, Plan &plan, rfb_nfa **out) {
    rfb_nfa *nfa = new (std::nothrow) rfb_nfa();
    if (!nfa) return fail(ctx, RFB_E_NOMEM, "out of host memory");
    nfa->ctx = ctx;
    nfa->device = ctx->device;
    nfa->host = std::move(plan.host);
    nfa->parts.resize(plan.parts.size());
    for (size_t g = 0; g < plan.parts.size(); g++) {
        nfa->parts[g].sub = std::move(plan.parts[g].sub);
        nfa->parts[g].img = std::move(plan.parts[g].img);
        nfa->parts[g].to_orig = std::move(plan.parts[g].to_orig);
    }
    cudaSetDevice(ctx->device);
    std::string err;
    for (size_t g = 0; g < nfa->parts.size(); g++) {
        Part &p = nfa->parts[g];
        // state 0 as seen by this part (parts.cpp keeps only the part's share of its row): see NfaDev::no_report_*
        const bool suppress = nfa->parts.size() > 1 && p.sub.degree(0) == 0 && (nfa->host.degree(0) != 0 || g > 0);
        const int rc = upload_part(ctx, p, nfa->host.n_states, suppress, err);
        if (rc) { rfb_nfa_destroy(nfa); return rc == RFB_E_CUDA ? rc : fail(ctx, rc, err); }
    }
    if (nfa->parts.size() == 1) nfa->full = nfa->parts[0].dev;
    else {
        cudaError_t e;
        if ((e = cudaMalloc(reinterpret_cast<void **>(&nfa->d_full), nfa->host.entries.size() * 4)) != cudaSuccess ||
            (e = cudaMemcpy(nfa->d_full, nfa->host.entries.data(), nfa->host.entries.size() * 4, cudaMemcpyHostToDevice)) != cudaSuccess) {
            rfb_nfa_destroy(nfa);
            return cuda_fail(ctx, e, "upload CSR");
        }
        nfa->full.n_states = nfa->host.n_states;
        nfa->full.n_ref_states = nfa->host.n_states;
        nfa->full.row_ptr = nfa->d_full;
        nfa->full.trans = nfa->d_full + nfa->host.n_states + 1;
    }
    if (nfa->parts.size() > 1 && nfa->parts.size() <= MAX_MULTI_PARTS && !std::getenv("RFB_NO_MULTI")) {
        std::vector<NfaDev> devs;
        bool same = true;
        for (const Part &p : nfa->parts) {
            if (!p.img.ok) { same = false; break; }
            NfaDev d = p.dev;
            d.hot_rows = lane_hot_rows(d.h, &d.hot_bytes);
            if (devs.empty()) { nfa->multi_words = d.h.sticky_words; nfa->multi_cap = lane_ring_cap(d.h); nfa->multi_smem = 0; }
            same = same && d.h.sticky_words == nfa->multi_words && lane_ring_cap(d.h) == nfa->multi_cap;
            nfa->multi_smem = std::max(nfa->multi_smem, lane_smem_bytes(d.h));
            devs.push_back(d);
        }
        if (same && nfa->multi_cap > 0) {
            cudaError_t e;
            if ((e = cudaMalloc(reinterpret_cast<void **>(&nfa->d_parts), devs.size() * sizeof(NfaDev))) != cudaSuccess ||
                (e = cudaMemcpy(nfa->d_parts, devs.data(), devs.size() * sizeof(NfaDev), cudaMemcpyHostToDevice)) != cudaSuccess) {
                rfb_nfa_destroy(nfa);
                return cuda_fail(ctx, e, "upload part table");
            }
            nfa->multi_ok = true;
        }
    }
    *out = nfa;
    return RFB_OK;
}

int rfb_nfa_from_entries(rfb_ctx *ctx, const uint32_t *entries, size_t n_entries, int64_t n_states, rfb_nfa **out) {
    if (!ctx || !out || !entries) return fail(ctx, RFB_E_INVALID, "NULL argument");
    *out = nullptr;
    Plan plan;
    std::string err;
    const int rc = plan_build_cached(entries, n_entries, n_states, default_image_options(), !std::getenv("RFB_NO_SPLIT"), plan, err);
    if (rc) return fail(ctx, rc, err);
    return nfa_from_plan(ctx, plan, out);
}

int rfb_nfa_load_image(rfb_ctx *ctx, const char *path, rfb_nfa **out) {
    if (!ctx || !path || !out) return fail(ctx, RFB_E_INVALID, "NULL argument");
    *out = nullptr;
    Plan plan;
    std::string err;
    const int rc = plan_read(path, plan, err);
    if (rc) return fail(ctx, rc, err);
    return nfa_from_plan(ctx, plan, out);
}

int rfb_nfa_save_image(const rfb_nfa *nfa, const char *path) {
    if (!nfa || !path) return fail(nullptr, RFB_E_INVALID, "NULL argument");
    Plan plan;
    plan.host = nfa->host;
    plan.parts.resize(nfa->parts.size());
    for (size_t g = 0; g < nfa->parts.size(); g++) {
        plan.parts[g].img = nfa->parts[g].img;
        plan.parts[g].to_orig = nfa->parts[g].to_orig;
    }
    std::string err;
    const int rc = plan_write(plan, path, err);
    return rc ? fail(nfa->ctx, rc, err) : RFB_OK;
}

int rfb_image_file_build(const uint32_t *entries, size_t n_entries, int64_t n_states, const char *path) {
    if (!entries || !path) return fail(nullptr, RFB_E_INVALID, "NULL argument");
    Plan plan;
    std::string err;
    int rc = plan_build(entries, n_entries, n_states, default_image_options(), !std::getenv("RFB_NO_SPLIT"), plan, err);
    if (rc == RFB_OK) rc = plan_write(plan, path, err);
    return rc ? fail(nullptr, rc, err) : RFB_OK;
}

int rfb_image_file_check(const char *path, rfb_nfa_info *info) {
    if (!path || !info) return fail(nullptr, RFB_E_INVALID, "NULL argument");
    Plan plan;
    std::string err;
    const int rc = plan_read(path, plan, err);
    if (rc) return fail(nullptr, rc, err);
    fill_info(plan.host, plan.parts[0].img, info);
    bool all_ok = true;
    for (const PlanPart &p : plan.parts) all_ok = all_ok && p.img.ok;
    info->image_ok = all_ok ? 1u : 0u;
    info->n_parts = (uint32_t)plan.parts.size();
    return RFB_OK;
}

int rfb_nfa_load_coe(rfb_ctx *ctx, const char *path, int64_t n_states, rfb_nfa **out) {
    if (!ctx || !path || !out) return fail(ctx, RFB_E_INVALID, "NULL argument");
    std::vector<uint32_t> e;
    std::string err;
    int rc = coe_parse_file(path, e, err);
    if (rc) return fail(ctx, rc, err);
    return rfb_nfa_from_entries(ctx, e.data(), e.size(), n_states, out);
}

void rfb_nfa_destroy(rfb_nfa *nfa) {
    if (!nfa) return;
    cudaSetDevice(nfa->device);
    for (Part &p : nfa->parts) p.release();
    cudaFree(nfa->d_full);
    cudaFree(nfa->d_parts);
    delete nfa;
}

int rfb_nfa_get_info(const rfb_nfa *nfa, rfb_nfa_info *info) {
    if (!nfa || !info) return fail(nullptr, RFB_E_INVALID, "NULL argument");
    fill_info(nfa->host, nfa->parts[0].img, info);
    bool all_ok = true;
    for (const Part &p : nfa->parts) all_ok = all_ok && p.img.ok;
    info->image_ok = all_ok ? 1u : 0u;
    info->n_parts = (uint32_t)nfa->parts.size();
    return RFB_OK;
}

int rfb_nfa_describe(const rfb_nfa *nfa, char *buf, size_t cap) {
    if (!nfa || !buf) return fail(nullptr, RFB_E_INVALID, "NULL argument");
    std::string s;
    for (size_t g = 0; g < nfa->parts.size(); g++) {
        const Part &p = nfa->parts[g];
        char line[512];
        if (p.img.ok)
            std::snprintf(line, sizeof line,
                          "part %zu: states %u kernel lane image_bytes %u slots %u bucket_bits %u sticky %u sticky_dropped %u "
                          "dfa %u dfa_states %u dfa_classes %u dfa_beyond_budget %u dfa_list_entries %zu dfa_absorbed_sticky %u\n",
                          g, p.sub.n_states, p.img.h.blob_bytes, p.img.h.n_slots, p.img.h.bucket_bits, p.img.n_sticky,
                          p.img.n_sticky_dropped, p.img.h.accel, p.img.dfa.n, p.img.dfa.ncls, p.img.dfa.n_frontier,
                          p.img.dfa.act.size(), p.img.n_absorbed);
        else
            std::snprintf(line, sizeof line, "part %zu: states %u kernel general (%s)\n", g, p.sub.n_states, p.img.why_not.c_str());
        s += line;
    }
    if (s.size() + 1 > cap) return fail(nfa->ctx, RFB_E_INVALID, "buffer too small");
    std::memcpy(buf, s.c_str(), s.size() + 1);
    return RFB_OK;
}

int rfb_image_check(const uint32_t *entries, size_t n_entries, int64_t n_states, int sticky_words, int bucket_bits,
                    rfb_nfa_info *info) {
    if (!entries || !info) return fail(nullptr, RFB_E_INVALID, "NULL argument");
    std::string err;
    ImageOptions opt = default_image_options();
    if (sticky_words > 0) opt.sticky_words = sticky_words;
    if (bucket_bits > 0) opt.bucket_bits = bucket_bits;
    else if (!std::getenv("RFB_BUCKET_BITS")) opt.bucket_bits = -1;
    Plan plan;
    const int rc = plan_build_cached(entries, n_entries, n_states, opt, false, plan, err);
    if (rc) return fail(nullptr, rc, err);
    fill_info(plan.host, plan.parts[0].img, info);
    if (!plan.parts[0].img.ok) g_err = plan.parts[0].img.why_not;
    return RFB_OK;
}

int rfb_nfa_get_entries(const rfb_nfa *nfa, uint32_t *entries, size_t capacity) {
    if (!nfa || !entries) return fail(nullptr, RFB_E_INVALID, "NULL argument");
    if (capacity < nfa->host.entries.size()) return fail(nfa->ctx, RFB_E_INVALID, "capacity too small");
    std::memcpy(entries, nfa->host.entries.data(), nfa->host.entries.size() * 4);
    return RFB_OK;
}

// ---- host-side format helpers ---------------------------------------------------------------------
int rfb_trace_load_mem(const char *path, uint8_t **bytes, size_t *n) {
    if (!path || !bytes || !n) return fail(nullptr, RFB_E_INVALID, "NULL argument");
    std::vector<uint8_t> v;
    std::string err;
    int rc = mem_parse_file(path, v, err);
    if (rc) return fail(nullptr, rc, err);
    *bytes = static_cast<uint8_t *>(std::malloc(v.size() ? v.size() : 1));
    if (!*bytes) return fail(nullptr, RFB_E_NOMEM, "out of host memory");
    std::memcpy(*bytes, v.data(), v.size());
    *n = v.size();
    return RFB_OK;
}
int rfb_trace_write_mem(const char *path, const uint8_t *bytes, size_t n) {
    if (!path || (!bytes && n)) return fail(nullptr, RFB_E_INVALID, "NULL argument");
    std::string err;
    int rc = mem_write_file(path, bytes, n, err);
    return rc ? fail(nullptr, rc, err) : RFB_OK;
}
int rfb_coe_parse(const char *path, uint32_t **entries, size_t *n_entries) {
    if (!path || !entries || !n_entries) return fail(nullptr, RFB_E_INVALID, "NULL argument");
    std::vector<uint32_t> v;
    std::string err;
    int rc = coe_parse_file(path, v, err);
    if (rc) return fail(nullptr, rc, err);
    *entries = static_cast<uint32_t *>(std::malloc(v.size() * 4));
    if (!*entries) return fail(nullptr, RFB_E_NOMEM, "out of host memory");
    std::memcpy(*entries, v.data(), v.size() * 4);
    *n_entries = v.size();
    return RFB_OK;
}
int rfb_coe_write(const char *path, const uint32_t *entries, size_t n_entries, int style) {
    if (!path || !entries) return fail(nullptr, RFB_E_INVALID, "NULL argument");
    std::string err;
    int rc = coe_write_file(path, entries, n_entries, style, err);
    return rc ? fail(nullptr, rc, err) : RFB_OK;
}
int64_t rfb_coe_detect_size(const uint32_t *entries, size_t n_entries) {
    if (!entries) return -1;
    return detect_size(entries, n_entries);
}
void rfb_free(void *p) { std::free(p); }
uint32_t rfb_tb_steps(uint32_t trace_entries) { return trace_entries ? trace_entries - 1 : 0; }

static int check_batch(rfb_ctx *ctx, const rfb_batch *b, bool host);

// ---- calibration: start-DFA rows ordered by measured visit frequency ---------------------------------
// The lane kernel keeps the first rows of the start-DFA table in shared memory (scan.cu: lane_hot_rows); a lookup
// there costs a bank-conflict-limited gather, a lookup in global memory one L1 wavefront per lane.  Which rows are
// hot depends on the traffic, not on the NFA alone (breadth-first order or a uniform-byte prior put 68 % of the
// lookups of the shipped traces into the first 400 rows, the measured order 92 %), so the library measures: it walks
// the start DFA over a sample of the caller's streams on the host (state 1 = "A alone" from the second symbol on,
// as in every stream of an unanchored ruleset), renumbers the states by visit count (0 and 1 stay) and replaces the
// DEVICE copy of the tables.  The host image -- and therefore execution-image files -- is unchanged: renumbering
// DFA states changes no result, only where a row lives.
static int calibrate_part(rfb_ctx *ctx, Part &p, const uint8_t *sample, size_t n_streams, size_t pitch, size_t n_steps,
                          uint64_t *symbols, double *hot_fraction) {
    const Image &im = p.img;
    if (!im.ok || !im.h.accel || im.dfa.n < 3) return RFB_OK;
    const Image::Dfa &D = im.dfa;
    const uint32_t *cmap = reinterpret_cast<const uint32_t *>(&im.blob[im.h.off_cmap]);
    uint8_t cls[256];
    for (int c = 0; c < 256; c++) cls[c] = (uint8_t)(cmap[c] & 0xFFu);
    std::vector<uint64_t> visits(D.n, 0);
    for (size_t s = 0; s < n_streams; s++) {
        const uint8_t *sp = sample + s * pitch;
        uint32_t d = 1;
        for (size_t k = 1; k < n_steps; k++) {
            visits[d]++;
            d = D.dt[(size_t)d * D.ncls + cls[sp[k]]] & 0x7FFFu;
            if (d == 0) d = 1;
        }
    }
    std::vector<uint32_t> order(D.n);                  // order[new id] = old id
    for (uint32_t i = 0; i < D.n; i++) order[i] = i;
    std::stable_sort(order.begin() + 2, order.end(), [&](uint32_t a, uint32_t b) { return visits[a] > visits[b]; });
    std::vector<uint32_t> perm(D.n);                   // perm[old id] = new id
    for (uint32_t i = 0; i < D.n; i++) perm[order[i]] = i;
    std::vector<uint16_t> dt((size_t)D.n * D.ncls + 8, 0);
    std::vector<uint32_t> dta((size_t)D.n * D.ncls, 0), mem_ptr(1, 0);
    std::vector<uint16_t> mem_ids;
    for (uint32_t nw = 0; nw < D.n; nw++) {
        const uint32_t o = order[nw];
        for (uint32_t q = 0; q < D.ncls; q++) {
            const uint16_t e = D.dt[(size_t)o * D.ncls + q];
            dt[(size_t)nw * D.ncls + q] = (uint16_t)(perm[e & 0x7FFFu] | (e & 0x8000u));
            dta[(size_t)nw * D.ncls + q] = D.dta[(size_t)o * D.ncls + q];
        }
        for (uint32_t j = D.mem_ptr[o]; j < D.mem_ptr[o + 1]; j++) mem_ids.push_back(D.mem_ids[j]);
        mem_ptr.push_back((uint32_t)mem_ids.size());
    }
    if (mem_ids.empty()) mem_ids.push_back(0);
    // no scan that reads the old order may still be running
    CU(ctx, cudaDeviceSynchronize());
    CU(ctx, cudaMemcpy(p.d_dt, dt.data(), dt.size() * 2, cudaMemcpyHostToDevice));
    CU(ctx, cudaMemcpy(p.d_dta, dta.data(), dta.size() * 4, cudaMemcpyHostToDevice));
    CU(ctx, cudaMemcpy(p.d_mem_ptr, mem_ptr.data(), mem_ptr.size() * 4, cudaMemcpyHostToDevice));
    CU(ctx, cudaMemcpy(p.d_mem_ids, mem_ids.data(), mem_ids.size() * 2, cudaMemcpyHostToDevice));
    p.calibrated = true;
    const uint32_t hot = lane_hot_rows(im.h, nullptr);
    uint64_t tot = 0, in_hot = 0;
    for (uint32_t nw = 0; nw < D.n; nw++) { tot += visits[order[nw]]; if (nw < hot) in_hot += visits[order[nw]]; }
    if (symbols) *symbols = tot;
    if (hot_fraction) *hot_fraction = tot ? (double)in_hot / (double)tot : 0.0;
    return RFB_OK;
}

static constexpr size_t CALIB_STREAMS = 2048;          // sample size (streams); ~3 M symbols for 1500-byte streams
static constexpr size_t CALIB_MIN_STREAMS = 8192;      // batches smaller than this are not worth measuring

// `sample`: n_streams rows of n_steps bytes at `pitch` on the HOST
static int calibrate_nfa(rfb_ctx *ctx, rfb_nfa *nfa, const uint8_t *sample, size_t n_streams, size_t pitch, size_t n_steps) {
    cudaSetDevice(ctx->device);
    for (Part &p : nfa->parts) {
        const int rc = calibrate_part(ctx, p, sample, n_streams, pitch, n_steps, &nfa->calib_symbols, &nfa->calib_hot_fraction);
        if (rc) return rc;
    }
    nfa->calibrated = true;
    return RFB_OK;
}

static bool auto_calibrate_enabled() {
    static const bool on = [] { const char *e = std::getenv("RFB_NO_CALIBRATE"); return !(e && *e && *e != '0'); }();
    return on;
}

// Which streams of a batch are measured: one per block of `step` consecutive streams, at a hashed position inside the
// block.  (A fixed stride aliases with any periodic structure of the batch: every 64th stream of a batch that alternates
// two kinds of traffic is always the same kind -- the first version calibrated W-mix on its quiet half only.)
__host__ __device__ static inline uint64_t calib_stream_index(uint64_t i, uint64_t step) {
    uint64_t z = i + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return i * step + z % step;
}
__global__ void calib_gather_kernel(const uint8_t *__restrict__ data, uint64_t stride, uint64_t step, uint32_t n_steps, uint8_t *__restrict__ out) {
    const uint8_t *src = data + calib_stream_index(blockIdx.x, step) * stride;
    for (uint32_t i = threadIdx.x; i < n_steps; i += blockDim.x) out[(size_t)blockIdx.x * n_steps + i] = src[i];
}

// first large uniformly strided batch an NFA sees: measure on a sample of it (host or device memory)
static int maybe_calibrate(rfb_ctx *ctx, const rfb_nfa *nfa_c, const rfb_batch *b, bool host) {
    rfb_nfa *nfa = const_cast<rfb_nfa *>(nfa_c);       // the device tables are a cache of the NFA, not part of its value
    if (nfa->calibrated || !auto_calibrate_enabled() || b->offsets || b->steps || b->n_streams < CALIB_MIN_STREAMS || b->n_steps < 64) return RFB_OK;
    const size_t n = CALIB_STREAMS, step = (size_t)(b->n_streams / n);
    std::vector<uint8_t> sample(n * (size_t)b->n_steps);
    cudaSetDevice(ctx->device);
    if (host) {
        for (size_t i = 0; i < n; i++)
            std::memcpy(&sample[i * b->n_steps], b->data + calib_stream_index(i, step) * b->stride, b->n_steps);
    } else {
        uint8_t *d_sample = nullptr;
        CU(ctx, cudaMalloc(&d_sample, sample.size()));
        calib_gather_kernel<<<(unsigned)n, 256>>>(b->data, b->stride, step, (uint32_t)b->n_steps, d_sample);
        cudaError_t e = cudaMemcpy(sample.data(), d_sample, sample.size(), cudaMemcpyDeviceToHost);
        cudaFree(d_sample);
        CU(ctx, e);
    }
    return calibrate_nfa(ctx, nfa, sample.data(), n, b->n_steps, b->n_steps);
}

int rfb_nfa_calibrate(rfb_ctx *ctx, rfb_nfa *nfa, const rfb_batch *sample) {
    if (!ctx || !nfa || !sample) return fail(ctx, RFB_E_INVALID, "NULL argument");
    if (nfa->ctx != ctx) return fail(ctx, RFB_E_INVALID, "nfa belongs to another context");
    if (sample->offsets || sample->steps) return fail(ctx, RFB_E_UNSUPPORTED, "rfb_nfa_calibrate takes a uniformly strided host batch");
    const int rc = check_batch(ctx, sample, true);
    if (rc) return rc;
    if (sample->n_streams == 0 || sample->n_steps < 2) return RFB_OK;
    return calibrate_nfa(ctx, nfa, sample->data, (size_t)sample->n_streams, (size_t)sample->stride, sample->n_steps);
}

int rfb_nfa_calibration(const rfb_nfa *nfa, uint64_t *sample_symbols, double *hot_fraction) {
    if (!nfa) return fail(nullptr, RFB_E_INVALID, "NULL argument");
    if (sample_symbols) *sample_symbols = nfa->calibrated ? nfa->calib_symbols : 0;
    if (hot_fraction) *hot_fraction = nfa->calibrated ? nfa->calib_hot_fraction : 0.0;
    return nfa->calibrated ? 1 : 0;
}

// ---- the scan ----------------------------------------------------------------------------------------
static int check_batch(rfb_ctx *ctx, const rfb_batch *b, bool host) {
    if (!b) return fail(ctx, RFB_E_INVALID, "batch is NULL");
    if (b->n_streams > 0xFFF00000ull) return fail(ctx, RFB_E_INVALID, "n_streams exceeds 2^32 - 2^20 (stream ids and the fetch counter are 32-bit)");
    if (b->n_streams && !b->data && b->data_bytes) return fail(ctx, RFB_E_INVALID, "data is NULL");
    if ((b->state_in || b->state_out) && (b->state_cap == 0 || b->state_cap > 255)) return fail(ctx, RFB_E_INVALID, "state_cap must be 1..255 when state_in/state_out are used");
    // every comparison below is written so that it cannot wrap (a huge offset or stride must fail, not pass)
    auto fits = [&](uint64_t off, uint64_t len) { return len <= b->data_bytes && off <= b->data_bytes - len; };
    if (!b->offsets && !b->steps) {
        if (b->n_streams) {
            const uint64_t last = b->n_streams - 1;
            if (last && b->stride > (UINT64_MAX - b->n_steps) / last) return fail(ctx, RFB_E_INVALID, "stride * n_streams overflows");
            if (!fits(last * b->stride, b->n_steps)) return fail(ctx, RFB_E_INVALID, "streams extend past data_bytes");
        }
    } else if (host) {  // host arrays can be read here; device arrays are the caller's responsibility
        for (uint64_t s = 0; s < b->n_streams; s++) {
            const uint64_t len = b->steps ? b->steps[s] : b->n_steps;
            uint64_t off;
            if (b->offsets) off = b->offsets[s];
            else if (s && b->stride > UINT64_MAX / s) return fail(ctx, RFB_E_INVALID, "stride * n_streams overflows");
            else off = s * b->stride;
            if (!fits(off, len)) return fail(ctx, RFB_E_INVALID, "stream " + std::to_string(s) + " extends past data_bytes");
        }
    }
    return RFB_OK;
}

// key bytes that can be non-zero for this batch (byte 0..3 state, 4..7 pos, 8..11 stream)
static uint32_t sort_key_bytes(const rfb_nfa *nfa, const rfb_batch *b, bool host_batch) {
    auto bytes_of = [](uint64_t maxv) { uint32_t m = 0; for (int i = 0; i < 4; i++) if (maxv >> (8 * i)) m |= 1u << i; return m ? m : 1u; };
    uint64_t max_steps = b->n_steps;
    if (b->steps && host_batch) { max_steps = 0; for (uint64_t s = 0; s < b->n_streams; s++) max_steps = std::max<uint64_t>(max_steps, b->steps[s]); }
    else if (b->steps) max_steps = 0xFFFFFFFFull;   // per-stream lengths live on the device in the device-pointer variant
    const uint64_t max_pos = std::min<uint64_t>(0xFFFFFFFFull, (uint64_t)b->pos_base + max_steps);
    const uint64_t max_stream = std::min<uint64_t>(0xFFFFFFFFull, (uint64_t)b->stream_id_base + b->n_streams);
    return bytes_of(nfa->host.n_states) | (bytes_of(max_pos) << 4) | (bytes_of(max_stream) << 8);
}

static int sort_on_device(rfb_ctx *ctx, const rfb_nfa *nfa, const rfb_batch *b, bool host_batch, rfb_match *d_records, uint64_t n, cudaStream_t st) {
    if (n < 2) return RFB_OK;
    CU(ctx, ensure(ctx->d_sort_tmp, ctx->d_sort_tmp_cap, (size_t)n));
    CU(ctx, ensure(ctx->d_sort_hist, ctx->d_sort_hist_cap, sort_hist_words(n)));
    CU(ctx, launch_sort_records(d_records, ctx->d_sort_tmp, n, sort_key_bytes(nfa, b, host_batch), ctx->d_sort_hist, st));
    return RFB_OK;
}

int rfb_scan_collect(rfb_ctx *ctx, rfb_result *res) {
    if (!ctx || !res) return fail(ctx, RFB_E_INVALID, "NULL argument");
    cudaSetDevice(ctx->device);
    CU(ctx, cudaMemcpyAsync(ctx->g_host, ctx->g, sizeof(ScanGlobals), cudaMemcpyDeviceToHost, ctx->last_stream));
    CU(ctx, cudaStreamSynchronize(ctx->last_stream));
    const ScanGlobals &g = *ctx->g_host;
    res->n_matches = g.n_matches;
    res->n_records = res->records ? std::min<uint64_t>(g.n_matches, res->record_capacity) : 0;
    res->n_dropped = g.n_matches - res->n_records;
    res->n_symbols = ctx->last_ragged ? g.n_symbols : ctx->last_symbols;
    res->n_rescanned = g.n_rescan_total;
    res->n_launches = ctx->last_launches;
    float ms = 0.f;
    CU(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    res->gpu_ms = ms;
    return RFB_OK;
}

// Enqueues the scan kernels of one batch (device pointers) on `st`.  The globals must already be reset.
static int enqueue_kernels(rfb_ctx *ctx, const rfb_nfa *nfa, const rfb_batch *b, uint32_t flags, const rfb_result *res,
                           cudaStream_t st, uint32_t *launches, unsigned int chunk_streams = 0,
                           ScanGlobals *g = nullptr, uint2 *rescan = nullptr) {
    if (!g) g = ctx->g;
    if (!rescan) rescan = ctx->rescan;
    BatchDev bd;
    bd.data = b->data; bd.n_streams = b->n_streams; bd.stride = b->stride;
    bd.offsets = reinterpret_cast<const unsigned long long *>(b->offsets);
    bd.steps = b->steps; bd.n_steps = b->n_steps; bd.stream_id_base = b->stream_id_base;
    bd.chunk_streams = chunk_streams; bd.ready = &g->chunks_ready;
    bd.pos_base = b->pos_base; bd.state_cap = b->state_cap; bd.state_in = b->state_in; bd.state_out = b->state_out;
    OutDev od;
    od.counts = (flags & RFB_SCAN_NO_COUNTS) ? nullptr : reinterpret_cast<unsigned long long *>(res->counts);
    od.records = res->records; od.capacity = res->records ? res->record_capacity : 0;
    od.g = g; od.rescan = rescan;
    // a row no kernel writes must read as "overflow", never as stale memory
    if (b->state_out && b->n_streams)
        CU(ctx, cudaMemsetAsync(b->state_out, 0xFF, (size_t)b->n_streams * (1 + (size_t)b->state_cap) * 4, st));
    od.q_rescan_n = nullptr; od.q_next_item = nullptr;
    // A cut NFA whose parts all have tables: ONE launch, CTAs bound to parts (scan.cu: scan_lane_multi_kernel), then one
    // hand-over pass per part.  Not when the final sets are wanted: parts append to the same rows, one after the other.
    if (b->n_streams && nfa->multi_ok && !(flags & RFB_SCAN_FORCE_WARP) && !b->state_out) {
        bd.count_symbols = 1u; bd.state_append = 0u;
        CU(ctx, launch_scan_lane_multi(nfa->d_parts, (uint32_t)nfa->parts.size(), nfa->multi_words, nfa->multi_cap, nfa->multi_smem,
                                       bd, od, ctx->n_sms, st));
        (*launches)++;
        bd.chunk_streams = 0;
        for (size_t pi = 0; pi < nfa->parts.size(); pi++) {
            OutDev op = od;
            op.rescan = rescan + pi * (size_t)b->n_streams;
            op.q_rescan_n = &g->part_rescan[pi];
            op.q_next_item = &g->part_item[pi];
            CU(ctx, launch_scan_warp(nfa->parts[pi].dev, bd, op, true, ctx->n_sms, st)); (*launches)++;
        }
        return RFB_OK;
    }
    if (b->n_streams) {
        bool first = true;
        for (const Part &p : nfa->parts) {   // one pass over the batch per part; reports of different parts are disjoint
            if (!first) {
                CU(ctx, cudaMemsetAsync(&g->next_stream, 0, 3 * sizeof(unsigned int), st));
                bd.chunk_streams = 0;        // the batch is resident once the first pass has consumed it
            }
            bd.count_symbols = first ? 1u : 0u;
            bd.state_append = first ? 0u : 1u;
            const bool lane = p.img.ok && !(flags & RFB_SCAN_FORCE_WARP);
            if (lane) {
                CU(ctx, launch_scan_lane(p.dev, bd, od, ctx->n_sms, st)); (*launches)++;
                CU(ctx, launch_scan_warp(p.dev, bd, od, true, ctx->n_sms, st)); (*launches)++;
            } else {
                CU(ctx, launch_scan_warp(p.dev, bd, od, false, ctx->n_sms, st)); (*launches)++;
            }
            first = false;
        }
    }
    return RFB_OK;
}

int rfb_scan_device(rfb_ctx *ctx, const rfb_nfa *nfa, const rfb_batch *b, uint32_t flags, void *cuda_stream, rfb_result *res) {
    if (!ctx || !nfa || !res) return fail(ctx, RFB_E_INVALID, "NULL argument");
    if (nfa->ctx != ctx) return fail(ctx, RFB_E_INVALID, "nfa belongs to another context");
    int rc = check_batch(ctx, b, false);
    if (rc) return rc;
    if ((flags & RFB_SCAN_SORT_RECORDS) && (flags & RFB_SCAN_ASYNC)) return fail(ctx, RFB_E_UNSUPPORTED, "RFB_SCAN_SORT_RECORDS needs the record count: not available with RFB_SCAN_ASYNC");
    if (!(flags & RFB_SCAN_FORCE_WARP) && (rc = maybe_calibrate(ctx, nfa, b, false)) != RFB_OK) return rc;
    cudaSetDevice(ctx->device);
    (void)cudaGetLastError();   // a stale error of an unrelated earlier call must not be blamed on this launch
    cudaStream_t st = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->stream;
    { const size_t want = (size_t)b->n_streams * (nfa->multi_ok ? nfa->parts.size() : 1); if (ctx->rescan_cap < want) CU(ctx, ensure(ctx->rescan, ctx->rescan_cap, want)); }

    if (flags & RFB_SCAN_ACCUMULATE) {
        CU(ctx, cudaMemsetAsync(&ctx->g->next_stream, 0, 3 * sizeof(unsigned int), st));
        CU(ctx, cudaMemsetAsync(ctx->g->part_next, 0, 3 * MAX_MULTI_PARTS * sizeof(unsigned int), st));
    } else {
        CU(ctx, cudaMemsetAsync(ctx->g, 0, sizeof(ScanGlobals), st));
        if (res->counts && !(flags & RFB_SCAN_NO_COUNTS))
            CU(ctx, cudaMemsetAsync(res->counts, 0, (size_t)nfa->host.n_states * sizeof(uint64_t), st));
    }
    uint32_t launches = 0;
    CU(ctx, cudaEventRecord(ctx->ev0, st));
    rc = enqueue_kernels(ctx, nfa, b, flags, res, st, &launches);
    if (rc) return rc;
    CU(ctx, cudaEventRecord(ctx->ev1, st));
    ctx->last_stream = st;
    ctx->last_ragged = b->steps != nullptr;
    ctx->last_symbols = b->steps ? 0 : b->n_streams * (unsigned long long)b->n_steps;
    ctx->last_launches = launches;
    if (flags & RFB_SCAN_ASYNC) return RFB_OK;
    rc = rfb_scan_collect(ctx, res);
    if (rc == RFB_OK && (flags & RFB_SCAN_SORT_RECORDS) && res->n_records > 1) {
        rc = sort_on_device(ctx, nfa, b, false, res->records, res->n_records, st);
        if (rc == RFB_OK) CU(ctx, cudaStreamSynchronize(st));
    }
    return rc;
}

// Host path: uniformly strided batches on the lane kernel are copied in up to 16 chunks of whole streams; after each
// chunk the copy stream bumps *ready (a counter in the scan's globals) and the kernel's warps wait for the chunk of the
// stream they take.  Returns the chunk size in streams (0: one copy, no gating).
static uint64_t plan_chunks(const rfb_nfa *nfa, const rfb_batch *b, uint32_t flags, uint64_t *n_chunks_out) {
    const bool lane_path = nfa->parts[0].img.ok && !(flags & RFB_SCAN_FORCE_WARP);
    uint64_t n_chunks = 1;
    if (lane_path && !b->offsets && b->n_streams >= 64 && b->stride > 0) {
        uint64_t max_chunks = 16, min_bytes = 64ull << 20;
        if (const char *e = getenv("RFB_CHUNKS")) max_chunks = std::min<uint64_t>(MAX_CHUNKS, std::max(1, atoi(e)));
        if (const char *e = getenv("RFB_CHUNK_MB")) min_bytes = (uint64_t)std::max(1, atoi(e)) << 20;
        n_chunks = std::min<uint64_t>(max_chunks, std::max<uint64_t>(1, b->data_bytes / min_bytes));
    }
    *n_chunks_out = n_chunks;
    // whole multiples of 128 streams: every chunk boundary is 128-byte aligned whatever the stride, so no cache line of
    // the kernel's (coherent) input loads straddles a chunk that has not landed yet
    return n_chunks > 1 ? ((b->n_streams + n_chunks - 1) / n_chunks + 127) / 128 * 128 : 0;
}
static int copy_chunks(rfb_ctx *ctx, const rfb_batch *b, uint8_t *d_data, unsigned int *ready, uint64_t n_chunks, uint64_t chunk_streams, cudaStream_t cs) {
    for (uint64_t c = 0; c < n_chunks; c++) {
        const uint64_t s0 = std::min<uint64_t>(c * chunk_streams, b->n_streams);
        const uint64_t s1 = std::min<uint64_t>((c + 1) * chunk_streams, b->n_streams);
        const size_t lo = (size_t)(s0 * b->stride);
        const size_t hi = (c + 1 == n_chunks || s1 == b->n_streams) ? (size_t)b->data_bytes : (size_t)(s1 * b->stride);
        if (hi > lo) CU(ctx, cudaMemcpyAsync(d_data + lo, b->data + lo, hi - lo, cudaMemcpyHostToDevice, cs));
        CU(ctx, cudaMemcpyAsync(ready, &ctx->chunk_vals[c], sizeof(unsigned int), cudaMemcpyHostToDevice, cs));
    }
    return RFB_OK;
}

// internal (rfb_group_scan): leave counts and records in the context's device buffers instead of copying them to the host
static constexpr uint32_t SCAN_KEEP_ON_DEVICE = 0x80000000u;

int rfb_scan(rfb_ctx *ctx, const rfb_nfa *nfa, const rfb_batch *b, uint32_t flags, rfb_result *res) {
    if (!ctx || !nfa || !res) return fail(ctx, RFB_E_INVALID, "NULL argument");
    if (nfa->ctx != ctx) return fail(ctx, RFB_E_INVALID, "nfa belongs to another context");
    int rc = check_batch(ctx, b, true);
    if (rc) return rc;
    if (b->state_in)   // host rows can be checked: an overflow mark or a foreign id cannot be resumed
        for (uint64_t s = 0; s < b->n_streams; s++) {
            const unsigned int *row = b->state_in + s * (1 + (uint64_t)b->state_cap);
            if (row[0] > b->state_cap)
                return fail(ctx, RFB_E_INVALID, "state_in of stream " + std::to_string(s) + (row[0] == 0xFFFFFFFFu ? " is an overflow mark (RFB_STATE_OVERFLOW): that stream cannot be resumed" : " holds more than state_cap ids"));
            for (unsigned int q = 0; q < row[0]; q++)
                if (row[1 + q] >= nfa->host.n_states) return fail(ctx, RFB_E_INVALID, "state_in of stream " + std::to_string(s) + " holds an id that is not a state of this NFA");
        }
    if (!(flags & RFB_SCAN_FORCE_WARP) && (rc = maybe_calibrate(ctx, nfa, b, true)) != RFB_OK) return rc;
    cudaSetDevice(ctx->device);
    (void)cudaGetLastError();
    cudaStream_t st = ctx->stream, cs = ctx->copy_stream;
    const size_t padded = ((size_t)b->data_bytes + 15) / 16 * 16 + 16;
    CU(ctx, ensure(ctx->d_data, ctx->d_data_cap, padded));
    { const size_t want = (size_t)b->n_streams * (nfa->multi_ok ? nfa->parts.size() : 1); if (ctx->rescan_cap < want) CU(ctx, ensure(ctx->rescan, ctx->rescan_cap, want)); }
    rfb_batch db = *b;
    db.data = ctx->d_data;
    if (b->offsets) {
        CU(ctx, ensure(ctx->d_offsets, ctx->d_offsets_cap, (size_t)b->n_streams));
        CU(ctx, cudaMemcpyAsync(ctx->d_offsets, b->offsets, b->n_streams * 8, cudaMemcpyHostToDevice, st));
        db.offsets = reinterpret_cast<const uint64_t *>(ctx->d_offsets);
    }
    if (b->steps) {
        CU(ctx, ensure(ctx->d_steps, ctx->d_steps_cap, (size_t)b->n_streams));
        CU(ctx, cudaMemcpyAsync(ctx->d_steps, b->steps, b->n_streams * 4, cudaMemcpyHostToDevice, st));
        db.steps = ctx->d_steps;
    }
    const size_t state_words = (size_t)b->n_streams * (1 + (size_t)b->state_cap);
    if (b->state_in) {
        CU(ctx, ensure(ctx->d_state_in, ctx->d_state_in_cap, state_words));
        CU(ctx, cudaMemcpyAsync(ctx->d_state_in, b->state_in, state_words * 4, cudaMemcpyHostToDevice, st));
        db.state_in = ctx->d_state_in;
    }
    if (b->state_out) {
        CU(ctx, ensure(ctx->d_state_out, ctx->d_state_out_cap, state_words));
        db.state_out = ctx->d_state_out;
    }
    rfb_result dr = *res;
    const bool want_counts = res->counts && !(flags & RFB_SCAN_NO_COUNTS);
    if (want_counts) {
        CU(ctx, ensure(ctx->d_counts, ctx->d_counts_cap, (size_t)nfa->host.n_states));
        dr.counts = reinterpret_cast<uint64_t *>(ctx->d_counts);
        CU(ctx, cudaMemsetAsync(ctx->d_counts, 0, (size_t)nfa->host.n_states * 8, st));
    } else dr.counts = nullptr;
    if (res->records && res->record_capacity) {
        CU(ctx, ensure(ctx->d_records, ctx->d_records_cap, (size_t)res->record_capacity));
        dr.records = ctx->d_records;
    } else { dr.records = nullptr; dr.record_capacity = 0; }
    const uint32_t kflags = flags & (RFB_SCAN_FORCE_WARP | RFB_SCAN_NO_COUNTS);
    CU(ctx, cudaMemsetAsync(ctx->g, 0, sizeof(ScanGlobals), st));

    // ONE lane-kernel launch over the whole batch runs beside the chunked copy (plan_chunks / copy_chunks above): its
    // warps take streams in ascending order and wait for their chunk's counter, so the kernel overlaps the H2D transfer
    // at full occupancy.  Other batches (explicit offsets, general kernel only) wait for the whole copy.
    uint32_t launches = 0;
    uint64_t n_chunks = 1;
    const uint64_t chunk_streams = plan_chunks(nfa, b, flags, &n_chunks);
    CU(ctx, cudaEventRecord(ctx->ev_chunk[0], st));                  // globals reset before the first flag write
    CU(ctx, cudaStreamWaitEvent(cs, ctx->ev_chunk[0], 0));
    CU(ctx, cudaEventRecord(ctx->ev0, st));
    if (n_chunks > 1) {
        // copies first: with pageable host memory they complete before the launch below (no overlap, no hazard)
        rc = copy_chunks(ctx, b, ctx->d_data, &ctx->g->chunks_ready, n_chunks, chunk_streams, cs);
        if (rc) return rc;
        rc = enqueue_kernels(ctx, nfa, &db, kflags, &dr, st, &launches, (unsigned int)chunk_streams);
        if (rc) return rc;
    } else {
        if (b->data_bytes) CU(ctx, cudaMemcpyAsync(ctx->d_data, b->data, b->data_bytes, cudaMemcpyHostToDevice, cs));
        CU(ctx, cudaEventRecord(ctx->ev_chunk[1], cs));
        CU(ctx, cudaStreamWaitEvent(st, ctx->ev_chunk[1], 0));
        rc = enqueue_kernels(ctx, nfa, &db, kflags, &dr, st, &launches);
        if (rc) return rc;
    }
    CU(ctx, cudaEventRecord(ctx->ev1, st));
    ctx->last_stream = st;
    ctx->last_ragged = b->steps != nullptr;
    ctx->last_symbols = b->steps ? 0 : b->n_streams * (unsigned long long)b->n_steps;
    ctx->last_launches = launches;
    rc = rfb_scan_collect(ctx, &dr);
    if (rc) return rc;
    if ((flags & RFB_SCAN_SORT_RECORDS) && dr.n_records > 1) {
        rc = sort_on_device(ctx, nfa, b, true, ctx->d_records, dr.n_records, st);
        if (rc) return rc;
    }
    if (!(flags & SCAN_KEEP_ON_DEVICE)) {
        if (want_counts) CU(ctx, cudaMemcpyAsync(res->counts, ctx->d_counts, (size_t)nfa->host.n_states * 8, cudaMemcpyDeviceToHost, st));
        if (dr.n_records) CU(ctx, cudaMemcpyAsync(res->records, ctx->d_records, dr.n_records * sizeof(rfb_match), cudaMemcpyDeviceToHost, st));
    }
    if (b->state_out && state_words) CU(ctx, cudaMemcpyAsync(b->state_out, ctx->d_state_out, state_words * 4, cudaMemcpyDeviceToHost, st));
    CU(ctx, cudaStreamSynchronize(st));
    res->n_matches = dr.n_matches; res->n_records = dr.n_records; res->n_dropped = dr.n_dropped;
    res->n_symbols = dr.n_symbols; res->n_rescanned = dr.n_rescanned; res->gpu_ms = dr.gpu_ms;
    res->n_launches = dr.n_launches;
    return RFB_OK;
}

// ---- pipelined host path ------------------------------------------------------------------------------
// rfb_scan spends its last milliseconds on the tail of the kernel, the record sort and the D2H of the results while
// the PCIe link idles.  With two batches in flight the next batch's H2D copy runs during all of that: the copy stream
// resets the slot's globals and streams the chunks, the compute stream runs the kernels (stream order keeps batches
// apart), and rfb_scan_wait finishes the oldest batch on a third stream.
static int slot_setup(rfb_ctx *ctx, rfb_ctx::Slot &s) {
    if (s.g) return RFB_OK;
    CU(ctx, cudaMalloc(reinterpret_cast<void **>(&s.g), sizeof(ScanGlobals)));
    CU(ctx, cudaMallocHost(reinterpret_cast<void **>(&s.g_host), sizeof(ScanGlobals)));
    CU(ctx, cudaEventCreateWithFlags(&s.ev_reset, cudaEventDisableTiming));
    CU(ctx, cudaEventCreate(&s.ev0));
    CU(ctx, cudaEventCreate(&s.ev1));
    if (!ctx->post_stream) CU(ctx, cudaStreamCreateWithFlags(&ctx->post_stream, cudaStreamNonBlocking));
    return RFB_OK;
}

int rfb_scan_submit(rfb_ctx *ctx, const rfb_nfa *nfa, const rfb_batch *b, uint32_t flags, rfb_result *res) {
    if (!ctx || !nfa || !res) return fail(ctx, RFB_E_INVALID, "NULL argument");
    if (nfa->ctx != ctx) return fail(ctx, RFB_E_INVALID, "nfa belongs to another context");
    int rc = check_batch(ctx, b, true);
    if (rc) return rc;
    if (b->offsets || b->steps || b->state_in || b->state_out)
        return fail(ctx, RFB_E_UNSUPPORTED, "rfb_scan_submit takes uniformly strided batches without per-stream lengths or resumed state");
    if (ctx->slots_busy == 2) return fail(ctx, RFB_E_INVALID, "two batches are in flight: call rfb_scan_wait first");
    if (ctx->slots_busy == 0 && !(flags & RFB_SCAN_FORCE_WARP) && (rc = maybe_calibrate(ctx, nfa, b, true)) != RFB_OK) return rc;
    cudaSetDevice(ctx->device);
    (void)cudaGetLastError();
    rfb_ctx::Slot &s = ctx->slot[(ctx->slot_head + ctx->slots_busy) & 1];
    rc = slot_setup(ctx, s);
    if (rc) return rc;
    cudaStream_t st = ctx->stream, cs = ctx->copy_stream;
    const size_t padded = ((size_t)b->data_bytes + 15) / 16 * 16 + 16;
    CU(ctx, ensure(s.d_data, s.d_data_cap, padded));
    { const size_t want = (size_t)b->n_streams * (nfa->multi_ok ? nfa->parts.size() : 1); if (s.rescan_cap < want) CU(ctx, ensure(s.rescan, s.rescan_cap, want)); }
    rfb_batch db = *b;
    db.data = s.d_data;
    rfb_result dr = *res;
    s.want_counts = res->counts && !(flags & RFB_SCAN_NO_COUNTS);
    if (s.want_counts) {
        CU(ctx, ensure(s.d_counts, s.d_counts_cap, (size_t)nfa->host.n_states));
        dr.counts = reinterpret_cast<uint64_t *>(s.d_counts);
        CU(ctx, cudaMemsetAsync(s.d_counts, 0, (size_t)nfa->host.n_states * 8, cs));
    } else dr.counts = nullptr;
    if (res->records && res->record_capacity) {
        CU(ctx, ensure(s.d_records, s.d_records_cap, (size_t)res->record_capacity));
        dr.records = s.d_records;
    } else { dr.records = nullptr; dr.record_capacity = 0; }
    const uint32_t kflags = flags & (RFB_SCAN_FORCE_WARP | RFB_SCAN_NO_COUNTS);
    // the copy stream owns the slot's globals until the kernel starts: it must not queue behind the previous
    // batch's kernel, or the copy could not overlap it
    CU(ctx, cudaMemsetAsync(s.g, 0, sizeof(ScanGlobals), cs));
    CU(ctx, cudaEventRecord(s.ev_reset, cs));
    uint64_t n_chunks = 1;
    const uint64_t chunk_streams = plan_chunks(nfa, b, flags, &n_chunks);
    uint32_t launches = 0;
    if (n_chunks > 1) {
        rc = copy_chunks(ctx, b, s.d_data, &s.g->chunks_ready, n_chunks, chunk_streams, cs);
        if (rc) return rc;
        CU(ctx, cudaStreamWaitEvent(st, s.ev_reset, 0));
        CU(ctx, cudaEventRecord(s.ev0, st));
        rc = enqueue_kernels(ctx, nfa, &db, kflags, &dr, st, &launches, (unsigned int)chunk_streams, s.g, s.rescan);
        if (rc) return rc;
    } else {
        if (b->data_bytes) CU(ctx, cudaMemcpyAsync(s.d_data, b->data, b->data_bytes, cudaMemcpyHostToDevice, cs));
        CU(ctx, cudaEventRecord(s.ev_reset, cs));
        CU(ctx, cudaStreamWaitEvent(st, s.ev_reset, 0));
        CU(ctx, cudaEventRecord(s.ev0, st));
        rc = enqueue_kernels(ctx, nfa, &db, kflags, &dr, st, &launches, 0, s.g, s.rescan);
        if (rc) return rc;
    }
    CU(ctx, cudaEventRecord(s.ev1, st));
    s.nfa = nfa; s.batch = *b; s.flags = flags; s.res = res; s.launches = launches; s.busy = true;
    ctx->slots_busy++;
    return RFB_OK;
}

int rfb_scan_wait(rfb_ctx *ctx, rfb_result **done) {
    if (!ctx) return fail(ctx, RFB_E_INVALID, "NULL argument");
    if (done) *done = nullptr;
    if (ctx->slots_busy == 0) return fail(ctx, RFB_E_INVALID, "no batch is in flight");
    cudaSetDevice(ctx->device);
    rfb_ctx::Slot &s = ctx->slot[ctx->slot_head];
    cudaStream_t ps = ctx->post_stream;
    rfb_result *res = s.res;
    // whatever happens below, the slot is released
    s.busy = false; ctx->slot_head ^= 1; ctx->slots_busy--;
    CU(ctx, cudaStreamWaitEvent(ps, s.ev1, 0));
    CU(ctx, cudaMemcpyAsync(s.g_host, s.g, sizeof(ScanGlobals), cudaMemcpyDeviceToHost, ps));
    CU(ctx, cudaStreamSynchronize(ps));
    const ScanGlobals &g = *s.g_host;
    const uint64_t n_records = res->records ? std::min<uint64_t>(g.n_matches, res->record_capacity) : 0;
    if ((s.flags & RFB_SCAN_SORT_RECORDS) && n_records > 1) {
        const int rc = sort_on_device(ctx, s.nfa, &s.batch, true, s.d_records, n_records, ps);
        if (rc) return rc;
    }
    if (s.want_counts) CU(ctx, cudaMemcpyAsync(res->counts, s.d_counts, (size_t)s.nfa->host.n_states * 8, cudaMemcpyDeviceToHost, ps));
    if (n_records) CU(ctx, cudaMemcpyAsync(res->records, s.d_records, n_records * sizeof(rfb_match), cudaMemcpyDeviceToHost, ps));
    CU(ctx, cudaStreamSynchronize(ps));
    float ms = 0.f;
    CU(ctx, cudaEventElapsedTime(&ms, s.ev0, s.ev1));
    res->n_matches = g.n_matches; res->n_records = n_records; res->n_dropped = g.n_matches - n_records;
    res->n_symbols = s.batch.n_streams * (unsigned long long)s.batch.n_steps;
    res->n_rescanned = g.n_rescan_total; res->gpu_ms = ms; res->n_launches = s.launches;
    if (done) *done = res;
    return RFB_OK;
}

int rfb_fpga_cycles(rfb_ctx *ctx, const rfb_nfa *nfa, const uint8_t *lo, const uint8_t *hi, uint32_t trace_entries, uint64_t *cycles) {
    if (!ctx || !nfa || !lo || !hi || !cycles) return fail(ctx, RFB_E_INVALID, "NULL argument");
    if (nfa->ctx != ctx) return fail(ctx, RFB_E_INVALID, "nfa belongs to another context");
    if (trace_entries == 0) return fail(ctx, RFB_E_INVALID, "trace_entries must be >= 1");
    cudaSetDevice(ctx->device);
    (void)cudaGetLastError();
    const Nfa &h = nfa->host;
    // cost(s): cycles the FSM spends on an active state (Design/FPGA.v:158-743):
    //   state 0 + state 1 + state 2 (two visits when s % 4 == 3: row_ptr[s+1] is on the next line, :187-205)
    //   + state 3 (1 for an accepting state, else lines + 2 for the 3-deep line pipeline and its drain) + state 4
    std::vector<uint32_t> cost(h.n_states);
    const uint32_t *rp = h.row_ptr();
    for (uint32_t s = 0; s < h.n_states; s++) {
        const uint64_t deg = rp[s + 1] - rp[s];
        const uint64_t s3 = deg == 0 ? 1 : ((((uint64_t)h.n_states + 1 + rp[s]) % 4 + deg + 3) / 4 + 2);
        cost[s] = (uint32_t)(1 + 1 + ((s % 4 == 3) ? 2 : 1) + s3 + 1);
    }
    const uint32_t n_steps = trace_entries - 1;          // the last entry is loaded but never processed (TB:71-86)
    uint32_t *d_cost = nullptr; uint8_t *d_tr = nullptr; unsigned long long *d_total = nullptr;
    cudaError_t e;
    int rc = RFB_OK;
    if ((e = cudaMalloc(reinterpret_cast<void **>(&d_cost), cost.size() * 4)) != cudaSuccess ||
        (e = cudaMalloc(reinterpret_cast<void **>(&d_tr), 2 * (size_t)trace_entries + 16)) != cudaSuccess ||
        (e = cudaMalloc(reinterpret_cast<void **>(&d_total), 8)) != cudaSuccess ||
        (e = cudaMemcpyAsync(d_cost, cost.data(), cost.size() * 4, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess ||
        (e = cudaMemcpyAsync(d_tr, lo, trace_entries, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess ||
        (e = cudaMemcpyAsync(d_tr + trace_entries, hi, trace_entries, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess ||
        (e = launch_tb_cycles(nfa->full, d_cost, d_tr, d_tr + trace_entries, n_steps, d_total, ctx->stream)) != cudaSuccess ||
        (e = cudaMemcpyAsync(cycles, d_total, 8, cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess ||
        (e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess)
        rc = cuda_fail(ctx, e, "rfb_fpga_cycles");
    cudaFree(d_cost); cudaFree(d_tr); cudaFree(d_total);
    return rc;
}

// ---- multi-GPU in one process (BASELINE config 4) ------------------------------------------
// A group is N contexts (one per GPU) plus one NCCL communicator per GPU (ncclCommInitAll).  rfb_group_scan cuts a
// HOST batch into N contiguous stream shards (streams are independent: each starts from {0}, Design/FPGA.v:146-147,
// and the two streams of the reference share only reads of the CSR, FPGA.v:54-57,264-268), scans every shard on its
// GPU concurrently (one host thread per GPU, rfb_scan's chunked copy/scan overlap per GPU), then
//   * per-state counts: ONE ncclAllReduce(sum, u64) over the device count vectors -- the path's only collective;
//   * records: every shard's (already canonical) records are copied from its GPU straight to their place in the caller's
//     buffer -- shards are ascending stream ranges, so the concatenation is the canonical order; the per-shard record
//     counts are host values in a single process, so no collective is needed for them.
// NCCL is loaded with dlopen when the first group is created: the library itself does not depend on it.
}  // extern "C"

namespace {
typedef struct ncclComm *nccl_comm_t;
struct NcclApi {
    void *so = nullptr;
    int (*CommInitAll)(nccl_comm_t *, int, const int *) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok() const { return CommInitAll && CommDestroy && AllReduce && GroupStart && GroupEnd; }
};
constexpr int NCCL_UINT64 = 5, NCCL_SUM = 0;          // ncclUint64 / ncclSum (nccl.h; stable since NCCL 2.0)

NcclApi &nccl_api() {
    static NcclApi api = [] {
        NcclApi a;
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
            a.so = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (a.so) break;
        }
        if (a.so) {
            a.CommInitAll = reinterpret_cast<decltype(a.CommInitAll)>(dlsym(a.so, "ncclCommInitAll"));
            a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(a.so, "ncclCommDestroy"));
            a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(dlsym(a.so, "ncclAllReduce"));
            a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(dlsym(a.so, "ncclGroupStart"));
            a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(dlsym(a.so, "ncclGroupEnd"));
            a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(a.so, "ncclGetErrorString"));
        }
        return a;
    }();
    return api;
}
}  // namespace

struct rfb_group {
    std::vector<rfb_ctx *> ctx;
    std::vector<nccl_comm_t> comm;
    std::string err;
};
struct rfb_group_nfa {
    rfb_group *group = nullptr;
    std::vector<rfb_nfa *> nfa;
};

static int group_fail(rfb_group *g, int code, const std::string &msg) {
    g_err = msg;
    if (g) g->err = msg;
    return code;
}

extern "C" {

int rfb_group_create(const int *device_ids, int n, rfb_group **out) {
    if (!out || !device_ids || n < 1 || n > 64) return fail(nullptr, RFB_E_INVALID, "rfb_group_create: bad arguments");
    *out = nullptr;
    for (int i = 0; i < n; i++) for (int j = 0; j < i; j++)
        if (device_ids[i] == device_ids[j]) return fail(nullptr, RFB_E_INVALID, "rfb_group_create: a device is listed twice");
    NcclApi &api = nccl_api();
    if (!api.ok()) return fail(nullptr, RFB_E_UNSUPPORTED, "NCCL (libnccl.so.2) could not be loaded: multi-GPU groups are unavailable");
    rfb_group *g = new (std::nothrow) rfb_group();
    if (!g) return fail(nullptr, RFB_E_NOMEM, "out of host memory");
    for (int i = 0; i < n; i++) {
        rfb_ctx *c = nullptr;
        const int rc = rfb_ctx_create(device_ids[i], &c);
        if (rc) { rfb_group_destroy(g); return rc; }
        g->ctx.push_back(c);
    }
    g->comm.assign((size_t)n, nullptr);
    const int nrc = api.CommInitAll(g->comm.data(), n, device_ids);
    if (nrc != 0) {
        const std::string msg = std::string("ncclCommInitAll: ") + (api.GetErrorString ? api.GetErrorString(nrc) : "failed");
        g->comm.clear();
        rfb_group_destroy(g);
        return fail(nullptr, RFB_E_CUDA, msg);
    }
    *out = g;
    return RFB_OK;
}

void rfb_group_destroy(rfb_group *g) {
    if (!g) return;
    for (nccl_comm_t c : g->comm) if (c) nccl_api().CommDestroy(c);
    for (rfb_ctx *c : g->ctx) rfb_ctx_destroy(c);
    delete g;
}

int rfb_group_size(const rfb_group *g) { return g ? (int)g->ctx.size() : 0; }
rfb_ctx *rfb_group_ctx(rfb_group *g, int i) { return (g && i >= 0 && i < (int)g->ctx.size()) ? g->ctx[(size_t)i] : nullptr; }
const char *rfb_group_last_error(const rfb_group *g) { return g ? g->err.c_str() : g_err.c_str(); }

int rfb_group_nfa_from_entries(rfb_group *g, const uint32_t *entries, size_t n_entries, int64_t n_states, rfb_group_nfa **out) {
    if (!g || !entries || !out) return fail(nullptr, RFB_E_INVALID, "NULL argument");
    *out = nullptr;
    rfb_group_nfa *gn = new (std::nothrow) rfb_group_nfa();
    if (!gn) return group_fail(g, RFB_E_NOMEM, "out of host memory");
    gn->group = g;
    for (rfb_ctx *c : g->ctx) {                        // the plan is built once and cached (plan_build_cached); every GPU gets a copy
        rfb_nfa *nfa = nullptr;
        const int rc = rfb_nfa_from_entries(c, entries, n_entries, n_states, &nfa);
        if (rc) { g->err = c->err; rfb_group_nfa_destroy(gn); return rc; }
        gn->nfa.push_back(nfa);
    }
    *out = gn;
    return RFB_OK;
}

int rfb_group_nfa_load_coe(rfb_group *g, const char *path, int64_t n_states, rfb_group_nfa **out) {
    if (!g || !path || !out) return fail(nullptr, RFB_E_INVALID, "NULL argument");
    std::vector<uint32_t> e;
    std::string err;
    const int rc = coe_parse_file(path, e, err);
    if (rc) return group_fail(g, rc, err);
    return rfb_group_nfa_from_entries(g, e.data(), e.size(), n_states, out);
}

void rfb_group_nfa_destroy(rfb_group_nfa *gn) {
    if (!gn) return;
    for (rfb_nfa *n : gn->nfa) rfb_nfa_destroy(n);
    delete gn;
}

rfb_nfa *rfb_group_nfa_member(rfb_group_nfa *gn, int i) { return (gn && i >= 0 && i < (int)gn->nfa.size()) ? gn->nfa[(size_t)i] : nullptr; }

int rfb_group_scan(rfb_group *g, const rfb_group_nfa *gn, const rfb_batch *b, uint32_t flags, rfb_result *res) {
    if (!g || !gn || !b || !res) return fail(nullptr, RFB_E_INVALID, "NULL argument");
    if (gn->group != g) return group_fail(g, RFB_E_INVALID, "nfa belongs to another group");
    if (b->offsets || b->steps || b->state_in || b->state_out)
        return group_fail(g, RFB_E_UNSUPPORTED, "rfb_group_scan takes uniformly strided batches without per-stream lengths or resumed state");
    if (flags & ~(RFB_SCAN_SORT_RECORDS | RFB_SCAN_FORCE_WARP | RFB_SCAN_NO_COUNTS))
        return group_fail(g, RFB_E_INVALID, "rfb_group_scan: unsupported flag");
    int rc = check_batch(g->ctx[0], b, true);
    if (rc) { g->err = g->ctx[0]->err; return rc; }
    const size_t N = g->ctx.size();
    const uint32_t n_states = gn->nfa[0]->host.n_states;
    const bool want_counts = res->counts && !(flags & RFB_SCAN_NO_COUNTS);
    // contiguous, balanced shards: GPU i scans streams [n*i/N, n*(i+1)/N)
    std::vector<rfb_batch> sb(N);
    std::vector<rfb_result> sr(N);
    std::vector<int> src(N, RFB_OK);
    std::vector<std::vector<uint64_t>> scratch(N);
    for (size_t i = 0; i < N; i++) {
        const uint64_t first = b->n_streams * i / N, last = b->n_streams * (i + 1) / N;
        sb[i] = *b;
        sb[i].n_streams = last - first;
        sb[i].data = b->data ? b->data + first * b->stride : nullptr;
        sb[i].data_bytes = last > first ? std::min<uint64_t>(b->data_bytes - first * b->stride, (last - first - 1) * b->stride + std::max<uint64_t>(b->n_steps, std::min<uint64_t>(b->stride, b->data_bytes - (last - 1) * b->stride))) : 0;
        sb[i].stream_id_base = b->stream_id_base + (uint32_t)first;
        std::memset(&sr[i], 0, sizeof(rfb_result));
        if (want_counts) { scratch[i].assign(n_states, 0); sr[i].counts = scratch[i].data(); }
        sr[i].records = res->records;                  // only its capacity matters: SCAN_KEEP_ON_DEVICE leaves the records on the GPU
        sr[i].record_capacity = res->records ? res->record_capacity : 0;
    }
    std::vector<std::thread> th;
    for (size_t i = 0; i < N; i++)
        th.emplace_back([&, i] { src[i] = rfb_scan(g->ctx[i], gn->nfa[i], &sb[i], flags | SCAN_KEEP_ON_DEVICE, &sr[i]); });
    for (auto &t : th) t.join();
    for (size_t i = 0; i < N; i++) if (src[i]) { g->err = g->ctx[i]->err; g_err = g->err; return src[i]; }
    // ---- the path's one collective: SUM all-reduce of the per-state counts over NCCL ----
    if (want_counts) {
        NcclApi &api = nccl_api();
        int nrc = api.GroupStart();
        for (size_t i = 0; i < N && nrc == 0; i++) {
            cudaSetDevice(g->ctx[i]->device);
            nrc = api.AllReduce(g->ctx[i]->d_counts, g->ctx[i]->d_counts, n_states, NCCL_UINT64, NCCL_SUM, g->comm[i], g->ctx[i]->stream);
        }
        const int erc = api.GroupEnd();
        if (nrc == 0) nrc = erc;
        if (nrc != 0) return group_fail(g, RFB_E_CUDA, std::string("ncclAllReduce: ") + (api.GetErrorString ? api.GetErrorString(nrc) : "failed"));
        cudaSetDevice(g->ctx[0]->device);
        cudaError_t e = cudaMemcpyAsync(res->counts, g->ctx[0]->d_counts, (size_t)n_states * 8, cudaMemcpyDeviceToHost, g->ctx[0]->stream);
        if (e != cudaSuccess) return group_fail(g, RFB_E_CUDA, std::string("D2H of the reduced counts: ") + cudaGetErrorString(e));
    }
    // ---- records: shard by shard to their place in the caller's buffer ----
    uint64_t n_matches = 0, n_symbols = 0, n_rescanned = 0, at = 0;
    uint32_t launches = 0;
    float ms = 0.f;
    for (size_t i = 0; i < N; i++) {
        n_matches += sr[i].n_matches; n_symbols += sr[i].n_symbols; n_rescanned += sr[i].n_rescanned;
        launches += sr[i].n_launches; ms = std::max(ms, sr[i].gpu_ms);
        const uint64_t room = res->records ? res->record_capacity - at : 0;
        const uint64_t take = std::min<uint64_t>(sr[i].n_records, room);
        if (take) {
            cudaSetDevice(g->ctx[i]->device);
            cudaError_t e = cudaMemcpyAsync(res->records + at, g->ctx[i]->d_records, take * sizeof(rfb_match), cudaMemcpyDeviceToHost, g->ctx[i]->stream);
            if (e != cudaSuccess) return group_fail(g, RFB_E_CUDA, std::string("D2H of the records: ") + cudaGetErrorString(e));
            at += take;
        }
    }
    for (size_t i = 0; i < N; i++) {
        cudaSetDevice(g->ctx[i]->device);
        cudaError_t e = cudaStreamSynchronize(g->ctx[i]->stream);
        if (e != cudaSuccess) return group_fail(g, RFB_E_CUDA, std::string("rfb_group_scan: ") + cudaGetErrorString(e));
    }
    res->n_matches = n_matches; res->n_records = at; res->n_dropped = n_matches - at;
    res->n_symbols = n_symbols; res->n_rescanned = n_rescanned; res->gpu_ms = ms; res->n_launches = launches;
    return RFB_OK;
}

}  // extern "C"
