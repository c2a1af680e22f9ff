// scan.cu -- the two scan kernels (sm_100a).
//
//  scan_lane_kernel : the hot path.  One THREAD per stream, 1024 threads per SM, the whole execution
//      image (image.cpp) staged into shared memory with bulk async copies (cp.async.bulk -> UBLKCP),
//      per-stream state = sticky bit mask and start-DFA state in registers (the DFA's tables are in global
//      memory, L1/L2-resident) + a short ring of transient state ids in shared memory (column-major,
//      bank-conflict-free).  Every lane walks its own stream at its own
//      pace (no per-symbol warp synchronisation): the loop is flattened so that one iteration costs
//      one random shared-memory lookup per lane.  Input bytes arrive as 16-byte ld.global.nc chunks
//      held in registers.
//  scan_warp_kernel : the general path.  One WARP per stream, original CSR read from global memory
//      (L2-resident), next-set de-duplicated in a shared-memory bit vector with atomicOr, long rows
//      expanded cooperatively by the 32 lanes.  Handles any NFA and any activity level; also re-runs
//      the (rare) streams whose transient list overflowed in the lane kernel.
//
// What both compute, per stream (Design/FPGA.v:158-407, 717-765; testbench_BLK_Mem.sv:53-69):
//   S_0 = {0};  for k in [0, n_steps):  report (stream, k, s) for every zero-out-degree s in S_k;
//   S_{k+1} = { t : (sym, t) in row(s), s in S_k, sym == data[k] }.
#include "device.h"
#include <cstdint>
#include <cstdlib>
#include <algorithm>


namespace rfb {

// ------------------------------------------------------------------------------------------------
// shared helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void emit_match(const OutDev &out, uint32_t stream, uint32_t pos, uint32_t state) {
    unsigned long long slot = atomicAdd(&out.g->n_matches, 1ull);
    if (out.records != nullptr && slot < out.capacity) {
        rfb_match m;
        m.stream = stream; m.pos = pos; m.state = state;
        out.records[slot] = m;
    }
    if (out.counts != nullptr) atomicAdd(&out.counts[state], 1ull);
}

__device__ __forceinline__ const uint8_t *stream_ptr(const BatchDev &b, unsigned long long s) {
    return b.data + (b.offsets ? b.offsets[s] : s * b.stride);
}

// ------------------------------------------------------------------------------------------------
// bulk async copy of the image into shared memory (TMA 1-D bulk copy, completes on an mbarrier)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void stage_image(uint8_t *dst, const uint8_t *src, uint32_t bytes, uint8_t *dst2, const uint8_t *src2, uint32_t bytes2, uint64_t *bar) {
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes + bytes2) : "memory");
        const uint32_t CH = 32768;
        for (uint32_t o = 0; o < bytes; o += CH) {
            uint32_t n = bytes - o < CH ? bytes - o : CH;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(dst + o)), "l"(src + o), "r"(n), "r"(smem_u32(bar)) : "memory");
        }
        for (uint32_t o = 0; o < bytes2; o += CH) {
            uint32_t n = bytes2 - o < CH ? bytes2 - o : CH;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(dst2 + o)), "l"(src2 + o), "r"(n), "r"(smem_u32(bar)) : "memory");
        }
    }
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// lane kernel
// ------------------------------------------------------------------------------------------------
// Shared memory of a lane-kernel CTA: image | hottest start-DFA rows | per-stream rings | barrier | class / attention copies.
// The rows are the first ones of the start-DFA table -- the library keeps that table ordered by measured visit frequency
// (api.cu: calibration), and a row in shared memory costs a bank-conflict-limited gather instead of one L1 wavefront per
// lane.
// tables a CTA builds in shared memory at kernel start for the quiet run: byte-wide symbol classes (256 B) and the
// attention masks at their natural stride (256 x 16 B reserved; 8 B used with a one-word mask)
constexpr size_t LANE_AUX_BYTES = 256 + 256 * 16;
int lane_ring_cap(const ImageHeader &h) {
    // 16 entries per stream: with the start DFA and the look-ahead masks no stream of the bench mixes (adversarial included)
    // ever holds more transient states at once; a fuller stream goes to the general kernel, as with any capacity.  The 32 KB
    // a 32-entry ring would add are worth more as L1 (below).
    static const int max_cap = [] { const char *e = getenv("RFB_RING_CAP"); const int v = e ? atoi(e) : 16; return v >= 64 ? 64 : v >= 32 ? 32 : 16; }();
    for (int cap = max_cap; cap >= 16; cap >>= 1)
        if ((size_t)h.blob_bytes + (size_t)cap * LANE_THREADS * 2 + 16 + LANE_AUX_BYTES <= MAX_DYN_SMEM) return cap;
    return 0;
}
// rows of the start-DFA table staged into shared memory, and the (16-byte multiple) bytes copied for them
uint32_t lane_hot_rows(const ImageHeader &h, uint32_t *copy_bytes) {
    static const long max_rows = [] { const char *e = getenv("RFB_HOT_ROWS"); return e ? atol(e) : 1L << 30; }();
    // 6 KB of every SM stay free: the record sort of the previous batch (sort.cu: 1 KB static + 1 KB reserved per CTA) must
    // be able to run beside a lane-kernel CTA, or the pipelined host path (rfb_scan_submit / _wait) stalls behind the scan
    constexpr size_t CORESIDENT_RESERVE = 6 * 1024;
    // Shared memory and L1 share 256 KB per SM in fixed splits (... 164 / 196 / 228 KB of shared memory).  The rows that do not
    // fit shared memory are served by L1 / L2, and a 60 KB L1 serves them far better than a 28 KB one (measured: W-mix
    // 8.69 -> 8.23 ms with the same 333 rows): the kernel stays inside the 196 KB split when that still leaves room for a
    // useful number of rows, and only takes the 228 KB split when the image is too large for that.
    constexpr size_t SPLIT_196 = 196 * 1024 - 1024 - CORESIDENT_RESERVE;   // - the CTA's reserved KB
    const size_t row = (size_t)std::max<uint32_t>(1u, h.dfa_ncls) * 2;
    const size_t used = (size_t)h.blob_bytes + (size_t)lane_ring_cap(h) * LANE_THREADS * 2 + 16 + LANE_AUX_BYTES;
    const size_t limit = used + 64 * row <= SPLIT_196 ? SPLIT_196 : MAX_DYN_SMEM - CORESIDENT_RESERVE;
    const size_t avail = used < limit ? (limit - used) & ~(size_t)15 : 0;
    size_t rows = std::min<size_t>(std::max<uint32_t>(1u, h.dfa_states), avail / row);
    if ((long)rows > max_rows) rows = (size_t)std::max<long>(0, max_rows);
    if (copy_bytes) *copy_bytes = (uint32_t)((rows * row + 15) & ~(size_t)15);
    return (uint32_t)rows;
}
size_t lane_smem_bytes(const ImageHeader &h) {
    uint32_t hot_bytes = 0;
    lane_hot_rows(h, &hot_bytes);
    return (size_t)h.blob_bytes + hot_bytes + (size_t)lane_ring_cap(h) * LANE_THREADS * 2 + 16 + LANE_AUX_BYTES;   // + barrier + the tables built at kernel start
}

// explicit shared-window accesses: 32-bit shared addresses never go through generic-pointer conversion.
// Table loads are plain asm (read-only data, the compiler may schedule them freely); ring accesses are volatile.
__device__ __forceinline__ uint32_t lds32(uint32_t a) { uint32_t v; asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds16(uint32_t a) { uint16_t v; asm("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint4 lds128(uint32_t a) { uint4 v; asm("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v; }
__device__ __forceinline__ uint2 lds64(uint32_t a) { uint2 v; asm("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t ring_ld(uint32_t a) { uint16_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void ring_st(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((uint16_t)v) : "memory"); }

// out-of-line copies for the lane kernel's hot loop (both are rare there)
__device__ __noinline__ void emit_match_cold(const OutDev &out, uint32_t stream, uint32_t pos, uint32_t state) {
    emit_match(out, stream, pos, state);
}
// exact duplicate check of a candidate against this step's new ring entries
__device__ __noinline__ bool ring_contains(uint32_t lb, uint32_t from, uint32_t to, uint32_t row, uint32_t rmask, uint32_t t) {
    for (uint32_t o = from; o != to; o = (o + row) & rmask)
        if (ring_ld(lb + o) == t) return true;
    return false;
}

// ---- resumable scans: per-stream active sets as original state ids (rare paths, kept out of line) ----
// S_{n_steps} of a finished stream = sticky bits + the ring's new entries + the never-materialised members of the
// stream's start-DFA state d (those that are not in the ring as well).
__device__ __noinline__ void lane_export_state(const uint32_t *orig_of_id, const uint32_t *mem_ptr, const uint16_t *mem_ids,
                                               unsigned int *dst, uint32_t cap, bool append, uint64_t P0, uint64_t P1, uint32_t lb,
                                               uint32_t from, uint32_t to, uint32_t row, uint32_t rmask, uint32_t d) {
    uint32_t n = append ? dst[0] : 0u;
    if (n == 0xFFFFFFFFu) return;                      // an earlier part's overflow mark stays
    for (int w = 0; w < 2; w++) {
        uint64_t bits = w ? P1 : P0;
        while (bits) {
            const uint32_t b = (uint32_t)__ffsll((long long)bits) - 1u + 64u * w;
            bits &= bits - 1;
            if (n < cap) dst[1 + n] = orig_of_id[b];
            n++;
        }
    }
    for (uint32_t o = from; o != to; o = (o + row) & rmask) {
        if (n < cap) dst[1 + n] = orig_of_id[ring_ld(lb + o)];
        n++;
    }
    if (d) for (uint32_t j = mem_ptr[d]; j < mem_ptr[d + 1]; j++) {
        const uint32_t id = mem_ids[j];
        if (ring_contains(lb, from, to, row, rmask, id)) continue;
        if (n < cap) dst[1 + n] = orig_of_id[id];
        n++;
    }
    dst[0] = n <= cap ? n : 0xFFFFFFFFu;
}
// a stream without steps keeps its state
__device__ __noinline__ void carry_state(const unsigned int *src, unsigned int *dst, uint32_t cap) {
    if (!src) { dst[0] = 1; dst[1] = 0; return; }
    const uint32_t n = src[0];
    dst[0] = n;
    for (uint32_t i = 0; i < n && i < cap; i++) dst[1 + i] = src[1 + i];
}

// input chunks: plain (coherent) 16-byte loads -- in the chunk-gated host path the copy engine is still writing LATER
// chunks of the same buffer while the kernel runs, so the non-coherent (.nc) path is not used for the batch; chunk
// boundaries are 128-byte aligned (api.cu: plan_chunks), so no cache line ever straddles a chunk that has not landed
__device__ __forceinline__ uint4 ld_in(const uint8_t *p) {
    uint4 v; asm("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p)); return v;
}
// first loads of a stream: ordered after the acquire of the chunk's arrival counter
__device__ __forceinline__ uint4 ld_in_ordered(const uint8_t *p) {
    uint4 v; asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ unsigned int ld_acquire(const unsigned int *p) {
    unsigned int v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
// start-DFA entry, sign-extended: negative <=> the transition has an insertion list
__device__ __forceinline__ int ldg_s16(const uint16_t *p) { int v; asm("ld.global.nc.s16 %0, [%1];" : "=r"(v) : "l"(p)); return v; }
// start-DFA entry `at` (= d * ncls + class) of state d: rows below hot_rows from shared memory, the rest from global
// memory; predicated, no branch.  Used by the general step; the quiet run reads the entry with one generic load from a
// selected base instead (lane_body: RFB_LD_DFA_Q), which is two instructions shorter.
__device__ __forceinline__ int ld_dfa(uint32_t d, uint32_t hot_rows, uint32_t hot_s, uint32_t at, const uint16_t *base) {
    int v = 0;   // defined on every path as far as ptxas can tell (the two loads are complementary)
    asm("{\n\t.reg .pred p;\n\t.reg .u64 a;\n\t.reg .u32 sa;\n\t"
        "setp.lt.u32 p, %1, %2;\n\t"
        "mad.lo.u32 sa, %4, 2, %3;\n\t"
        "mad.wide.u32 a, %4, 2, %5;\n\t"
        "@p ld.shared.s16 %0, [sa];\n\t"
        "@!p ld.global.nc.s16 %0, [a];\n\t}"
        : "+r"(v) : "r"(d), "r"(hot_rows), "r"(hot_s), "r"(at), "l"(base));
    return v;
}
__device__ __forceinline__ uint32_t lds8r(uint32_t a) { uint32_t v; asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds16r(uint32_t a) { uint32_t v; asm("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ int lds_s16(uint32_t a) { int v; asm("ld.shared.s16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }

#ifndef RFB_QUIET_STEPS
#define RFB_QUIET_STEPS 16   // symbols per quiet run (<= 16: one input chunk)
#endif
#ifndef RFB_STEP_REPS
#define RFB_STEP_REPS 2      // general steps a busy stream may take per iteration (measured: 1 / 2 / 4 / 8, profiles/README.md)
#endif

// v >>= 8 * nb (nb in 0..15), branch-free
__device__ __forceinline__ void shr_bytes(uint4 &v, uint32_t nb) {
    const bool b8 = (nb & 8u) != 0, b4 = (nb & 4u) != 0;
    uint32_t x = b8 ? v.z : v.x, y = b8 ? v.w : v.y, z = b8 ? 0u : v.z, w = b8 ? 0u : v.w;
    x = b4 ? y : x; y = b4 ? z : y; z = b4 ? w : z; w = b4 ? 0u : w;
    const uint32_t sh = (nb & 3u) * 8u;
    v.x = __funnelshift_r(x, y, sh); v.y = __funnelshift_r(y, z, sh); v.z = __funnelshift_r(z, w, sh); v.w = w >> sh;
}

// The body of the lane kernel.  q_next / q_rescan_n / q_rescan: the stream-fetch counter and the hand-over queue this
// CTA works on (the batch's own, or -- when several parts of a cut NFA run in one launch -- those of the CTA's part).
template <int W, int RING_CAP>
__device__ __forceinline__ void lane_body(const NfaDev &nfa, const BatchDev &batch, const OutDev &out, unsigned int *q_next,
                                          unsigned int *q_rescan_n, uint2 *q_rescan, uint8_t *smem) {
    const ImageHeader &h = nfa.h;
    constexpr uint32_t ROW = LANE_THREADS * 2;          // bytes between consecutive ring entries of one lane
    constexpr uint32_t RING = RING_CAP * ROW;
    constexpr uint32_t RMASK = RING - 1;
    constexpr uint32_t NONE = 0xFFFFFFFFu;
    // shared memory: image | first hot_rows rows of the start-DFA table | rings | barrier
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + h.blob_bytes + nfa.hot_bytes + RING);
    stage_image(smem, nfa.blob, h.blob_bytes, smem + h.blob_bytes, reinterpret_cast<const uint8_t *>(nfa.dfa_dt), nfa.hot_bytes, bar);

    // 32-bit shared-window addresses of the staged tables and of this lane's ring
    // ptxas re-derives the shared window base (S2R CgaCtaId + LEA) and threadIdx at every use to save a register;
    // routing both through a shuffle makes them opaque, so they stay in registers
    const uint32_t sbase = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
    // fixed-size tables sit at offsets that depend only on W (image.cpp): immediates in the load instructions
    constexpr uint32_t OFF_CMAP = 256u * 32u * W, OFF_SDESC = OFF_CMAP + 1024u, OFF_TAB = OFF_SDESC + 64u * W * 4u;
    const uint32_t mask_s = sbase, cmap_s = sbase + OFF_CMAP, sdesc_s = sbase + OFF_SDESC, tab_s = sbase + OFF_TAB;
    const uint32_t memb_s = sbase + h.off_memb, look_s = sbase + h.off_look;
    const uint32_t hot_s = sbase + h.blob_bytes, hot_rows = nfa.hot_rows;
    // The quiet run reads its start-DFA entry with ONE generic load from a selected base (the hot rows' shared window or the
    // table in global memory): ISETP + 2 SEL + IMAD.WIDE + LD instead of the predicated LDS / LDG pair with its two address
    // computations -- two instructions fewer per symbol (uniform bytes 1.38 -> 1.29 ms, lo windows 3.44 -> 3.37)
    const int16_t *hot_g = reinterpret_cast<const int16_t *>(smem + h.blob_bytes);
    const int16_t *dt_g = reinterpret_cast<const int16_t *>(nfa.dfa_dt);
#define RFB_LD_DFA_Q(D, AT) ((int)((D) < hot_rows ? hot_g : dt_g)[(AT)])
    const uint32_t lb = __shfl_sync(0xffffffffu, sbase + h.blob_bytes + nfa.hot_bytes + threadIdx.x * 2, threadIdx.x & 31);   // ring entry at byte offset o: lb + o; bank-conflict free
    // The quiet run looks the class of every symbol up: a byte-wide copy of that table (256 bytes = 64 words, so lanes that
    // read the same word share one access and at most two words share a bank) instead of the 32-bit cmap entries (256
    // words: ~3.5 conflicting accesses per warp-wide lookup -- on quiet traffic the shared-memory pipe was 91 % busy).
    const uint32_t cls8_s = sbase + h.blob_bytes + nfa.hot_bytes + RING + 16;
    // ... and its attention masks are read for every symbol by every lane of a warp that holds any sticky state: in the
    // image they sit in 32-byte (64-byte) rows together with K and M, i.e. on 4 (2) distinct bank groups, which makes a
    // warp-wide lookup an ~8-way conflict (63 % of all shared-memory wavefronts on W-mix); a copy at stride 8 (16) bytes
    // spreads them over all banks.
    const uint32_t att_s = cls8_s + 256;
    if (threadIdx.x < 256) {
        const uint32_t cm = lds32(cmap_s + threadIdx.x * 4);
        asm volatile("st.shared.u8 [%0], %1;" ::"r"(cls8_s + threadIdx.x), "r"(cm & 0xFFu) : "memory");
        if (W == 1) {
            const uint2 a = lds64(mask_s + threadIdx.x * 32u * W);
            asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(att_s + threadIdx.x * 8), "r"(a.x), "r"(a.y) : "memory");
        } else {
            const uint4 a = lds128(mask_s + threadIdx.x * 32u * W);
            asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(att_s + threadIdx.x * 16), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w) : "memory");
        }
    }
    __syncthreads();
    const uint32_t gbase = h.gbase, nsb = h.nsb;
    const uint32_t nbm = (1u << h.bucket_bits) - 1u;
    const uint32_t acc_base = h.acc_base, n_acc = h.n_acc, ncls = h.dfa_ncls;
    const uint16_t *__restrict__ dfa_dt = nfa.dfa_dt;
    const uint32_t abit = h.accel ? 1u : 0u;            // without a start DFA the table is one inert entry and d stays 0
    constexpr uint32_t MSTRIDE = 32u * W;
    constexpr uint32_t CM_OFF = W == 1 ? 8u : 48u;      // {class, hash} of symbol c in the padding of its mask row

    // One THREAD per stream, every thread at its own pace.  An iteration of the loop offers each lane two blocks and
    // the lane takes the one its state asks for:
    //   QUIET  a lane whose transient set is empty runs up to 16 symbols straight from its input registers (bytes at
    //          static positions) with one start-DFA lookup and, if anybody in the warp holds a sticky state besides A,
    //          one attention test per symbol -- the state in which most symbols of most streams are scanned; the run
    //          stops in front of the first symbol that needs anything else (flagged DFA transition, a sticky state
    //          dying, a sticky state firing into something that can outlive the next symbol);
    //   STEP   one whole symbol step the general way: finish the stream / take the next one, open the step (start
    //          DFA, sticky masks), drain every work item of it (entries of a start-DFA insertion list, members of the
    //          current set, rows of firing sticky states: one edge-table lookup and one insertion each), close it
    //          (current <= next, Design/FPGA.v:733-737).
    // All per-step bookkeeping is local to STEP, so the quiet run keeps few registers alive.
    uint64_t P0 = 0, P1 = 0;                            // sticky set (P1 unused when W == 1)
    uint32_t rp = 0, re = 0;                            // ring byte offsets of the current set S_k: [rp, re)
    uint32_t d = 0;                                     // start-DFA state: 0 = A not active yet, 1 = A alone
    uint4 cur = make_uint4(0, 0, 0, 0);                 // the next nv bytes of the stream, next symbol in cur.x[7:0]
    uint4 pre = make_uint4(0, 0, 0, 0);                 // the aligned 16-byte chunk after them
    uint32_t nv = 0;
    const uint8_t *np = nullptr;                        // address of pre
    uint32_t sid = 0, k = 0, nsteps = 0;
    bool have = false;
    bool evt = false;                                   // the quiet run stopped in front of an event: STEP takes that symbol
    bool done = false;                                  // no stream left for this lane
    bool parked = false;                                // host path: the lane holds a stream whose input chunk has not landed yet
#ifdef RFB_STATS
    unsigned long long st_iter = 0, st_qruns = 0, st_qsteps = 0, st_qzero = 0, st_steps = 0, st_items = 0, st_qfast = 0;
#define RFB_STAT(x) x
#else
#define RFB_STAT(x)
#endif

    for (;;) {
        // All 32 lanes stay in the loop until the last one is done, and meet here once per iteration: both blocks are
        // entered by converged lanes (a lane that leaves a block early waits for the others at the next meeting point
        // instead of running ahead through private copies of the code).
        if (__all_sync(0xffffffffu, done)) break;
        if (batch.chunk_streams && __all_sync(0xffffffffu, done || parked)) __nanosleep(500);   // nothing to do but wait for the copy
        RFB_STAT(if (!done) st_iter++;)
#define RFB_NEXT_CHUNK()                                                                        \
        do {                                                                                     \
            cur = pre; nv = 16u;                                                                 \
            if (nsteps - k > 16u) { np += 16; pre = ld_in(np); }   /* the stream goes on behind it */ \
        } while (0)
        if (have && nv == 0u) RFB_NEXT_CHUNK();
        // ---- QUIET run: to the end of the bytes at hand, or to the first symbol with an event ----
        if (have && !evt && rp == re && nv != 0u && k < nsteps) {
            d = max(d, (uint32_t)P0 & abit);             // A entered the set (it never leaves): 0 -> 1
            const uint32_t plo = (uint32_t)P0, phi = (uint32_t)(P0 >> 32), qlo = (uint32_t)P1, qhi = (uint32_t)(P1 >> 32);
            // a lane whose only sticky state is A (whose edges the start DFA follows) has nothing to attend to; the
            // decision is taken per WARP (one extra load per step for everybody beats a divergent branch for a few)
            const bool needmask = __ballot_sync(__activemask(), ((plo & ~abit) | phi | qlo | qhi) != 0u) != 0u;
            const uint32_t n = min(nv, nsteps - k);
            uint32_t cnt = 0;
            // one step of the run.  JJ: index of the symbol among the bytes at hand; W0 / W1: the registers that hold this
            // symbol and the next one; JB: this symbol's (static) byte position in W0
#define RFB_QUIET_STEP(JJ, JB, W0, W1, CHECK_N)                                                                            \
            {                                                                                                              \
                if (CHECK_N && (uint32_t)(JJ) >= n) break;                                                                 \
                const uint32_t cc = __byte_perm((W0), 0u, 0x4440u | (uint32_t)(JB));   /* PRMT: one instruction */          \
                const uint32_t cls = lds8r(cls8_s + cc);                                                                   \
                uint32_t t = 0;                                                                                            \
                if (needmask) {                                                                                            \
                    const uint32_t mrow = mask_s + cc * MSTRIDE;                                                           \
                    if (W == 1) { const uint2 a = lds64(att_s + cc * 8); t = (plo & a.x) | (phi & a.y); }                  \
                    else { const uint4 a = lds128(att_s + cc * 16); t = (plo & a.x) | (phi & a.y) | (qlo & a.z) | (qhi & a.w); } \
                    if (t != 0u && (uint32_t)(JJ) + 1u < n) {                                                              \
                        /* attention: a state that dies always counts; a state that fires only if what it enters can      \
                           outlive the NEXT symbol (look-ahead masks, image.cpp) -- most firings cannot */                 \
                        const uint32_t cn = __byte_perm((W1), 0u, 0x4440u | (((uint32_t)(JB) + 1u) & 3u));                 \
                        if (W == 1) {                                                                                      \
                            const uint4 km = lds128(mrow + 16);                                                            \
                            const uint2 lk = lds64(look_s + cn * 8);                                                       \
                            t = (plo & ~km.x) | (phi & ~km.y) | (plo & km.z & lk.x) | (phi & km.w & lk.y);                 \
                        } else {                                                                                           \
                            const uint4 kk = lds128(mrow + 16), mm = lds128(mrow + 32), lk = lds128(look_s + cn * 16);     \
                            t = (plo & ~kk.x) | (phi & ~kk.y) | (qlo & ~kk.z) | (qhi & ~kk.w) |                            \
                                (plo & mm.x & lk.x) | (phi & mm.y & lk.y) | (qlo & mm.z & lk.z) | (qhi & mm.w & lk.w);     \
                        }                                                                                                  \
                    }                                                                                                      \
                }                                                                                                          \
                const int e = RFB_LD_DFA_Q(d, d * ncls + cls);                                                             \
                if ((t | ((uint32_t)e & 0x80000000u)) != 0u) break;   /* STEP takes this symbol */                         \
                d = (uint32_t)e;                                                                                           \
                cnt = (uint32_t)(JJ) + 1u;                                                                                 \
            }
            // one version of the run for the whole warp: without per-step limit checks only if EVERY lane in the run has a
            // whole chunk ahead (two versions side by side would run one after the other)
            const bool full = __all_sync(__activemask(), n == (uint32_t)RFB_QUIET_STEPS);
#define RFB_WORD(J) ((J) < 4 ? cur.x : (J) < 8 ? cur.y : (J) < 12 ? cur.z : cur.w)
            if (full) {
#pragma unroll
                for (int J = 0; J < RFB_QUIET_STEPS; J++) RFB_QUIET_STEP(J, J & 3, RFB_WORD(J), RFB_WORD(J + 1), false)
            } else {
#pragma unroll
                for (int J = 0; J < RFB_QUIET_STEPS; J++) RFB_QUIET_STEP(J, J & 3, RFB_WORD(J), RFB_WORD(J + 1), true)
            }
#undef RFB_WORD
#undef RFB_QUIET_STEP
#undef RFB_LD_DFA_Q
            k += cnt; nv -= cnt;
            RFB_STAT(st_qruns++; st_qsteps += cnt; st_qzero += cnt == 0; st_qfast += n == 16u;)
            evt = cnt != n;
            if (nv && cnt) shr_bytes(cur, cnt);
        }
        __syncwarp();
        // ---- STEP: one whole symbol step ----
        const bool want_step = !done && (!have || k == nsteps || (nv != 0u && (evt || rp != re)));
        if (want_step) {
            if (have && k == nsteps) {                   // stream finished
                if (batch.state_out)
                    lane_export_state(nfa.orig_of_id, nfa.dfa_mem_ptr, nfa.dfa_mem_ids, batch.state_out + (size_t)sid * (1u + batch.state_cap),
                                      batch.state_cap, batch.state_append != 0, P0, P1, lb, rp, re, ROW, RMASK, d);
                have = false;
            }
            if (!have) {   // ---- next stream ----
                for (;;) {
                    if (!parked) {
                        sid = atomicAdd(q_next, 1u);
                        if (sid >= batch.n_streams) break;
                        nsteps = batch.steps ? batch.steps[sid] : batch.n_steps;
                        if (nsteps == 0) {
                            if (batch.state_out && !batch.state_append)
                                carry_state(batch.state_in ? batch.state_in + (size_t)sid * (1u + batch.state_cap) : nullptr,
                                            batch.state_out + (size_t)sid * (1u + batch.state_cap), batch.state_cap);
                            continue;
                        }
                        if (batch.steps && batch.count_symbols) atomicAdd(&out.g->n_symbols, (unsigned long long)nsteps);
                    }
                    if (batch.chunk_streams) {   // host path: has the H2D copy of this stream's chunk landed?  If not the lane
                                                 // parks its stream and asks again next iteration: its warp-mates go on
                        parked = ld_acquire(batch.ready) < sid / batch.chunk_streams + 1u;
                        if (parked) break;
                    }
                    P0 = 0; P1 = 0; rp = 0; re = 0; d = 0; k = 0;
                    if (batch.state_in) {   // resume: the stream's active set as left by an earlier call
                        const unsigned int *stt = batch.state_in + (size_t)sid * (1u + batch.state_cap);
                        const uint32_t ns = stt[0] <= batch.state_cap ? stt[0] : 0u;   // an overflow mark cannot be resumed: treated as empty (rfb_scan rejects it)
                        bool fits = true;
                        for (uint32_t q = 0; q < ns; q++) {
                            if (stt[1 + q] >= nfa.n_ref_states) continue;                // not a state: ignored
                            const uint32_t id = nfa.id_of_orig[stt[1 + q]];
                            if (id == 0xFFFFFFFFu) continue;          // a state of another part
                            if (id < nsb) { if (W == 1 || id < 64) P0 |= 1ull << (id & 63); else P1 |= 1ull << (id & 63); }
                            else if (((re + ROW) & RMASK) == rp) fits = false;
                            else { ring_st(lb + re, id); re = (re + ROW) & RMASK; }
                        }
                        if (!fits) {   // more transient members than the ring holds: the general kernel takes the whole stream
                            const unsigned int slot = atomicAdd(q_rescan_n, 1u);
                            q_rescan[slot] = make_uint2(sid, 0u);
                            continue;
                        }
                    } else if (h.start_id < nsb) {                                        // Design/FPGA.v:146-147
                        if (W == 1 || h.start_id < 64) P0 = 1ull << (h.start_id & 63); else P1 = 1ull << (h.start_id & 63);
                    } else { ring_st(lb, h.start_id); re = ROW; }
                    break;
                }
                if (parked) {}
                else if (sid >= batch.n_streams) done = true;   // this lane is done
                else {
                    // input: aligned 16-byte chunks; the first one is shifted down to the stream's first byte
                    const uint8_t *sp = stream_ptr(batch, sid);
                    np = reinterpret_cast<const uint8_t *>(reinterpret_cast<uintptr_t>(sp) & ~(uintptr_t)15);
                    const uint32_t off = (uint32_t)(sp - np);
                    cur = ld_in_ordered(np);
                    shr_bytes(cur, off);
                    nv = 16u - off;
                    if (nsteps > nv) { np += 16; pre = ld_in_ordered(np); }
                    have = true;
                }
            }
            // a stream whose transient set stays non-empty takes its next symbols here as well (up to RFB_STEP_REPS per
            // iteration): busy streams then need fewer iterations, each of which also pays for a quiet block
#pragma unroll 1
            for (int rep = 0; !done && !parked && rep < RFB_STEP_REPS; rep++) {
                if (rep != 0) {
                    if (!(have && k != nsteps && rp != re)) break;
                    if (nv == 0u) RFB_NEXT_CHUNK();
                }
                // ---- open step k: next symbol ----
                RFB_STAT(st_steps++;)
                const uint32_t c = cur.x & 0xFFu;
                cur.x = __funnelshift_r(cur.x, cur.y, 8); cur.y = __funnelshift_r(cur.y, cur.z, 8); cur.z = __funnelshift_r(cur.z, cur.w, 8); cur.w >>= 8;
                nv--;
                const uint32_t mrow = mask_s + c * MSTRIDE;
                const uint4 a = lds128(mrow);           // attention masks (W == 1: | start-DFA class | symbol hash)
                uint32_t cls, hf;
                if (W == 1) { cls = a.z; hf = a.w; }
                else { const uint2 ch = lds64(mrow + CM_OFF); cls = ch.x; hf = ch.y; }
                const uint32_t hc = hf & nbm;           // a hashed row uses the low bits of the symbol hash
                uint32_t x = NONE;                      // next entry of a pending start-DFA insertion list
                // start DFA: one lookup steps all the never-materialised successors of the always-active state A
                {
                    d = max(d, (uint32_t)P0 & abit);
                    const uint32_t at = d * ncls + cls;
                    const int e = ld_dfa(d, hot_rows, hot_s, at, dfa_dt);
                    d = (uint32_t)e & 0x7FFFu;
                    if (e < 0) x = __ldg(nfa.dfa_dta + at);   // sticky / accepting / untracked successors to insert
                }
                // sticky states: survivors P & K[c]; those in P & M[c] fire their rows
                uint32_t i0 = 0, i1 = 0, i2 = 0, i3 = 0;  // firing sticky bits not yet expanded
                {
                    bool attn;
                    if (W == 1) attn = (((uint32_t)P0 & a.x) | ((uint32_t)(P0 >> 32) & a.y)) != 0;
                    else attn = (((uint32_t)P0 & a.x) | ((uint32_t)(P0 >> 32) & a.y) | ((uint32_t)P1 & a.z) | ((uint32_t)(P1 >> 32) & a.w)) != 0;
                    if (attn) {
                        if (W == 1) {
                            const uint4 km = lds128(mrow + 16);
                            i0 = (uint32_t)P0 & km.z; i1 = (uint32_t)(P0 >> 32) & km.w;
                            P0 &= (uint64_t)km.x | ((uint64_t)km.y << 32);
                        } else {
                            const uint4 kk = lds128(mrow + 16);
                            const uint4 mm = lds128(mrow + 32);
                            i0 = (uint32_t)P0 & mm.x; i1 = (uint32_t)(P0 >> 32) & mm.y;
                            i2 = (uint32_t)P1 & mm.z; i3 = (uint32_t)(P1 >> 32) & mm.w;
                            P0 &= (uint64_t)kk.x | ((uint64_t)kk.y << 32);
                            P1 &= (uint64_t)kk.z | ((uint64_t)kk.w << 32);
                        }
                        if (nv != 0u && k + 1u < nsteps) {   // look-ahead: skip firings whose targets cannot outlive the next symbol
                            const uint32_t cn = cur.x & 0xFFu;
                            if (W == 1) { const uint2 lk = lds64(look_s + cn * 8); i0 &= lk.x; i1 &= lk.y; }
                            else { const uint4 lk = lds128(look_s + cn * 16); i0 &= lk.x; i1 &= lk.y; i2 &= lk.z; i3 &= lk.w; }
                        }
                    }
                }
                // ---- drain every work item of the step ----
                uint32_t wp = re;                       // next write: S_{k+1} grows behind S_k
                uint32_t filt = 0;                      // 32-bit membership filter of this step's new entries
                uint32_t idx = 0;
                bool walking = false, ovf = false;
                for (;;) {
                    bool hit = false, look = walking;
                    uint32_t t = 0;
                    if (!walking) {
                        if (x != NONE) {                                  // entry of a start-DFA insertion list: the target itself
                            hit = true;
                            const uint32_t tl = __ldg(nfa.dfa_act + x);
                            t = tl & 0x7FFFu;
                            x = (tl & 0x8000u) ? x + 1 : NONE;
                        } else if (rp != re) {                            // a member of S_k
                            const uint32_t u = ring_ld(lb + rp);
                            rp = (rp + ROW) & RMASK;
                            idx = u + (u >= gbase ? hc : 0u);
                            look = true;
                            if (u - acc_base < n_acc) {                   // accepting (Design/FPGA.v:210-226)
                                if (u != nfa.no_report_lane) emit_match_cold(out, sid + batch.stream_id_base, k + batch.pos_base, nfa.orig_of_id[u]);
                                look = false;
                            }
                        } else if ((i0 | i1 | i2 | i3) != 0u) {          // row of a firing sticky state
                            uint32_t wsel, wbase;
                            if (i0) { wsel = i0; wbase = 0; i0 &= i0 - 1; }
                            else if (i1) { wsel = i1; wbase = 32; i1 &= i1 - 1; }
                            else if (i2) { wsel = i2; wbase = 64; i2 &= i2 - 1; }
                            else { wsel = i3; wbase = 96; i3 &= i3 - 1; }
                            const uint32_t sd = lds32(sdesc_s + (wbase + (uint32_t)__ffs((int)wsel) - 1u) * 4);
                            idx = (sd & 0xFFFFu) + (hf & (sd >> 16));
                            look = true;
                        } else break;                                     // the step is drained
                    }
                    RFB_STAT(st_items++;)
                    if (look) {
                        const uint32_t e = lds32(tab_s + idx * 4);
                        const uint32_t ea = e & 0xFFu, eb = (e >> 8) & 0xFFu;
                        t = (e >> 16) & 0x7FFFu;
                        walking = (e & TAB_MORE) != 0;
                        idx++;
                        if (ea <= eb) hit = (c == ea) | (c == eb);
                        else if (ea == 0xFFu) { idx = t; walking = true; }                     // indirect -> chain
                        else hit = (lds32(memb_s + (((0xFEu - ea) * 253u + eb) * 8 + (c >> 5)) * 4) >> (c & 31)) & 1u;
                    }
                    if (hit && !ovf) {   // ---- the one insertion site: add t to S_{k+1} ----
                        if (t < nsb) {                                    // entering a sticky state
                            const uint64_t sb = 1ull << (t & 63);
                            // P was read when the step opened (survivors, firing bits) and insertions happen after that:
                            // a state entered now is first seen by the next step, as it must be
                            if (W == 1 || t < 64) P0 |= sb; else P1 |= sb;
                        } else {
                            const uint32_t fb = 1u << (t & 31);
                            if (!((filt & fb) && ring_contains(lb, re, wp, ROW, RMASK, t))) {
                                const uint32_t nw = (wp + ROW) & RMASK;
                                if (nw == rp) ovf = true;                 // ring full: hand the stream to the general kernel
                                else {
                                    ring_st(lb + wp, t);
                                    wp = nw;
                                    filt |= fb;
                                }
                            }
                        }
                    }
                }
                // ---- close: current <= next (Design/FPGA.v:733-737); rp == re: S_k is consumed ----
                re = wp; k++; evt = false;
                if (ovf) {   // a full ring hands the stream to the general kernel: S_{k-1} was fully examined, that kernel
                             // re-runs the stream and reports from step k on (after the last step there is nothing left to
                             // report, but the caller may want S_{n_steps}: only that kernel has it)
                    if (k < nsteps || batch.state_out) {
                        const unsigned int slot_ = atomicAdd(q_rescan_n, 1u);
                        q_rescan[slot_] = make_uint2(sid, k);
                    }
                    have = false;
                }
            }
        }
    }
#ifdef RFB_STATS
    if (out.counts) {   // dev build only: loop statistics land in the LAST seven counters of the count vector
        unsigned long long *cs = out.counts + (nfa.n_ref_states - 7);
        atomicAdd(cs + 0, st_iter); atomicAdd(cs + 1, st_qruns); atomicAdd(cs + 2, st_qsteps); atomicAdd(cs + 3, st_qzero);
        atomicAdd(cs + 4, st_steps); atomicAdd(cs + 5, st_items); atomicAdd(cs + 6, st_qfast);
    }
#endif
#undef RFB_STAT
#undef RFB_NEXT_CHUNK
}

// 56 registers, not 58: 1024 threads x 64 (the allocation unit rounds 58 up) would take the whole register file and the
// record sort of the previous batch could not run beside this kernel (pipelined host path: +1.9 ms per batch)
template <int W, int RING_CAP>
__global__ void __maxnreg__(56)
scan_lane_kernel(const NfaDev nfa, const BatchDev batch, const OutDev out) {
    extern __shared__ __align__(128) uint8_t smem[];
    lane_body<W, RING_CAP>(nfa, batch, out, &out.g->next_stream, &out.g->n_rescan, out.rescan, smem);
}

// Several parts of a cut NFA (csrc/parts.cpp) in ONE launch: CTA b scans the whole batch against part b % n_parts, whose
// tables it stages into its shared memory; every part has its own stream counter and hand-over queue.  The FPGA holds
// the whole NFA in one memory and scans it in one pass (Design/FPGA.v:773-795); here the batch is still read once per
// part, but by CTAs that run side by side: one kernel, one tail, streams spread over 148 / n_parts SMs per part instead
// of a pass per part in which every lane gets one or two streams.
template <int W, int RING_CAP>
__global__ void __launch_bounds__(LANE_THREADS, 1)
scan_lane_multi_kernel(const NfaDev *__restrict__ parts, const uint32_t n_parts, const BatchDev batch_in, const OutDev out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ NfaDev s_nfa;
    const uint32_t p = blockIdx.x % n_parts;
    {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(parts + p);
        uint32_t *dst = reinterpret_cast<uint32_t *>(&s_nfa);
        for (uint32_t i = threadIdx.x; i < sizeof(NfaDev) / 4; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    BatchDev batch = batch_in;
    batch.count_symbols = batch_in.count_symbols && p == 0;       // ragged batches: the stream lengths are summed once
    lane_body<W, RING_CAP>(s_nfa, batch, out, &out.g->part_next[p], &out.g->part_rescan[p], out.rescan + (size_t)p * batch.n_streams, smem);
}

cudaError_t launch_scan_lane(const NfaDev &nfa_in, const BatchDev &batch, const OutDev &out, int n_sms, cudaStream_t stream) {
    const size_t smem = lane_smem_bytes(nfa_in.h);
    NfaDev nfa = nfa_in;
    nfa.hot_rows = lane_hot_rows(nfa.h, &nfa.hot_bytes);
    unsigned long long want = (batch.n_streams + LANE_THREADS - 1) / LANE_THREADS;
    int grid = (int)(want < (unsigned long long)n_sms ? (want ? want : 1) : (unsigned long long)n_sms);
    const int cap = lane_ring_cap(nfa.h);

    const bool w1 = nfa.h.sticky_words == 1;
#define RFB_LAUNCH(W_, C_) scan_lane_kernel<W_, C_><<<grid, LANE_THREADS, smem, stream>>>(nfa, batch, out)
    if (cap == 64) { if (w1) RFB_LAUNCH(1, 64); else RFB_LAUNCH(2, 64); }
    else if (cap == 32) { if (w1) RFB_LAUNCH(1, 32); else RFB_LAUNCH(2, 32); }
    else if (cap == 16) { if (w1) RFB_LAUNCH(1, 16); else RFB_LAUNCH(2, 16); }
    else return cudaErrorInvalidValue;
#undef RFB_LAUNCH
    return cudaGetLastError();
}

// parts[0..n_parts) on the device (hot_rows / hot_bytes already set), all with `sticky_words` and ring capacity `cap`
cudaError_t launch_scan_lane_multi(const NfaDev *d_parts, uint32_t n_parts, uint32_t sticky_words, int cap, size_t smem,
                                   const BatchDev &batch, const OutDev &out, int n_sms, cudaStream_t stream) {
    if (n_parts == 0 || n_parts > MAX_MULTI_PARTS) return cudaErrorInvalidValue;
    // every part gets the same number of CTAs (a CTA cannot help another part: the tables in its shared memory are its part's)
    unsigned long long per_part = (batch.n_streams + LANE_THREADS - 1) / LANE_THREADS;
    const unsigned long long max_per_part = std::max<unsigned long long>(1, (unsigned long long)n_sms / n_parts);
    if (per_part > max_per_part) per_part = max_per_part;
    if (per_part == 0) per_part = 1;
    const int grid = (int)(per_part * n_parts);
    const bool w1 = sticky_words == 1;
#define RFB_LAUNCH(W_, C_) scan_lane_multi_kernel<W_, C_><<<grid, LANE_THREADS, smem, stream>>>(d_parts, n_parts, batch, out)
    if (cap == 64) { if (w1) RFB_LAUNCH(1, 64); else RFB_LAUNCH(2, 64); }
    else if (cap == 32) { if (w1) RFB_LAUNCH(1, 32); else RFB_LAUNCH(2, 32); }
    else if (cap == 16) { if (w1) RFB_LAUNCH(1, 16); else RFB_LAUNCH(2, 16); }
    else return cudaErrorInvalidValue;
#undef RFB_LAUNCH
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// warp kernel
// ------------------------------------------------------------------------------------------------
size_t warp_smem_bytes(uint32_t n_states, int warps) {
    const size_t nw = (n_states + 31) / 32;
    return (size_t)warps * (2 * nw * 4 + 2 * (size_t)WARP_LCAP * 4 + 16);
}

__global__ void __launch_bounds__(WARP_THREADS)
scan_warp_kernel(const NfaDev nfa, const BatchDev batch, const OutDev out, const int from_rescan, const int warps_per_cta) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if ((int)wid >= warps_per_cta) return;
    const uint32_t nw = (nfa.n_states + 31) / 32;
    const size_t per_warp = 2 * (size_t)nw * 4 + 2 * (size_t)WARP_LCAP * 4 + 16;
    uint8_t *my = smem + wid * per_warp;
    uint32_t *bits_cur = reinterpret_cast<uint32_t *>(my);
    uint32_t *bits_nxt = bits_cur + nw;
    uint32_t *list_cur = bits_nxt + nw;
    uint32_t *list_nxt = list_cur + WARP_LCAP;
    uint32_t *n_next = list_nxt + WARP_LCAP;
    const uint32_t *__restrict__ ep = nfa.eptr;
    const unsigned long long *__restrict__ er = nfa.erec;
    const uint32_t *__restrict__ em = nfa.emembs;
    const uint32_t *__restrict__ smap = nfa.state_map;

    for (uint32_t w = lane; w < 2 * nw; w += 32) bits_cur[w] = 0;
    __syncwarp();
    // hand-over queue: the batch's own, or (parts of a cut NFA scanned in one launch) the part's
    const unsigned int *q_n = out.q_rescan_n ? out.q_rescan_n : &out.g->n_rescan;
    unsigned int *q_item = out.q_next_item ? out.q_next_item : &out.g->next_item;
    if (from_rescan && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&out.g->n_rescan_total, *q_n);

    auto insert = [&](uint32_t t) {
        const uint32_t bit = 1u << (t & 31);
        const uint32_t old = atomicOr(&bits_nxt[t >> 5], bit);
        if (!(old & bit)) {
            const uint32_t p = atomicAdd(n_next, 1u);
            if (p < (uint32_t)WARP_LCAP) list_nxt[p] = t;
        }
    };

    for (;;) {
        // ---- fetch a stream ----
        unsigned int item = 0;
        if (lane == 0) item = atomicAdd(q_item, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        uint32_t sid, emit_from = 0;
        if (from_rescan) {
            if (item >= *q_n) break;
            const uint2 r = out.rescan[item];
            sid = r.x; emit_from = r.y;
        } else {
            if (item >= batch.n_streams) break;
            sid = item;
        }
        const uint32_t nsteps = batch.steps ? batch.steps[sid] : batch.n_steps;
        if (!from_rescan && batch.steps && batch.count_symbols && lane == 0) atomicAdd(&out.g->n_symbols, (unsigned long long)nsteps);
        const uint8_t *sp = stream_ptr(batch, sid);

        uint32_t ncur = 1;
        bool dense = false;
        if (batch.state_in) {   // resume from the set an earlier call left
            const unsigned int *stt = batch.state_in + (size_t)sid * (1u + batch.state_cap);
            const uint32_t ns = stt[0] <= batch.state_cap ? stt[0] : 0u;
            ncur = 0;
            for (uint32_t base = 0; base < ns; base += 32) {             // keep the members that belong to this (sub-)NFA
                const uint32_t i = base + lane;
                uint32_t s = i < ns ? stt[1 + i] : 0xFFFFFFFFu;
                if (s >= nfa.n_ref_states) s = 0xFFFFFFFFu;               // not a state: ignored
                else if (nfa.sub_of_ref) s = nfa.sub_of_ref[s];
                const uint32_t m = __ballot_sync(0xffffffffu, s != 0xFFFFFFFFu);
                if (s != 0xFFFFFFFFu) list_cur[ncur + __popc(m & ((1u << lane) - 1u))] = s;
                ncur += __popc(m);
            }
        } else if (lane == 0) list_cur[0] = 0;                            // Design/FPGA.v:146-147
        __syncwarp();
        for (uint32_t k = 0; k < nsteps; k++) {
            const uint32_t c = __ldg(sp + k);
            if (lane == 0) *n_next = 0;
            __syncwarp();
            const bool report = k >= emit_from;
            // expand S_k through the edge-grouped rows: one active state per lane, its few edges serially
            auto expand = [&](uint32_t s) {
                const uint32_t e0 = ep[s], e1 = ep[s + 1];
                if (e0 == e1) { if (report && s != nfa.no_report_sub) emit_match(out, sid + batch.stream_id_base, k + batch.pos_base, smap ? smap[s] : s); return; }   // FPGA.v:210-226
                for (uint32_t j = e0; j < e1; j++) {
                    const unsigned long long r = er[j];
                    const uint32_t lo = (uint32_t)r;
                    const bool hit = (lo & 0x10000u) ? ((em[(lo >> 17) * 8 + (c >> 5)] >> (c & 31)) & 1u) != 0
                                                     : (c == (lo & 0xFFu) || c == ((lo >> 8) & 0xFFu));
                    if (hit) insert((uint32_t)(r >> 32));
                }
            };
            if (!dense && ncur <= 8) {
                // few active states: the 32 lanes share each state's edges (coalesced record loads)
                for (uint32_t i = 0; i < ncur; i++) {
                    const uint32_t s = list_cur[i];
                    const uint32_t e0 = ep[s], e1 = ep[s + 1];
                    if (e0 == e1 && report && lane == 0 && s != nfa.no_report_sub) emit_match(out, sid + batch.stream_id_base, k + batch.pos_base, smap ? smap[s] : s);   // FPGA.v:210-226
                    for (uint32_t j = e0 + lane; j < e1; j += 32) {
                        const unsigned long long r = er[j];
                        const uint32_t lo = (uint32_t)r;
                        const bool hit = (lo & 0x10000u) ? ((em[(lo >> 17) * 8 + (c >> 5)] >> (c & 31)) & 1u) != 0
                                                         : (c == (lo & 0xFFu) || c == ((lo >> 8) & 0xFFu));
                        if (hit) insert((uint32_t)(r >> 32));
                    }
                }
            } else if (!dense) {
                for (uint32_t i = lane; i < ncur; i += 32) expand(list_cur[i]);
            } else {
                for (uint32_t wd = lane; wd < nw; wd += 32) {
                    uint32_t bits = bits_cur[wd];
                    while (bits) {
                        const uint32_t s = wd * 32 + (uint32_t)__ffs((int)bits) - 1u;
                        bits &= bits - 1;
                        expand(s);
                    }
                }
            }
            __syncwarp();
            // clear the consumed set, then current <= next (Design/FPGA.v:733-737)
            if (!dense) { for (uint32_t i = lane; i < ncur; i += 32) bits_cur[list_cur[i] >> 5] = 0; }
            else { for (uint32_t wd = lane; wd < nw; wd += 32) bits_cur[wd] = 0; }
            __syncwarp();
            ncur = *n_next;
            dense = ncur > (uint32_t)WARP_LCAP;
            uint32_t *tb = bits_cur; bits_cur = bits_nxt; bits_nxt = tb;
            uint32_t *tl = list_cur; list_cur = list_nxt; list_nxt = tl;
            __syncwarp();
        }
        if (batch.state_out) {   // S_{n_steps}: handed to the caller instead of being dropped
            unsigned int *dst = batch.state_out + (size_t)sid * (1u + batch.state_cap);
            const uint32_t n0 = batch.state_append ? dst[0] : 0u;          // an earlier part's overflow mark stays
            const bool fits = !dense && n0 != 0xFFFFFFFFu && n0 + ncur <= batch.state_cap;
            __syncwarp();
            if (lane == 0) dst[0] = fits ? n0 + ncur : 0xFFFFFFFFu;
            if (fits) for (uint32_t i = lane; i < ncur; i += 32) dst[1 + n0 + i] = smap ? smap[list_cur[i]] : list_cur[i];
        }
        // leave both bit vectors clean for the next stream (S_{n_steps} is never examined, TB:71-86)
        if (!dense) { for (uint32_t i = lane; i < ncur; i += 32) bits_cur[list_cur[i] >> 5] = 0; }
        else { for (uint32_t wd = lane; wd < nw; wd += 32) bits_cur[wd] = 0; }
        __syncwarp();
    }
}

static int warp_warps_for(uint32_t n_states) {
    int warps = WARP_THREADS / 32;
    while (warps > 1 && warp_smem_bytes(n_states, warps) > 200 * 1024) warps >>= 1;
    return warps;
}

cudaError_t launch_scan_warp(const NfaDev &nfa, const BatchDev &batch, const OutDev &out, bool from_rescan, int n_sms, cudaStream_t stream) {
    const int warps = warp_warps_for(nfa.n_states);
    const size_t smem = warp_smem_bytes(nfa.n_states, warps);
    if (smem > MAX_DYN_SMEM) return cudaErrorInvalidValue;
    unsigned long long want = from_rescan ? (unsigned long long)n_sms * 2 : (batch.n_streams + warps - 1) / warps;
    unsigned long long cap = (unsigned long long)n_sms * 8;
    int grid = (int)(want < cap ? (want ? want : 1) : cap);
    scan_warp_kernel<<<grid, WARP_THREADS, smem, stream>>>(nfa, batch, out, from_rescan ? 1 : 0, warps);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// testbench cycle model: "Total no. cycles" (testbench_BLK_Mem.sv:52,84) of the two-stream lock-step run
// ------------------------------------------------------------------------------------------------
// The FSM spends 1 cycle on every state index that is inactive in both streams (Design/FPGA.v:744-752) and
// cost(s) cycles on every state active in either (FPGA.v:158-743; closed form in DESIGN.md section 9):
//   step_cycles(k) = (size - |U_k|) + sum_{s in U_k} cost(s),  U_k = S_k(lo) | S_k(hi);  total = 1 + sum_k.
// One CTA walks the pair: four bit vectors in shared memory, threads own bitmap words.
__global__ void __launch_bounds__(256)
tb_cycles_kernel(const NfaDev nfa, const uint32_t *__restrict__ cost, const uint8_t *__restrict__ lo,
                 const uint8_t *__restrict__ hi, const uint32_t n_steps, unsigned long long *total) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t nw = (nfa.n_states + 31) / 32;
    uint32_t *c1 = reinterpret_cast<uint32_t *>(smem), *c2 = c1 + nw, *n1 = c2 + nw, *n2 = n1 + nw;
    unsigned long long &acc = *reinterpret_cast<unsigned long long *>(smem + (((size_t)4 * nw * 4 + 15) & ~(size_t)15));
    const uint32_t *__restrict__ rp = nfa.row_ptr;
    const uint32_t *__restrict__ tr = nfa.trans;
    for (uint32_t w = threadIdx.x; w < 4 * nw; w += blockDim.x) c1[w] = 0;
    if (threadIdx.x == 0) acc = 1;                       // the reset edge (testbench_BLK_Mem.sv:31-39)
    __syncthreads();
    if (threadIdx.x == 0) { c1[0] = 1; c2[0] = 1; }      // current[0] <= 1 for both streams (FPGA.v:146-147)
    __syncthreads();
    for (uint32_t k = 0; k < n_steps; k++) {
        const uint32_t s1 = lo[k], s2 = hi[k];
        unsigned long long mine = 0;
        for (uint32_t w = threadIdx.x; w < nw; w += blockDim.x) {
            const uint32_t a1 = c1[w], a2 = c2[w];
            uint32_t u = a1 | a2;
            const uint32_t valid = min(32u, nfa.n_states - w * 32);
            mine += valid - __popc(u);
            while (u) {
                const uint32_t b = (uint32_t)__ffs((int)u) - 1u;
                u &= u - 1;
                const uint32_t s = w * 32 + b;
                mine += cost[s];
                const bool in1 = (a1 >> b) & 1u, in2 = (a2 >> b) & 1u;
                for (uint32_t j = rp[s]; j < rp[s + 1]; j++) {
                    const uint32_t e = tr[j], sy = e >> 24, t = e & 0xFFFFFFu;
                    if (in1 && sy == s1) atomicOr(&n1[t >> 5], 1u << (t & 31));
                    if (in2 && sy == s2) atomicOr(&n2[t >> 5], 1u << (t & 31));
                }
            }
        }
        if (mine) atomicAdd(&acc, mine);
        __syncthreads();
        for (uint32_t w = threadIdx.x; w < nw; w += blockDim.x) { c1[w] = n1[w]; c2[w] = n2[w]; n1[w] = 0; n2[w] = 0; }
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = acc;
}

cudaError_t launch_tb_cycles(const NfaDev &nfa, const uint32_t *cost, const uint8_t *lo, const uint8_t *hi, uint32_t n_steps,
                             unsigned long long *total, cudaStream_t stream) {
    const size_t smem = (((size_t)((nfa.n_states + 31) / 32) * 16 + 15) & ~(size_t)15) + 16;
    if (smem > MAX_DYN_SMEM) return cudaErrorInvalidValue;
    tb_cycles_kernel<<<1, 256, smem, stream>>>(nfa, cost, lo, hi, n_steps, total);
    return cudaGetLastError();
}

cudaError_t configure_kernels() {
    cudaError_t e;
#define RFB_ATTR(W_, C_) if ((e = cudaFuncSetAttribute(scan_lane_kernel<W_, C_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MAX_DYN_SMEM)) != cudaSuccess) return e
    RFB_ATTR(1, 16); RFB_ATTR(2, 16); RFB_ATTR(1, 32); RFB_ATTR(2, 32); RFB_ATTR(1, 64); RFB_ATTR(2, 64);
#undef RFB_ATTR
    // (the multi-part kernel keeps its part's descriptor in 256 bytes of static shared memory)
#define RFB_ATTR(W_, C_) if ((e = cudaFuncSetAttribute(scan_lane_multi_kernel<W_, C_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MAX_DYN_SMEM - 1024)) != cudaSuccess) return e
    RFB_ATTR(1, 16); RFB_ATTR(2, 16); RFB_ATTR(1, 32); RFB_ATTR(2, 32); RFB_ATTR(1, 64); RFB_ATTR(2, 64);
#undef RFB_ATTR
    if ((e = cudaFuncSetAttribute(tb_cycles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MAX_DYN_SMEM)) != cudaSuccess) return e;
    return cudaFuncSetAttribute(scan_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MAX_DYN_SMEM);
}

}  // namespace rfb
