// scan.cu -- the two scan kernels (sm_100a).
//
//  scan_lane_kernel : the hot path.  One THREAD per stream, 1024 threads per SM, the whole execution
//      image (image.cpp) staged into shared memory with bulk async copies (cp.async.bulk -> UBLKCP),
//      per-stream state = sticky bit mask in registers + a short ring of transient state ids in
//      shared memory (column-major, bank-conflict-free).  Every lane walks its own stream at its own
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

__device__ __forceinline__ void stage_image(uint8_t *dst, const uint8_t *src, uint32_t bytes, uint64_t *bar) {
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
        const uint32_t CH = 32768;
        for (uint32_t o = 0; o < bytes; o += CH) {
            uint32_t n = bytes - o < CH ? bytes - o : CH;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(dst + o)), "l"(src + o), "r"(n), "r"(smem_u32(bar)) : "memory");
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
size_t lane_smem_bytes(const ImageHeader &h) {
    return (size_t)h.blob_bytes + (size_t)LANE_CAP * LANE_THREADS * sizeof(uint16_t) + 16;
}

template <int W>
__global__ void __launch_bounds__(LANE_THREADS, 1)
scan_lane_kernel(const NfaDev nfa, const BatchDev batch, const OutDev out) {
    extern __shared__ __align__(128) uint8_t smem[];
    const ImageHeader &h = nfa.h;
    uint16_t *lists = reinterpret_cast<uint16_t *>(smem + h.blob_bytes);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + h.blob_bytes + (size_t)LANE_CAP * LANE_THREADS * sizeof(uint16_t));
    stage_image(smem, nfa.blob, h.blob_bytes, bar);

    const uint32_t *tab = reinterpret_cast<const uint32_t *>(smem + h.off_tab);
    const uint16_t *inj = reinterpret_cast<const uint16_t *>(smem + h.off_inj);
    const uint8_t *mask = smem + h.off_mask;
    const uint32_t *memb = reinterpret_cast<const uint32_t *>(smem + h.off_memb);
    const uint16_t *tlist = reinterpret_cast<const uint16_t *>(smem + h.off_tlist);
    uint16_t *list = lists + threadIdx.x;  // entry i of this lane: list[i * LANE_THREADS]
    const uint32_t gbase = h.gbase, nsb = h.nsb, hmul = h.hash_mul, hsh = h.hash_shift;
    const uint32_t nbm = (1u << h.bucket_bits) - 1u;
    constexpr uint32_t MSTRIDE = 32u * W;
    constexpr uint32_t FULL = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31u;

    // A warp takes 32 consecutive streams at a time and steps them in lock-step, one symbol per
    // iteration of the k loop, so that the per-symbol bookkeeping runs with all 32 lanes converged.
    for (;;) {
        unsigned int base = 0;
        if (lane == 0) base = atomicAdd(&out.g->next_stream, 32u);
        base = __shfl_sync(FULL, base, 0);
        if (base >= batch.n_streams) break;
        const uint32_t sid = base + lane;
        const bool valid = sid < batch.n_streams;
        const uint32_t nsteps = valid ? (batch.steps ? batch.steps[sid] : batch.n_steps) : 0u;
        const uint32_t maxsteps = __reduce_max_sync(FULL, nsteps);

        // ---- per-lane stream state ----
        uint64_t P[W], Pn[W];
#pragma unroll
        for (int w = 0; w < W; w++) { P[w] = 0; Pn[w] = 0; }
        uint32_t head = 0, rd = 0, ncur = 0, nnew = 0, flo = 0, fhi = 0;
        bool ovf = false;
        uint32_t ovf_at = 0;
        if (nsteps) {
            if (h.start_id < nsb) {                                                   // Design/FPGA.v:146-147
                if (W == 1 || h.start_id < 64) P[0] |= 1ull << (h.start_id & 63); else P[W - 1] |= 1ull << (h.start_id & 63);
            } else { list[0] = (uint16_t)h.start_id; ncur = 1; }
        }
        // input: aligned 16-byte chunks kept in registers, the first one shifted to the stream's first byte
        uint64_t blo = 0, bhi = 0, plo = 0, phi = 0;
        uint32_t bufn = 16;
        const uint8_t *nextp = nullptr, *endp = nullptr;
        if (nsteps) {
            const uint8_t *sp = stream_ptr(batch, sid);
            endp = sp + nsteps;
            const uint8_t *b16 = reinterpret_cast<const uint8_t *>(reinterpret_cast<uintptr_t>(sp) & ~(uintptr_t)15);
            const uint32_t off = (uint32_t)(sp - b16);
            uint4 v = __ldg(reinterpret_cast<const uint4 *>(b16));
            blo = (uint64_t)v.x | ((uint64_t)v.y << 32);
            bhi = (uint64_t)v.z | ((uint64_t)v.w << 32);
            const uint32_t sh = off * 8;
            if (sh >= 64) { blo = bhi >> (sh - 64); bhi = 0; }
            else if (sh) { blo = (blo >> sh) | (bhi << (64 - sh)); bhi >>= sh; }
            bufn = 16 - off;
            nextp = b16 + 16;
            if (nextp < endp) {
                v = __ldg(reinterpret_cast<const uint4 *>(nextp));
                plo = (uint64_t)v.x | ((uint64_t)v.y << 32);
                phi = (uint64_t)v.z | ((uint64_t)v.w << 32);
            }
            nextp += 16;
        }

        auto push = [&](uint32_t t) {
            if (t < nsb) {
                const uint64_t bit = 1ull << (t & 63);
                if (W == 1 || t < 64) Pn[0] |= bit; else Pn[W - 1] |= bit;
                return;
            }
            const uint32_t bit = 1u << (t & 31);
            const bool hi = (t & 32) != 0;
            const uint32_t f = hi ? fhi : flo;
            bool dup = false;
            if (f & bit) {  // possible duplicate: exact check against this step's new entries
                for (uint32_t j = 0; j < nnew; j++)
                    if (list[((head + ncur + j) & (LANE_CAP - 1)) * LANE_THREADS] == t) { dup = true; break; }
            }
            if (!dup) {
                if (ncur - rd + nnew >= (uint32_t)LANE_CAP) { ovf = true; return; }
                list[((head + ncur + nnew) & (LANE_CAP - 1)) * LANE_THREADS] = (uint16_t)t;
                nnew++;
                if (hi) fhi |= bit; else flo |= bit;
            }
        };

        for (uint32_t k = 0; k < maxsteps; k++) {
            const bool act = k < nsteps && !ovf;
            // ---- next symbol ----
            if (bufn == 0) {
                blo = plo; bhi = phi; bufn = 16;
                if (nextp < endp) {
                    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(nextp));
                    plo = (uint64_t)v.x | ((uint64_t)v.y << 32);
                    phi = (uint64_t)v.z | ((uint64_t)v.w << 32);
                }
                nextp += 16;
            }
            const uint32_t c = (uint32_t)blo & 0xFFu;
            blo = (blo >> 8) | (bhi << 56);
            bhi >>= 8;
            bufn--;
            const uint32_t hc = ((c * hmul) >> hsh) & nbm;

            // ---- transient states of S_k: one table lookup per lane per iteration ----
            bool walking = false;
            uint32_t idx = 0;
            bool work = act && rd < ncur;
            while (__any_sync(FULL, work)) {
                if (work) {
                    if (!walking) {
                        const uint32_t u = list[((head + rd) & (LANE_CAP - 1)) * LANE_THREADS];
                        rd++;
                        idx = u + (u >= gbase ? hc : 0u);
                    }
                    const uint32_t e = tab[idx];
                    const uint32_t a = e & 0xFFu, b = (e >> 8) & 0xFFu, t = (e >> 16) & 0x7FFFu;
                    bool hit = false, redirect = false;
                    if (a <= b) hit = (c == a) | (c == b);
                    else if (a == 0xFFu) {
                        if (b == CODE_ACCEPT) emit_match(out, sid + batch.stream_id_base, k, nfa.orig_of_id[idx]);
                        else if (b == CODE_INDIRECT) redirect = true;
                    } else {
                        const uint32_t n = (0xFEu - a) * 253u + b;
                        hit = (memb[n * 8 + (c >> 5)] >> (c & 31)) & 1u;
                    }
                    if (hit && !ovf) push(t);
                    if (redirect) { idx = t; walking = true; }
                    else if (e & TAB_MORE) { idx++; walking = true; }
                    else walking = false;
                    work = walking || rd < ncur;
                }
            }

            // ---- sticky states: P' = (P & K[c]) | entered ; injections from P & M[c] ----
            if (act) {
                const uint8_t *mrow = mask + c * MSTRIDE;
                bool attn = false;
#pragma unroll
                for (int w = 0; w < W; w++) attn |= (P[w] & reinterpret_cast<const uint64_t *>(mrow)[w]) != 0;
                if (attn) {
#pragma unroll
                    for (int w = 0; w < W; w++) {
                        const uint64_t K = reinterpret_cast<const uint64_t *>(mrow + 16)[w];
                        const uint64_t M = reinterpret_cast<const uint64_t *>(mrow + 16)[W + w];
                        uint64_t im = P[w] & M;
                        P[w] &= K;
                        while (im) {
                            const uint32_t bpos = (uint32_t)__ffsll((long long)im) - 1u;
                            im &= im - 1;
                            const uint32_t x = inj[(w * 64 + bpos) * 256 + c];
                            if (ovf) continue;
                            if (x < 0x8000u) push(x);
                            else if (x != 0xFFFFu) {
                                uint32_t q = x & 0x7FFFu, tl;
                                do { tl = tlist[q++]; push(tl & 0x7FFFu); } while ((tl & 0x8000u) && !ovf);
                            }
                        }
                    }
                }
#pragma unroll
                for (int w = 0; w < W; w++) { P[w] |= Pn[w]; Pn[w] = 0; }
                // current <= next (Design/FPGA.v:733-737)
                head += ncur; rd = 0; ncur = nnew; nnew = 0; flo = 0; fhi = 0;
                if (ovf) ovf_at = k + 1;   // S_k was fully examined; the general kernel reports from step k+1 on
            }
        }
        if (ovf && ovf_at < nsteps) {
            const unsigned int slot = atomicAdd(&out.g->n_rescan, 1u);
            out.rescan[slot] = make_uint2(sid, ovf_at);
        }
        if (batch.steps) {
            unsigned long long tot = nsteps;
#pragma unroll
            for (int o = 16; o; o >>= 1) tot += __shfl_xor_sync(FULL, tot, o);
            if (lane == 0) atomicAdd(&out.g->n_symbols, tot);
        }
    }
}

cudaError_t launch_scan_lane(const NfaDev &nfa, const BatchDev &batch, const OutDev &out, int n_sms, cudaStream_t stream) {
    const size_t smem = lane_smem_bytes(nfa.h);
    unsigned long long want = (batch.n_streams + LANE_THREADS - 1) / LANE_THREADS;
    int grid = (int)(want < (unsigned long long)n_sms ? (want ? want : 1) : (unsigned long long)n_sms);
    if (nfa.h.sticky_words == 1) scan_lane_kernel<1><<<grid, LANE_THREADS, smem, stream>>>(nfa, batch, out);
    else scan_lane_kernel<2><<<grid, LANE_THREADS, smem, stream>>>(nfa, batch, out);
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
    const uint32_t *__restrict__ rp = nfa.row_ptr;
    const uint32_t *__restrict__ tr = nfa.trans;

    for (uint32_t w = lane; w < 2 * nw; w += 32) bits_cur[w] = 0;
    __syncwarp();

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
        if (lane == 0) item = atomicAdd(&out.g->next_item, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        uint32_t sid, emit_from = 0;
        if (from_rescan) {
            if (item >= out.g->n_rescan) break;
            const uint2 r = out.rescan[item];
            sid = r.x; emit_from = r.y;
        } else {
            if (item >= batch.n_streams) break;
            sid = item;
        }
        const uint32_t nsteps = batch.steps ? batch.steps[sid] : batch.n_steps;
        if (!from_rescan && batch.steps && lane == 0) atomicAdd(&out.g->n_symbols, (unsigned long long)nsteps);
        const uint8_t *sp = stream_ptr(batch, sid);

        uint32_t ncur = 1;
        bool dense = false;
        if (lane == 0) list_cur[0] = 0;                                   // Design/FPGA.v:146-147
        __syncwarp();
        for (uint32_t k = 0; k < nsteps; k++) {
            const uint32_t c = __ldg(sp + k);
            if (lane == 0) *n_next = 0;
            __syncwarp();
            const bool report = k >= emit_from;
            if (!dense) {
                for (uint32_t base = 0; base < ncur; base += 32) {
                    const uint32_t i = base + lane;
                    const bool valid = i < ncur;
                    uint32_t s = 0, r0 = 0, r1 = 0;
                    if (valid) { s = list_cur[i]; r0 = rp[s]; r1 = rp[s + 1]; }
                    const uint32_t len = r1 - r0;
                    if (valid && len == 0 && report) emit_match(out, sid + batch.stream_id_base, k, s);   // FPGA.v:210-226
                    const bool is_long = valid && len > 8;
                    if (valid && !is_long)
                        for (uint32_t j = r0; j < r1; j++) { const uint32_t w = tr[j]; if ((w >> 24) == c) insert(w & 0xFFFFFFu); }
                    uint32_t big = __ballot_sync(0xffffffffu, is_long);
                    while (big) {   // one long row at a time, 32 transitions per pass (FPGA.v:227-407 does 4)
                        const int l = __ffs((int)big) - 1;
                        big &= big - 1;
                        const uint32_t R0 = __shfl_sync(0xffffffffu, r0, l), R1 = __shfl_sync(0xffffffffu, r1, l);
                        for (uint32_t j = R0 + lane; j < R1; j += 32) { const uint32_t w = tr[j]; if ((w >> 24) == c) insert(w & 0xFFFFFFu); }
                    }
                }
            } else {
                for (uint32_t wd = lane; wd < nw; wd += 32) {
                    uint32_t bits = bits_cur[wd];
                    while (bits) {
                        const uint32_t s = wd * 32 + (uint32_t)__ffs((int)bits) - 1u;
                        bits &= bits - 1;
                        const uint32_t r0 = rp[s], r1 = rp[s + 1];
                        if (r0 == r1 && report) emit_match(out, sid + batch.stream_id_base, k, s);
                        for (uint32_t j = r0; j < r1; j++) { const uint32_t w = tr[j]; if ((w >> 24) == c) insert(w & 0xFFFFFFu); }
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

cudaError_t configure_kernels() {
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(scan_lane_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MAX_DYN_SMEM)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(scan_lane_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MAX_DYN_SMEM)) != cudaSuccess) return e;
    return cudaFuncSetAttribute(scan_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MAX_DYN_SMEM);
}

}  // namespace rfb
