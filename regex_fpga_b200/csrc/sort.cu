// sort.cu -- canonical ordering of match records on the device.
//
// The FSM reports matches in time order per stream, ascending state index within a step (Design/FPGA.v:725-728);
// the kernels append records in arrival order.  RFB_SCAN_SORT_RECORDS restores the canonical (stream, pos, state)
// order with a stable LSD radix sort over the 12-byte records, 8 bits per pass, skipping the key bytes that are
// zero for the whole batch (the host knows the largest stream id, position and state id).
#include "device.h"

namespace rfb {

constexpr int SORT_TILE = 4096;      // records per block (one warp ranks them in order: stable)

__device__ __forceinline__ uint32_t sort_digit(const rfb_match &m, int byte) {
    // byte 0..3: state, 4..7: pos, 8..11: stream (least significant key byte first)
    const uint32_t w = byte < 4 ? m.state : (byte < 8 ? m.pos : m.stream);
    return (w >> ((byte & 3) * 8)) & 0xFFu;
}

__global__ void __launch_bounds__(256) sort_hist_kernel(const rfb_match *in, unsigned long long n, int byte, uint32_t *hist, uint32_t n_blocks) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const unsigned long long lo = (unsigned long long)blockIdx.x * SORT_TILE;
    const unsigned long long hi = lo + SORT_TILE < n ? lo + SORT_TILE : n;
    for (unsigned long long i = lo + threadIdx.x; i < hi; i += blockDim.x) atomicAdd(&h[sort_digit(in[i], byte)], 1u);
    __syncthreads();
    hist[(size_t)threadIdx.x * n_blocks + blockIdx.x] = h[threadIdx.x];     // digit-major: scan order = output order
}

// exclusive prefix sum over hist[256 * n_blocks] (one CTA; n_blocks is a few hundred).  256 threads at <= 32
// registers: small enough to run on an SM that a lane-kernel CTA occupies (pipelined host path, api.cu)
constexpr int SCAN_THREADS = 256;
__global__ void __launch_bounds__(SCAN_THREADS) sort_scan_kernel(uint32_t *hist, uint32_t total) {
    __shared__ uint32_t part[SCAN_THREADS];
    const uint32_t per = (total + SCAN_THREADS - 1) / SCAN_THREADS;
    const uint32_t lo = threadIdx.x * per, hi = lo + per < total ? lo + per : total;
    uint32_t s = 0;
    for (uint32_t i = lo; i < hi; i++) s += hist[i];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) { uint32_t run = 0; for (int i = 0; i < SCAN_THREADS; i++) { const uint32_t v = part[i]; part[i] = run; run += v; } }
    __syncthreads();
    uint32_t run = part[threadIdx.x];
    for (uint32_t i = lo; i < hi; i++) { const uint32_t v = hist[i]; hist[i] = run; run += v; }
}

__global__ void __launch_bounds__(32) sort_scatter_kernel(const rfb_match *in, rfb_match *outp, unsigned long long n, int byte,
                                                         const uint32_t *hist, uint32_t n_blocks) {
    __shared__ uint32_t off[256];
    const uint32_t lane = threadIdx.x;
    for (uint32_t d = lane; d < 256; d += 32) off[d] = hist[(size_t)d * n_blocks + blockIdx.x];
    __syncwarp();
    const unsigned long long lo = (unsigned long long)blockIdx.x * SORT_TILE;
    const unsigned long long hi = lo + SORT_TILE < n ? lo + SORT_TILE : n;
    for (unsigned long long base = lo; base < hi; base += 32) {
        const unsigned long long i = base + lane;
        const bool valid = i < hi;
        rfb_match m;
        m.stream = 0; m.pos = 0; m.state = 0;
        if (valid) m = in[i];
        const uint32_t d = valid ? sort_digit(m, byte) : 0xFFFFFFFFu;
        const uint32_t peers = __match_any_sync(0xffffffffu, d);             // lanes with the same digit, in record order
        const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
        if (valid) outp[off[d] + rank] = m;
        __syncwarp();
        if (valid && rank == 0) off[d] += __popc(peers);
        __syncwarp();
    }
}

// Sorts records[0..n) in place (tmp: same capacity; hist: 256 * ceil(n / SORT_TILE) words).  key_bytes: bit i set
// when key byte i (see sort_digit) can be non-zero.
cudaError_t launch_sort_records(rfb_match *records, rfb_match *tmp, unsigned long long n, uint32_t key_bytes, uint32_t *hist,
                                cudaStream_t stream) {
    if (n < 2) return cudaSuccess;
    const uint32_t n_blocks = (uint32_t)((n + SORT_TILE - 1) / SORT_TILE);
    rfb_match *src = records, *dst = tmp;
    for (int byte = 0; byte < 12; byte++) {
        if (!((key_bytes >> byte) & 1u)) continue;
        sort_hist_kernel<<<n_blocks, 256, 0, stream>>>(src, n, byte, hist, n_blocks);
        sort_scan_kernel<<<1, SCAN_THREADS, 0, stream>>>(hist, 256u * n_blocks);
        sort_scatter_kernel<<<n_blocks, 32, 0, stream>>>(src, dst, n, byte, hist, n_blocks);
        rfb_match *t = src; src = dst; dst = t;
    }
    if (src != records) {
        cudaError_t e = cudaMemcpyAsync(records, src, n * sizeof(rfb_match), cudaMemcpyDeviceToDevice, stream);
        if (e != cudaSuccess) return e;
    }
    return cudaGetLastError();
}

size_t sort_hist_words(unsigned long long n) { return 256 * (size_t)((n + SORT_TILE - 1) / SORT_TILE) + 256; }

}  // namespace rfb
