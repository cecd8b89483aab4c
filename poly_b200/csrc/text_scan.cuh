// text_scan.cuh -- building blocks shared by the text ingest kernels (fastq_ingest.cu,
// fasta_ingest.cu): newline positions of a byte buffer and an ordered device-wide exclusive scan
// over an arbitrary associative operator (sums, and the composition of parser-state functions).
#pragma once
#include <algorithm>

#include "common.cuh"

namespace pg {
namespace text {

constexpr int BLOCK_BYTES = 4096;
constexpr int THREADS = 256;  // 16 bytes per thread

// ---- ordered exclusive scan --------------------------------------------------------------------
// Op: typedef In, T; static T identity(); static T lift(In); static T combine(T earlier, T later).
// `combine` need not commute.  out has n + 1 entries, out[n] = the total.
constexpr int SCAN_TILE = 4096;  // 256 threads x 16 consecutive items

template <typename Op>
__device__ __forceinline__ typename Op::T block_exclusive(typename Op::T mine, typename Op::T *s_warp,
                                                          typename Op::T *block_total) {
    using T = typename Op::T;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    T incl = mine;
    for (int d = 1; d < 32; d <<= 1) {
        const T y = __shfl_up_sync(0xffffffffu, incl, d);
        if ((int)lane >= d) incl = Op::combine(y, incl);
    }
    T excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = Op::identity();
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    T before = Op::identity(), all = Op::identity();
    for (uint32_t w = 0; w < blockDim.x / 32; ++w) {
        if (w == warp) before = all;
        all = Op::combine(all, s_warp[w]);
    }
    if (block_total) *block_total = all;
    return Op::combine(before, excl);
}

template <typename Op>
__global__ void __launch_bounds__(256) tile_reduce_kernel(const typename Op::In *__restrict__ in, uint64_t n,
                                                          typename Op::T *__restrict__ partial) {
    using T = typename Op::T;
    __shared__ T s_warp[8];
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + threadIdx.x * 16;
    T acc = Op::identity();
    for (int j = 0; j < 16; ++j)
        if (base + j < n) acc = Op::combine(acc, Op::lift(in[base + j]));
    T total;
    (void)block_exclusive<Op>(acc, s_warp, &total);
    if (threadIdx.x == 0) partial[blockIdx.x] = total;
}

// in-place exclusive scan of `data` (n entries) by one CTA; *total = combination of everything
template <typename Op>
__global__ void __launch_bounds__(1024) single_scan_kernel(typename Op::T *__restrict__ data, uint64_t n,
                                                           typename Op::T *__restrict__ total) {
    using T = typename Op::T;
    __shared__ T s_part[1024];
    const uint32_t tid = threadIdx.x;
    const uint64_t per = (n + 1023) / 1024, lo = min((uint64_t)tid * per, n), hi = min(lo + per, n);
    T acc = Op::identity();
    for (uint64_t i = lo; i < hi; ++i) acc = Op::combine(acc, data[i]);
    s_part[tid] = acc;
    __syncthreads();
    if (tid == 0) {
        T run = Op::identity();
        for (int i = 0; i < 1024; ++i) { const T t = s_part[i]; s_part[i] = run; run = Op::combine(run, t); }
        *total = run;
    }
    __syncthreads();
    T run = s_part[tid];
    for (uint64_t i = lo; i < hi; ++i) { const T t = data[i]; data[i] = run; run = Op::combine(run, t); }
}

template <typename Op>
__global__ void __launch_bounds__(256) tile_scan_kernel(const typename Op::In *__restrict__ in, uint64_t n,
                                                        const typename Op::T *__restrict__ partial_ex,
                                                        typename Op::T *__restrict__ out) {
    using T = typename Op::T;
    __shared__ T s_warp[8];
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + threadIdx.x * 16;
    T v[16];
    T acc = Op::identity();
    for (int j = 0; j < 16; ++j) {
        v[j] = base + j < n ? Op::lift(in[base + j]) : Op::identity();
        acc = Op::combine(acc, v[j]);
    }
    T run = Op::combine(partial_ex[blockIdx.x], block_exclusive<Op>(acc, s_warp, nullptr));
    for (int j = 0; j < 16; ++j) {
        if (base + j < n) out[base + j] = run;
        run = Op::combine(run, v[j]);
    }
}

template <typename Op>
int device_scan(const typename Op::In *d_in, uint64_t n, typename Op::T *d_out, cudaStream_t st) {
    using T = typename Op::T;
    const uint64_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    StreamScratch tmp(st);
    T *d_partial = nullptr;
    PG_CUDA(tmp.alloc(&d_partial, ntiles));
    if (ntiles) {
        tile_reduce_kernel<Op><<<(unsigned)ntiles, 256, 0, st>>>(d_in, n, d_partial);
        note_launch("tile_reduce_kernel");
    }
    single_scan_kernel<Op><<<1, 1024, 0, st>>>(d_partial, ntiles, d_out + n);
    note_launch("single_scan_kernel");
    if (ntiles) {
        tile_scan_kernel<Op><<<(unsigned)ntiles, 256, 0, st>>>(d_in, n, d_partial, d_out);
        note_launch("tile_scan_kernel");
    }
    return PG_OK;
}

template <typename TIn>
struct SumOp {
    using In = TIn;
    using T = unsigned long long;
    __device__ static T identity() { return 0ull; }
    __device__ static T lift(In x) { return (T)x; }
    __device__ static T combine(T a, T b) { return a + b; }
};

// ---- newline positions ---------------------------------------------------------------------------
static __global__ void __launch_bounds__(THREADS)
count_newlines_kernel(const uint8_t *__restrict__ text, uint64_t n, uint32_t *__restrict__ block_count) {
    const uint64_t base = (uint64_t)blockIdx.x * BLOCK_BYTES + threadIdx.x * 16;
    uint32_t c = 0;
    for (int j = 0; j < 16; ++j) c += (base + j < n && __ldg(text + base + j) == '\n');
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    __shared__ uint32_t s[THREADS / 32];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < THREADS / 32; ++w) t += s[w];
        block_count[blockIdx.x] = t;
    }
}

static __global__ void __launch_bounds__(THREADS)
write_newlines_kernel(const uint8_t *__restrict__ text, uint64_t n, const unsigned long long *__restrict__ block_start,
                      uint64_t *__restrict__ nl) {
    const uint64_t base = (uint64_t)blockIdx.x * BLOCK_BYTES + threadIdx.x * 16;
    uint32_t mask = 0;
    for (int j = 0; j < 16; ++j) mask |= (uint32_t)(base + j < n && __ldg(text + base + j) == '\n') << j;
    const uint32_t c = __popc(mask);
    uint32_t incl = c;
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
        if ((int)(threadIdx.x & 31) >= d) incl += y;
    }
    __shared__ uint32_t s[THREADS / 32];
    if ((threadIdx.x & 31) == 31) s[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint32_t wbase = 0;
    for (uint32_t w = 0; w < (threadIdx.x >> 5); ++w) wbase += s[w];
    uint64_t pos = block_start[blockIdx.x] + wbase + incl - c;
    for (int j = 0; j < 16; ++j)
        if (mask & (1u << j)) nl[pos++] = base + j;
}

// Positions of every '\n' of d_text[0, nbytes) in ascending order.  *d_nl is allocated on `st`
// (cudaFreeAsync it); *n_lines = their number; *last_plus1 = offset just past the last newline.
static inline int newline_positions(const uint8_t *d_text, uint64_t nbytes, uint64_t **d_nl, uint64_t *n_lines,
                                    uint64_t *last_plus1, cudaStream_t st) {
    const uint64_t nblocks = (nbytes + BLOCK_BYTES - 1) / BLOCK_BYTES;
    uint32_t *d_bcount = nullptr;
    unsigned long long *d_bstart = nullptr;
    *d_nl = nullptr; *n_lines = 0; *last_plus1 = 0;
    if (!nblocks) return PG_OK;
    StreamScratch tmp(st);
    PG_CUDA(tmp.alloc(&d_bcount, nblocks));
    PG_CUDA(tmp.alloc(&d_bstart, nblocks + 1));
    count_newlines_kernel<<<(unsigned)nblocks, THREADS, 0, st>>>(d_text, nbytes, d_bcount);
    note_launch("count_newlines_kernel");
    int rc = device_scan<SumOp<uint32_t>>(d_bcount, nblocks, d_bstart, st);
    if (rc != PG_OK) return rc;
    unsigned long long nlines = 0;
    PG_CUDA(cudaMemcpyAsync(&nlines, d_bstart + nblocks, 8, cudaMemcpyDeviceToHost, st));
    PG_CUDA(cudaStreamSynchronize(st));
    PG_CUDA(cudaMallocAsync(d_nl, std::max<uint64_t>(nlines, 1) * 8, st));
    write_newlines_kernel<<<(unsigned)nblocks, THREADS, 0, st>>>(d_text, nbytes, d_bstart, *d_nl);
    note_launch("write_newlines_kernel");
    if (nlines) {
        uint64_t last = 0;
        PG_CUDA(cudaMemcpyAsync(&last, *d_nl + nlines - 1, 8, cudaMemcpyDeviceToHost, st));
        PG_CUDA(cudaStreamSynchronize(st));
        *last_plus1 = last + 1;
    }
    *n_lines = nlines;
    return PG_OK;
}

}  // namespace text
}  // namespace pg
