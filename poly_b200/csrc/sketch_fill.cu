// sketch_fill.cu -- K1: mash.Sketch in the fill regime (L-k < s): every k-mer hash is
// written positionally (/root/reference/search/mash/mash.go:73-84).
//
// Fast path (fixed-length reads, k in the instantiated set, compact output):
//   * one CTA = one tile of R reads; the R*L input bytes are contiguous in HBM and are
//     brought in by ONE 1-D TMA bulk copy (cp.async.bulk + mbarrier; SASS UBLKCP);
//   * one THREAD per read walks the read 4 bytes per step.  The murmur3 block pre-mix
//     K(p) = rotl(w(p)*c1,15)*c2 of the 4 bytes at position p is computed ONCE and kept
//     in a register ring shared by the k/4 k-mers that consume it, so the cost per
//     k-mer is the body chain + fmix, not k/4 pre-mixes.  The kernel is bound by the
//     ALU pipe (SHF/LOP3 issue at one warp instruction per 2 cycles per SM sub-partition,
//     tools/pipe_bench.cu), i.e. by ALU instruction COUNT, so everything that is not
//     murmur3 arithmetic is kept off that pipe: for k % 4 == 1 the tail pre-mix comes from
//     a 256-entry shared-memory table (2 LSU instructions, table address formed by IMAD on
//     the FMA pipe), the main loop has no bounds tests or predicated stores, and raw words
//     are fetched one step ahead (DESIGN.md "K1", profiles/);
//   * hashes are staged in shared memory ([read][pos], stride L-k words: bank-conflict
//     free when L-k is odd) and leave as ONE bulk store per tile (the compact output
//     [n][L-k] of a tile is contiguous in HBM).
// All global traffic is therefore full-line TMA traffic; the SM only touches smem.
//
// Generic path (any k >= 0, ragged reads, any row stride): one warp per read, every
// k-mer hashed from scratch.  Slower, same results.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "murmur3.cuh"
#include "tma.cuh"

namespace pg {

// ---- K1 fast path ----------------------------------------------------------------
// smem: [0,16) mbarrier | in tile (R*L bytes + 16 B over-read pad, 16-B rounded) | out tile
// (+ 1 KB static table when k % 4 == 1)
__host__ __device__ inline uint32_t k1_in_bytes(uint32_t R, uint32_t L) {
    return (R * L + 16u + 15u) & ~15u;
}

__device__ __forceinline__ uint32_t lds_u32(uint32_t saddr) {
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}

// kmix(b) for a single tail byte b, built at compile time
struct KmixTable {
    uint32_t v[256];
};
constexpr uint32_t c_rotl(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
constexpr KmixTable make_kmix_table() {
    KmixTable t{};
    for (uint32_t b = 0; b < 256; ++b) t.v[b] = c_rotl(b * MM3_C1, 15) * MM3_C2;
    return t;
}
__device__ __align__(16) const KmixTable g_kmix_byte = make_kmix_table();

// One word step = 4 k-mers (ring slot U).  Uses the locals of the enclosing kernel.
#define PG_K1_STEP(U, CHECKED)                                                                 \
    {                                                                                          \
        uint32_t w[4];                                                                         \
        w[0] = w_cur;                                                                          \
        w[1] = __funnelshift_r(w_cur, w_nxt, 8);                                               \
        w[2] = __funnelshift_r(w_cur, w_nxt, 16);                                              \
        w[3] = __funnelshift_r(w_cur, w_nxt, 24);                                              \
        uint32_t h[4];                                                                         \
        _Pragma("unroll") for (int r = 0; r < 4; ++r) {                                        \
            uint32_t x = mm3_round0(ring[r][U]);                                               \
            _Pragma("unroll") for (int j = 1; j < NB; ++j) x = mm3_round(x, ring[r][((U) + j) % NB]); \
            if (LUT) x ^= lds_u32(sb[i + r] * lut_stride + lut_base);                          \
            else if (TAIL) x ^= mm3_kmix(w[r] & TAILMASK);                                     \
            x ^= (uint32_t)K;                                                                  \
            h[r] = mm3_fmix(x);                                                                \
            ring[r][U] = mm3_kmix(w[r]);                                                       \
        }                                                                                      \
        my_out[i] = h[0];                                                                      \
        if (!(CHECKED) || i + 1 < nk) my_out[i + 1] = h[1];                                    \
        if (!(CHECKED) || i + 2 < nk) my_out[i + 2] = h[2];                                    \
        if (!(CHECKED) || i + 3 < nk) my_out[i + 3] = h[3];                                    \
        w_cur = w_nxt;                                                                         \
        w_nxt = __funnelshift_r(raw_a, raw_b, sh);                                             \
        raw_a = raw_b;                                                                         \
        raw_b = *swp++;                                                                        \
        i += 4;                                                                                \
    }

// lut_stride == 4 arrives as a kernel argument so that the table address byte*4 + base
// stays an IMAD (FMA pipe) instead of being strength-reduced to LEA (ALU pipe).
template <int K, int R>
__global__ void __launch_bounds__(R)
sketch_fill_uniform_kernel(const uint8_t *__restrict__ bases, const SketchDst dst,
                           uint32_t L, uint32_t nk, uint32_t lut_stride) {
    constexpr int NB = K / 4;    // 4-byte body blocks per k-mer
    constexpr int TAIL = K % 4;  // tail bytes per k-mer
    constexpr uint32_t TAILMASK = TAIL == 1 ? 0xffu : TAIL == 2 ? 0xffffu : 0xffffffu;
    constexpr bool LUT = TAIL == 1;
    static_assert(NB >= 1, "fast path needs k >= 4");

    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(16) uint32_t s_lut[LUT ? 256 : 4];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem);
    uint8_t *s_in = smem + 16;
    uint32_t *s_out = reinterpret_cast<uint32_t *>(s_in + k1_in_bytes(R, L));

    const uint32_t tid = threadIdx.x;
    const uint64_t tile = blockIdx.x;
    const uint32_t in_bytes = R * L;  // multiple of 16 (host guarantees)

    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
        mbar_expect_tx(bar, in_bytes + (LUT ? 1024u : 0u));
        bulk_g2s(s_in, bases + tile * in_bytes, in_bytes, bar);
        if (LUT) bulk_g2s(s_lut, &g_kmix_byte, 1024u, bar);
    }
    __syncthreads();  // barrier init visible to the waiters
    mbar_wait(bar, 0);

    // my read: bytes [tid*L, tid*L+L) of the tile, realigned to words on the fly
    const uint32_t b0 = tid * L;
    const uint32_t *sw = reinterpret_cast<const uint32_t *>(s_in) + (b0 >> 2);
    const uint8_t *sb = s_in + b0 + 4 * NB;  // tail byte of k-mer i is sb[i]
    const uint32_t sh = (b0 & 3u) * 8u;
    uint32_t *my_out = s_out + tid * nk;
    const uint32_t lut_base = smem_u32(s_lut);

    uint32_t raw_a = sw[0], raw_b = sw[1];
    uint32_t w_cur = __funnelshift_r(raw_a, raw_b, sh);  // bytes 4Q .. 4Q+3 of the read
    raw_a = raw_b; raw_b = sw[2];
    uint32_t w_nxt = __funnelshift_r(raw_a, raw_b, sh);  // bytes 4Q+4 .. 4Q+7
    raw_a = raw_b; raw_b = sw[3];                        // one word ahead of its use
    const uint32_t *swp = sw + 4;

    uint32_t ring[4][NB];  // ring[r][*]: pre-mixes K(4q+r) of the last NB word steps
    // prologue: word steps 0 .. NB-1 only fill the ring
#pragma unroll
    for (int q = 0; q < NB; ++q) {
        ring[0][q] = mm3_kmix(w_cur);
        ring[1][q] = mm3_kmix(__funnelshift_r(w_cur, w_nxt, 8));
        ring[2][q] = mm3_kmix(__funnelshift_r(w_cur, w_nxt, 16));
        ring[3][q] = mm3_kmix(__funnelshift_r(w_cur, w_nxt, 24));
        w_cur = w_nxt;
        w_nxt = __funnelshift_r(raw_a, raw_b, sh);
        raw_a = raw_b;
        raw_b = *swp++;
    }

    // At word step Q = NB + i/4 the k-mers i+r (r = 0..3) complete: their body blocks are
    // ring[r][oldest .. newest], their tail bytes start at position i + r + 4*NB.
    uint32_t i = 0;
    const uint32_t n_main = (nk / (4 * NB)) * (4 * NB);
    while (i < n_main) {  // full groups of NB word steps: no bounds tests
#pragma unroll
        for (int u = 0; u < NB; ++u) PG_K1_STEP(u, false)
    }
#pragma unroll
    for (int u = 0; u < NB; ++u) {  // remainder: < NB full steps plus possibly a partial one
        if (i < nk) PG_K1_STEP(u, true)
    }

    // hand the staged tile to the async proxy and bulk-store it (positional, mash.go:81-84)
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
        // one bulk store per destination: the local buffer, or (fused all-gather) the gathered
        // buffer of every rank, peers reached through their NVLink-mapped addresses
        for (int p = 0; p < dst.n; ++p)
            bulk_s2g(dst.ptr[p] + tile * (uint64_t)(R * nk), s_out, R * nk * 4u);  // multiple of 16 (R % 4 == 0)
        bulk_wait_read0();
    }
}
#undef PG_K1_STEP

// ---- generic fill path -----------------------------------------------------------
// One warp per read.  Handles reads with n = max(len-k,0) < s (others are left to the
// select kernel).  Also writes count / zero padding / status for those reads.
__global__ void __launch_bounds__(256)
sketch_fill_generic_kernel(const uint8_t *__restrict__ bases, const uint64_t *__restrict__ offsets,
                           uint32_t uniform_len, uint64_t n_reads, uint64_t first_read, uint32_t k,
                           uint32_t s, uint32_t flags, uint32_t *__restrict__ out,
                           uint64_t row_stride, uint32_t *__restrict__ count,
                           int32_t *__restrict__ status) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t r = (((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); r < n_reads;
         r += warps) {
        const uint64_t row = first_read + r;
        uint64_t beg, len;
        if (offsets) {
            beg = offsets[row];
            len = offsets[row + 1] - beg;
        } else {
            beg = row * (uint64_t)uniform_len;
            len = uniform_len;
        }
        const uint64_t n = len > k ? len - k : 0;  // mash.go:73
        if (n >= s && !(n == 0 && s == 0)) continue;  // select regime (incl. s in {0,1} panics)
        const uint8_t *seq = bases + beg;
        uint32_t *dst = out + row * row_stride;
        for (uint64_t i = lane; i < n; i += 32) {
            const uint8_t *p = seq + i;
            dst[i] = mm3_bytes([p](uint32_t j) { return __ldg(p + j); }, k);
        }
        if (flags & PG_SKETCH_PAD_ZERO)
            for (uint64_t i = n + lane; i < s; i += 32) dst[i] = 0u;
        if (lane == 0) {
            if (count) count[row] = (uint32_t)n;
            if (status) status[row] = PG_ITEM_OK;
        }
    }
}

// ---- launchers -------------------------------------------------------------------
template <int K, int R>
static int launch_k1(const uint8_t *d_bases, uint64_t n_tiles, uint32_t L, uint32_t nk,
                     const SketchDst &dst, cudaStream_t st) {
    const size_t smem = 16 + k1_in_bytes(R, L) + (size_t)R * nk * 4;
    static size_t configured = 0;  // per instantiation (one process = one device)
    if (smem > configured) {
        PG_CUDA(cudaFuncSetAttribute(sketch_fill_uniform_kernel<K, R>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    sketch_fill_uniform_kernel<K, R><<<(unsigned)n_tiles, R, smem, st>>>(d_bases, dst, L, nk, 4u);
    PG_LAUNCH_CHECK("sketch_fill_uniform_kernel");
    return PG_OK;
}

template <int R>
static int dispatch_k1(int k, const uint8_t *d_bases, uint64_t n_tiles, uint32_t L, uint32_t nk,
                       const SketchDst &dst, cudaStream_t st, bool *handled) {
    *handled = true;
    switch (k) {
#define PG_K1_CASE(KK) \
    case KK: return launch_k1<KK, R>(d_bases, n_tiles, L, nk, dst, st);
        PG_K1_CASE(11) PG_K1_CASE(13) PG_K1_CASE(15) PG_K1_CASE(16) PG_K1_CASE(17) PG_K1_CASE(19)
        PG_K1_CASE(21) PG_K1_CASE(23) PG_K1_CASE(24) PG_K1_CASE(25) PG_K1_CASE(27) PG_K1_CASE(29)
        PG_K1_CASE(31) PG_K1_CASE(32)
#undef PG_K1_CASE
    default: *handled = false; return PG_OK;
    }
}

// reads per tile == threads per CTA.  32 (one warp per CTA, ~10 CTAs per SM at L=150, k=21)
// measured fastest: 64 loses ~9 % (profiles/r01_k1_tuning.md).
constexpr int K1_R = 32;

static int launch_fill_generic(const uint8_t *d_bases, const uint64_t *d_offsets, uint32_t ulen,
                               uint64_t n_reads, uint64_t first_read, int k, int s, uint32_t flags,
                               uint32_t *d_out, uint64_t row_stride, uint32_t *d_count,
                               int32_t *d_status, cudaStream_t st) {
    if (n_reads == 0) return PG_OK;
    const uint64_t blocks = std::min<uint64_t>((n_reads + 7) / 8, (uint64_t)sm_count() * 32);
    sketch_fill_generic_kernel<<<(unsigned)blocks, 256, 0, st>>>(
        d_bases, d_offsets, ulen, n_reads, first_read, (uint32_t)k, (uint32_t)s, flags, d_out,
        row_stride, d_count, d_status);
    PG_LAUNCH_CHECK("sketch_fill_generic_kernel");
    return PG_OK;
}

int launch_sketch_uniform(const uint8_t *d_bases, uint64_t n_reads, uint32_t L, int k, int s,
                          uint32_t flags, uint32_t *d_out, uint64_t row_stride, int32_t *d_status,
                          cudaStream_t st, const SketchDst *extra) {
    if (n_reads == 0) return PG_OK;
    const uint64_t n = L > (uint32_t)k ? L - (uint32_t)k : 0;
    if (n >= (uint64_t)s && !(n == 0 && s == 0))  // select regime, mash.go:87-102
        return launch_sketch_select(d_bases, nullptr, L, n_reads, k, s, flags, d_out, row_stride,
                                    nullptr, d_status, st, extra);
    // fill regime
    const uint32_t nk = (uint32_t)n;
    uint64_t done = 0;
    const bool compact = row_stride == nk && !(flags & PG_SKETCH_PAD_ZERO);
    bool aligned = ((uintptr_t)d_bases % 16 == 0) && ((uintptr_t)d_out % 16 == 0);
    if (extra)
        for (int p = 0; p < extra->n; ++p) aligned = aligned && ((uintptr_t)extra->ptr[p] % 16 == 0);
    const size_t smem = 16 + k1_in_bytes(K1_R, L) + (size_t)K1_R * nk * 4;
    if (compact && aligned && nk > 0 && (K1_R * (uint64_t)L) % 16 == 0 && smem <= 200 * 1024 &&
        n_reads >= K1_R) {
        const uint64_t tiles = n_reads / K1_R;
        bool handled = false;
        uint64_t t0 = 0;
        while (t0 < tiles) {  // grid.x limit
            const uint64_t nt = std::min<uint64_t>(tiles - t0, 0x7fffffffull);
            SketchDst dst;
            dst.n = extra ? extra->n : 1;
            for (int p = 0; p < dst.n; ++p)
                dst.ptr[p] = (extra ? extra->ptr[p] : d_out) + t0 * K1_R * (uint64_t)nk;
            int rc = dispatch_k1<K1_R>(k, d_bases + t0 * K1_R * (uint64_t)L, nt, L, nk, dst, st, &handled);
            if (rc != PG_OK) return rc;
            if (!handled) break;
            t0 += nt;
        }
        if (handled) {
            done = tiles * K1_R;
            if (d_status) PG_CUDA(cudaMemsetAsync(d_status, 0, done * sizeof(int32_t), st));
        }
    }
    if (done < n_reads) {
        if (extra) {
            set_error("fused sketch+gather needs the TMA fast path (k instantiated, n_local %% 32 == 0, 16-byte aligned buffers)");
            return PG_ERR_UNSUPPORTED;
        }
        return launch_fill_generic(d_bases, nullptr, L, n_reads - done, done, k, s, flags, d_out,
                                   row_stride, nullptr, d_status, st);
    }
    return PG_OK;
}

int launch_sketch_ragged(const uint8_t *d_bases, const uint64_t *d_offsets, uint64_t n_reads,
                         uint64_t max_read_len, int k, int s, uint32_t flags, uint32_t *d_out,
                         uint64_t row_stride, uint32_t *d_count, int32_t *d_status,
                         cudaStream_t st) {
    if (n_reads == 0) return PG_OK;
    int rc = launch_fill_generic(d_bases, d_offsets, 0, n_reads, 0, k, s, flags, d_out, row_stride,
                                 d_count, d_status, st);
    if (rc != PG_OK) return rc;
    const uint64_t nmax = max_read_len > (uint64_t)k ? max_read_len - (uint64_t)k : 0;
    if (nmax >= (uint64_t)s && nmax > 0)  // some read may be in the select regime
        return launch_sketch_select(d_bases, d_offsets, 0, n_reads, k, s, flags, d_out, row_stride,
                                    d_count, d_status, st);
    return PG_OK;
}

}  // namespace pg
