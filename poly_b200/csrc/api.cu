// api.cu -- the C ABI of libpolyb200.so (include/poly_b200.h): argument checking,
// device selection, host<->device staging for the host-pointer entry points, and
// dispatch to the kernel launchers.  No compute happens on the host: when no sm_100
// device is usable every compute entry point fails with PG_ERR_NO_DEVICE.
#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace pg {

static thread_local char t_err[512] = "";
static thread_local const char *t_last_kernel = "";
static std::atomic<uint64_t> g_launches{0};
static std::mutex g_mu;         // serialises the host-pointer entry points
static int g_device = -1;       // bound device (-1: not initialised)
static int g_sms = 0;
static cudaStream_t g_streams[3] = {nullptr, nullptr, nullptr};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof t_err, fmt, ap);
    va_end(ap);
}
int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
    set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    return PG_ERR_CUDA;
}
void note_launch(const char *name) {
    t_last_kernel = name;
    g_launches.fetch_add(1, std::memory_order_relaxed);
}
int sm_count() { return g_sms > 0 ? g_sms : 148; }

static int init_locked(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        set_error("no CUDA device available (%s); libpolyb200 has no CPU fallback",
                  e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        cudaGetLastError();
        return PG_ERR_NO_DEVICE;
    }
    if (device < 0) {
        if (cudaGetDevice(&device) != cudaSuccess) device = 0;
    }
    if (device >= n) {
        set_error("device %d out of range (count %d)", device, n);
        return PG_ERR_ARG;
    }
    cudaDeviceProp prop;
    PG_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("device %d is sm_%d%d; libpolyb200 is built for sm_100a only", device, prop.major,
                  prop.minor);
        return PG_ERR_NO_DEVICE;
    }
    PG_CUDA(cudaSetDevice(device));
    if (g_device != device) {
        for (int i = 0; i < 3; ++i) {
            if (g_streams[i]) { cudaStreamDestroy(g_streams[i]); g_streams[i] = nullptr; }
        }
    }
    for (int i = 0; i < 3; ++i) {
        if (!g_streams[i]) PG_CUDA(cudaStreamCreateWithFlags(&g_streams[i], cudaStreamNonBlocking));
    }
    {   // Stream-ordered temporaries (cudaMallocAsync) are freed at the end of every call and most calls
        // synchronise: with the default release threshold of 0 the pool would hand its memory back to the
        // driver each time.  Keep up to 1 GiB cached so that small calls do not pay a driver allocation.
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t keep = 1ull << 30, cur = 0;
            if (cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &cur) == cudaSuccess && cur < keep)
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    g_device = device;
    g_sms = prop.multiProcessorCount;
    return PG_OK;
}

// every entry point: make sure a device is bound and current on this thread
static int ensure_device() {
    if (g_device < 0) {
        std::lock_guard<std::mutex> lk(g_mu);
        if (g_device < 0) {
            int rc = init_locked(-1);
            if (rc != PG_OK) return rc;
        }
    }
    PG_CUDA(cudaSetDevice(g_device));
    return PG_OK;
}

// grow-only device scratch used by the host-pointer entry points (under g_mu)
struct Scratch {
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return PG_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) {
            set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
            cudaGetLastError();
            return PG_ERR_NOMEM;
        }
        cap = bytes;
        return PG_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};
static Scratch g_in[3], g_out[3], g_aux[3], g_st[3];

static inline uint64_t kmers_of(uint64_t len, int k) { return len > (uint64_t)k ? len - (uint64_t)k : 0; }

}  // namespace pg

// small RAII helper for temporary device buffers of the non-pipelined host paths
namespace pg {
namespace {
struct Tmp {
    void *p = nullptr;
    cudaStream_t st;
    explicit Tmp(cudaStream_t s) : st(s) {}
    int alloc(size_t bytes) {
        cudaError_t e = cudaMallocAsync(&p, bytes ? bytes : 16, st);
        if (e != cudaSuccess) { set_error("cudaMallocAsync(%zu): %s", bytes, cudaGetErrorString(e)); cudaGetLastError(); return PG_ERR_NOMEM; }
        return PG_OK;
    }
    ~Tmp() { if (p) cudaFreeAsync(p, st); }
    template <typename T> T *as() { return (T *)p; }
};
}  // namespace
}  // namespace pg

using namespace pg;

extern "C" {

int pg_version(void) { return 100; }

int pg_init(int device) {
    std::lock_guard<std::mutex> lk(g_mu);
    return init_locked(device);
}

int pg_shutdown(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_device >= 0) {
        cudaSetDevice(g_device);
        cudaDeviceSynchronize();
        for (int i = 0; i < 3; ++i) {
            g_in[i].release(); g_out[i].release(); g_aux[i].release(); g_st[i].release();
            if (g_streams[i]) { cudaStreamDestroy(g_streams[i]); g_streams[i] = nullptr; }
        }
    }
    g_device = -1;
    return PG_OK;
}

const char *pg_last_error(void) { return t_err; }
const char *pg_last_kernel(void) { return t_last_kernel; }
uint64_t pg_launch_count(void) { return g_launches.load(); }

int pg_device_count(int *count) {
    if (!count) return PG_ERR_ARG;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); n = 0; }
    *count = n;
    return PG_OK;
}
int pg_device_sm_count(int *sms) {
    if (!sms) return PG_ERR_ARG;
    int rc = ensure_device();
    if (rc != PG_OK) return rc;
    *sms = g_sms;
    return PG_OK;
}

int pg_host_alloc(void **ptr, size_t bytes) {
    if (!ptr) return PG_ERR_ARG;
    int rc = ensure_device();
    if (rc != PG_OK) return rc;
    PG_CUDA(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return PG_OK;
}
int pg_host_free(void *ptr) {
    if (ptr) PG_CUDA(cudaFreeHost(ptr));
    return PG_OK;
}
int pg_dev_alloc(void **dptr, size_t bytes) {
    if (!dptr) return PG_ERR_ARG;
    int rc = ensure_device();
    if (rc != PG_OK) return rc;
    PG_CUDA(cudaMalloc(dptr, bytes ? bytes : 1));
    return PG_OK;
}
int pg_dev_free(void *dptr) {
    if (dptr) PG_CUDA(cudaFree(dptr));
    return PG_OK;
}
int pg_memcpy_h2d(void *dst, const void *src, size_t bytes, void *stream) {
    int rc = ensure_device();
    if (rc != PG_OK) return rc;
    PG_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return PG_OK;
}
int pg_memcpy_d2h(void *dst, const void *src, size_t bytes, void *stream) {
    int rc = ensure_device();
    if (rc != PG_OK) return rc;
    PG_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return PG_OK;
}
int pg_stream_sync(void *stream) {
    int rc = ensure_device();
    if (rc != PG_OK) return rc;
    PG_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return PG_OK;
}

// ---------------------------------------------------------------------------------
// mash.Sketch
// ---------------------------------------------------------------------------------
static int check_ks(int32_t k, int32_t s) {
    if (k < 0) { set_error("kmerSize %d < 0: the reference panics (slice bounds)", k); return PG_ERR_ARG; }
    if (s < 0) { set_error("sketchSize %d < 0: the reference panics (makeslice)", s); return PG_ERR_ARG; }
    return PG_OK;
}

int pg_mash_sketch_uniform_dev(const uint8_t *d_bases, uint64_t n_reads, uint32_t read_len,
                               int32_t k, int32_t s, uint32_t flags, uint32_t *d_out,
                               uint64_t row_stride, int32_t *d_status, void *stream) {
    int rc = ensure_device();
    if (rc != PG_OK) return rc;
    if ((rc = check_ks(k, s)) != PG_OK) return rc;
    if (n_reads && (!d_bases || !d_out)) { set_error("null buffer"); return PG_ERR_ARG; }
    const uint64_t cnt = std::min<uint64_t>(kmers_of(read_len, k), (uint64_t)s);
    const uint64_t need = (flags & PG_SKETCH_PAD_ZERO) ? (uint64_t)s : cnt;
    if (row_stride < need) { set_error("row_stride %llu < %llu", (unsigned long long)row_stride, (unsigned long long)need); return PG_ERR_ARG; }
    return launch_sketch_uniform(d_bases, n_reads, read_len, k, s, flags, d_out, row_stride,
                                 d_status, (cudaStream_t)stream);
}

int pg_mash_sketch_batch_dev(const uint8_t *d_bases, const uint64_t *d_offsets, uint64_t n_reads,
                             uint64_t max_read_len, int32_t k, int32_t s, uint32_t flags,
                             uint32_t *d_out, uint64_t row_stride, uint32_t *d_count,
                             int32_t *d_status, void *stream) {
    int rc = ensure_device();
    if (rc != PG_OK) return rc;
    if ((rc = check_ks(k, s)) != PG_OK) return rc;
    if (n_reads && (!d_offsets || !d_out)) { set_error("null buffer"); return PG_ERR_ARG; }
    const uint64_t cnt = std::min<uint64_t>(kmers_of(max_read_len, k), (uint64_t)s);
    const uint64_t need = (flags & PG_SKETCH_PAD_ZERO) ? (uint64_t)s : cnt;
    if (row_stride < need) { set_error("row_stride %llu < %llu", (unsigned long long)row_stride, (unsigned long long)need); return PG_ERR_ARG; }
    return launch_sketch_ragged(d_bases, d_offsets, n_reads, max_read_len, k, s, flags, d_out,
                                row_stride, d_count, d_status, (cudaStream_t)stream);
}

int pg_ipc_export(void *dptr, uint8_t handle[PG_IPC_HANDLE_BYTES]) {
    int rc = ensure_device();
    if (rc != PG_OK) return rc;
    static_assert(sizeof(cudaIpcMemHandle_t) == PG_IPC_HANDLE_BYTES, "handle size");
    cudaIpcMemHandle_t h;
    PG_CUDA(cudaIpcGetMemHandle(&h, dptr));
    memcpy(handle, &h, sizeof h);
    return PG_OK;
}
int pg_ipc_import(const uint8_t handle[PG_IPC_HANDLE_BYTES], void **dptr) {
    int rc = ensure_device();
    if (rc != PG_OK) return rc;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    PG_CUDA(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
    return PG_OK;
}
int pg_ipc_close(void *dptr) {
    if (dptr) PG_CUDA(cudaIpcCloseMemHandle(dptr));
    return PG_OK;
}

int pg_mash_sketch_uniform_gather_dev(const uint8_t *d_bases, uint64_t n_local, uint32_t read_len,
                                      int32_t k, int32_t s, void *const *gathered_ptrs,
                                      int32_t world, int32_t rank, void *stream) {
    int rc = ensure_device();
    if (rc != PG_OK) return rc;
    if ((rc = check_ks(k, s)) != PG_OK) return rc;
    if (world < 1 || world > PG_MAX_PEERS || rank < 0 || rank >= world || !gathered_ptrs) {
        set_error("bad world/rank (%d/%d)", world, rank);
        return PG_ERR_ARG;
    }
    const uint64_t cnt = std::min<uint64_t>(kmers_of(read_len, k), (uint64_t)s);
    SketchDst dst;
    dst.n = world;
    for (int p = 0; p < world; ++p) {
        if (!gathered_ptrs[p]) { set_error("null gathered pointer for rank %d", p); return PG_ERR_ARG; }
        dst.ptr[p] = (uint32_t *)gathered_ptrs[p] + (uint64_t)rank * n_local * cnt;  // this rank's row block
    }
    return launch_sketch_uniform(d_bases, n_local, read_len, k, s, 0, dst.ptr[rank], cnt, nullptr,
                                 (cudaStream_t)stream, &dst);
}

// Pipelined host path shared by the uniform and ragged entry points: chunks of reads
// cycle through 3 stream/buffer slots (H2D, kernel, D2H overlap across slots).
static int sketch_host(const uint8_t *bases, const uint64_t *offsets, uint32_t ulen, uint64_t n_reads,
                       int32_t k, int32_t s, uint32_t flags, uint32_t *out, uint64_t row_stride,
                       uint32_t *count, int32_t *status) {
    int rc = ensure_device();
    if (rc != PG_OK) return rc;
    if ((rc = check_ks(k, s)) != PG_OK) return rc;
    if (n_reads == 0) return PG_OK;
    if (!bases || !out) { set_error("null buffer"); return PG_ERR_ARG; }
    std::lock_guard<std::mutex> lk(g_mu);
    const bool want_status = status != nullptr || s <= 1;
    const uint64_t target_bytes = 192ull << 20;  // input bytes per chunk
    bool any_panic = false;
    std::vector<int32_t> st_host;
    uint64_t r0 = 0;
    int slot = 0;
    struct Pending { uint64_t r0, nr; bool used; } pend[3] = {{0, 0, false}, {0, 0, false}, {0, 0, false}};
    std::vector<int32_t> st_slot[3];
    auto drain = [&](int sl) -> int {
        if (!pend[sl].used) return PG_OK;
        PG_CUDA(cudaStreamSynchronize(g_streams[sl]));
        if (want_status) {
            for (uint64_t i = 0; i < pend[sl].nr; ++i) {
                if (st_slot[sl][i] != PG_ITEM_OK) any_panic = true;
                if (status) status[pend[sl].r0 + i] = st_slot[sl][i];
            }
        }
        pend[sl].used = false;
        return PG_OK;
    };
    while (r0 < n_reads) {
        // chunk [r0, r1)
        uint64_t r1, beg, end, maxlen, minlen;
        if (offsets) {
            beg = offsets[r0];
            r1 = r0;
            maxlen = 0;
            minlen = ~0ull;
            while (r1 < n_reads && (r1 == r0 || offsets[r1 + 1] - beg <= target_bytes)) {
                if (offsets[r1 + 1] < offsets[r1]) { set_error("offsets not monotone at %llu", (unsigned long long)r1); return PG_ERR_ARG; }
                const uint64_t len = offsets[r1 + 1] - offsets[r1];
                maxlen = std::max(maxlen, len);
                minlen = std::min(minlen, len);
                ++r1;
            }
            end = offsets[r1];
        } else {
            uint64_t per = std::max<uint64_t>(1, target_bytes / std::max<uint32_t>(ulen, 1));
            per = (per + 31) & ~31ull;  // keep chunk starts on K1 tile boundaries
            r1 = std::min(n_reads, r0 + per);
            beg = r0 * (uint64_t)ulen;
            end = r1 * (uint64_t)ulen;
            maxlen = minlen = ulen;
        }
        const uint64_t nr = r1 - r0;
        const bool uniform = maxlen == minlen && maxlen <= 0xffffffffull;
        const uint64_t cnt_max = std::min<uint64_t>(kmers_of(maxlen, k), (uint64_t)s);
        const uint64_t dev_stride = (flags & PG_SKETCH_PAD_ZERO) ? (uint64_t)s : cnt_max;
        if (row_stride < dev_stride) { set_error("row_stride %llu < %llu", (unsigned long long)row_stride, (unsigned long long)dev_stride); return PG_ERR_ARG; }

        if ((rc = drain(slot)) != PG_OK) return rc;
        cudaStream_t stx = g_streams[slot];
        const uint64_t in_bytes = end - beg;
        if ((rc = g_in[slot].reserve(in_bytes + 64)) != PG_OK) return rc;
        if ((rc = g_out[slot].reserve(std::max<uint64_t>(nr * dev_stride * 4, 16))) != PG_OK) return rc;
        uint8_t *d_in = (uint8_t *)g_in[slot].p;
        uint32_t *d_out = (uint32_t *)g_out[slot].p;
        int32_t *d_status = nullptr;
        uint32_t *d_count = nullptr;
        uint64_t *d_off = nullptr;
        if (want_status) {
            if ((rc = g_st[slot].reserve(nr * 4)) != PG_OK) return rc;
            d_status = (int32_t *)g_st[slot].p;
            PG_CUDA(cudaMemsetAsync(d_status, 0, nr * 4, stx));
            st_slot[slot].assign(nr, 0);
        }
        if (in_bytes) PG_CUDA(cudaMemcpyAsync(d_in, bases + beg, in_bytes, cudaMemcpyHostToDevice, stx));
        if (uniform) {
            rc = launch_sketch_uniform(d_in, nr, (uint32_t)maxlen, k, s, flags, d_out, dev_stride, d_status, stx);
        } else {
            if ((rc = g_aux[slot].reserve((nr + 1) * 8 + nr * 4)) != PG_OK) return rc;
            d_off = (uint64_t *)g_aux[slot].p;
            d_count = (uint32_t *)(d_off + nr + 1);
            PG_CUDA(cudaMemcpyAsync(d_off, offsets + r0, (nr + 1) * 8, cudaMemcpyHostToDevice, stx));
            // kernels index bases with absolute offsets: shift the base pointer
            rc = launch_sketch_ragged(d_in - beg, d_off, nr, maxlen, k, s, flags, d_out, dev_stride, d_count, d_status, stx);
        }
        if (rc != PG_OK) return rc;
        if (dev_stride) {
            if (row_stride == dev_stride)
                PG_CUDA(cudaMemcpyAsync(out + r0 * row_stride, d_out, nr * dev_stride * 4, cudaMemcpyDeviceToHost, stx));
            else
                PG_CUDA(cudaMemcpy2DAsync(out + r0 * row_stride, row_stride * 4, d_out, dev_stride * 4, dev_stride * 4, nr, cudaMemcpyDeviceToHost, stx));
        }
        if (count) {
            if (uniform) {
                for (uint64_t i = 0; i < nr; ++i) count[r0 + i] = (uint32_t)cnt_max;
            } else {
                PG_CUDA(cudaMemcpyAsync(count + r0, d_count, nr * 4, cudaMemcpyDeviceToHost, stx));
            }
        }
        if (want_status) PG_CUDA(cudaMemcpyAsync(st_slot[slot].data(), d_status, nr * 4, cudaMemcpyDeviceToHost, stx));
        pend[slot] = {r0, nr, true};
        slot = (slot + 1) % 3;
        r0 = r1;
    }
    for (int i = 0; i < 3; ++i)
        if ((rc = drain(i)) != PG_OK) return rc;
    if (any_panic) {
        set_error("at least one read hits an input on which mash.Sketch panics (sketchSize <= 1)");
        return PG_ERR_PANIC;
    }
    return PG_OK;
}

int pg_mash_sketch_batch(const uint8_t *bases, const uint64_t *offsets, uint64_t n_reads, int32_t k,
                         int32_t s, uint32_t flags, uint32_t *out, uint64_t row_stride,
                         uint32_t *count, int32_t *status) {
    if (n_reads && !offsets) { set_error("null offsets"); return PG_ERR_ARG; }
    return sketch_host(bases, offsets, 0, n_reads, k, s, flags, out, row_stride, count, status);
}

int pg_mash_sketch_uniform(const uint8_t *bases, uint64_t n_reads, uint32_t read_len, int32_t k,
                           int32_t s, uint32_t flags, uint32_t *out, uint64_t row_stride,
                           int32_t *status) {
    return sketch_host(bases, nullptr, read_len, n_reads, k, s, flags, out, row_stride, nullptr, status);
}

// ---------------------------------------------------------------------------------
// Similarity / Distance
// ---------------------------------------------------------------------------------
int pg_mash_similarity_pairs_dev(const uint32_t *d_sketches, const uint64_t *d_sk_offsets,
                                 uint64_t n_sketches, const uint32_t *d_pair_a,
                                 const uint32_t *d_pair_b, uint64_t n_pairs, int64_t *d_same,
                                 double *d_similarity, double *d_distance, int32_t *d_status,
                                 void *stream) {
    int rc = ensure_device();
    if (rc != PG_OK) return rc;
    return launch_similarity_pairs(d_sketches, d_sk_offsets, n_sketches, d_pair_a, d_pair_b, n_pairs,
                                   d_same, d_similarity, d_distance, d_status, (cudaStream_t)stream);
}


int pg_mash_similarity_pairs(const uint32_t *sketches, const uint64_t *sk_offsets,
                             uint64_t n_sketches, const uint32_t *pair_a, const uint32_t *pair_b,
                             uint64_t n_pairs, int64_t *same, double *similarity, double *distance,
                             int32_t *status) {
    int rc = ensure_device();
    if (rc != PG_OK) return rc;
    if (n_pairs == 0) return PG_OK;
    if (!sk_offsets || !pair_a || !pair_b) { set_error("null buffer"); return PG_ERR_ARG; }
    std::lock_guard<std::mutex> lk(g_mu);
    cudaStream_t st = g_streams[0];
    const uint64_t words = sk_offsets[n_sketches];
    Tmp d_sk(st), d_off(st), d_a(st), d_b(st), d_same(st), d_sim(st), d_dist(st), d_st(st);
    if ((rc = d_sk.alloc(words * 4)) || (rc = d_off.alloc((n_sketches + 1) * 8)) ||
        (rc = d_a.alloc(n_pairs * 4)) || (rc = d_b.alloc(n_pairs * 4)) ||
        (rc = d_same.alloc(n_pairs * 8)) || (rc = d_sim.alloc(n_pairs * 8)) ||
        (rc = d_dist.alloc(n_pairs * 8)) || (rc = d_st.alloc(n_pairs * 4)))
        return rc;
    if (words) PG_CUDA(cudaMemcpyAsync(d_sk.p, sketches, words * 4, cudaMemcpyHostToDevice, st));
    PG_CUDA(cudaMemcpyAsync(d_off.p, sk_offsets, (n_sketches + 1) * 8, cudaMemcpyHostToDevice, st));
    PG_CUDA(cudaMemcpyAsync(d_a.p, pair_a, n_pairs * 4, cudaMemcpyHostToDevice, st));
    PG_CUDA(cudaMemcpyAsync(d_b.p, pair_b, n_pairs * 4, cudaMemcpyHostToDevice, st));
    PG_CUDA(cudaMemsetAsync(d_st.p, 0, n_pairs * 4, st));
    rc = launch_similarity_pairs(d_sk.as<uint32_t>(), d_off.as<uint64_t>(), n_sketches, d_a.as<uint32_t>(),
                                 d_b.as<uint32_t>(), n_pairs, d_same.as<int64_t>(), d_sim.as<double>(),
                                 d_dist.as<double>(), d_st.as<int32_t>(), st);
    if (rc != PG_OK) return rc;
    std::vector<int32_t> hst(n_pairs);
    if (same) PG_CUDA(cudaMemcpyAsync(same, d_same.p, n_pairs * 8, cudaMemcpyDeviceToHost, st));
    if (similarity) PG_CUDA(cudaMemcpyAsync(similarity, d_sim.p, n_pairs * 8, cudaMemcpyDeviceToHost, st));
    if (distance) PG_CUDA(cudaMemcpyAsync(distance, d_dist.p, n_pairs * 8, cudaMemcpyDeviceToHost, st));
    PG_CUDA(cudaMemcpyAsync(hst.data(), d_st.p, n_pairs * 4, cudaMemcpyDeviceToHost, st));
    PG_CUDA(cudaStreamSynchronize(st));
    bool panic = false;
    for (uint64_t i = 0; i < n_pairs; ++i) {
        if (status) status[i] = hst[i];
        panic |= hst[i] != PG_ITEM_OK;
    }
    if (panic) { set_error("at least one pair involves an empty sketch or an invalid index (reference panics)"); return PG_ERR_PANIC; }
    return PG_OK;
}

int pg_mash_distance_block_dev(const uint32_t *d_sketches, uint64_t n, int32_t s, uint64_t row_begin,
                               uint64_t row_end, uint32_t *d_same, double *d_distance, void *stream) {
    int rc = ensure_device();
    if (rc != PG_OK) return rc;
    if (row_end > n || row_begin > row_end) { set_error("bad row range"); return PG_ERR_ARG; }
    return launch_distance_block(d_sketches, n, s, row_begin, row_end, d_same, d_distance, (cudaStream_t)stream);
}

int pg_mash_distance_block(const uint32_t *sketches, uint64_t n, int32_t s, uint64_t row_begin,
                           uint64_t row_end, uint32_t *same, double *distance) {
    int rc = ensure_device();
    if (rc != PG_OK) return rc;
    if (row_end > n || row_begin > row_end) { set_error("bad row range"); return PG_ERR_ARG; }
    if (row_end == row_begin || n == 0) return PG_OK;
    if (s <= 0) { set_error("distance over sketches of size %d: the reference panics", s); return PG_ERR_PANIC; }
    std::lock_guard<std::mutex> lk(g_mu);
    cudaStream_t st = g_streams[0];
    Tmp d_sk(st);
    if ((rc = d_sk.alloc(n * (uint64_t)s * 4))) return rc;
    PG_CUDA(cudaMemcpyAsync(d_sk.p, sketches, n * (uint64_t)s * 4, cudaMemcpyHostToDevice, st));
    // one plan (sortedness + join index) for the whole call; row blocks of <= 64 Mi pairs per pass
    DistancePlan plan;
    rc = distance_plan_create(d_sk.as<uint32_t>(), n, s, st, &plan);
    const uint64_t rows_per = std::max<uint64_t>(8, ((64ull << 20) / n) & ~7ull);
    for (uint64_t rb = row_begin; rb < row_end && rc == PG_OK; rb += rows_per) {
        const uint64_t re = std::min(row_end, rb + rows_per);
        Tmp d_same(st), d_dist(st);
        if (same && (rc = d_same.alloc((re - rb) * n * 4))) break;
        if (distance && (rc = d_dist.alloc((re - rb) * n * 8))) break;
        rc = distance_plan_rows(plan, rb, re, same ? d_same.as<uint32_t>() : nullptr, distance ? d_dist.as<double>() : nullptr, st);
        if (rc != PG_OK) break;
        cudaError_t e = cudaSuccess;
        if (same) e = cudaMemcpyAsync(same + (rb - row_begin) * n, d_same.p, (re - rb) * n * 4, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess && distance)
            e = cudaMemcpyAsync(distance + (rb - row_begin) * n, d_dist.p, (re - rb) * n * 8, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) rc = cuda_fail(e, "distance block copy", __FILE__, __LINE__);
    }
    distance_plan_destroy(plan, st);
    return rc;
}

// ---------------------------------------------------------------------------------
// Smith-Waterman score
// ---------------------------------------------------------------------------------
static int align_score_dev(int global, const uint8_t *d_queries, const uint64_t *d_q_offsets, uint64_t n_queries,
                          uint64_t max_query_len, const uint8_t *d_templ, uint64_t templ_len,
                          int32_t query_is_a, const int16_t *lut_a_host, const int16_t *lut_b_host,
                          const int64_t *table_host, int32_t n_a, int32_t n_b, int64_t gap,
                          int64_t *d_score, int32_t *d_err_code, int64_t *d_err_pos, void *stream) {
    int rc = ensure_device();
    if (rc != PG_OK) return rc;
    if (!lut_a_host || !lut_b_host || !table_host || (n_queries && (!d_q_offsets || !d_score))) { set_error("null buffer"); return PG_ERR_ARG; }
    return launch_sw_score(d_queries, d_q_offsets, n_queries, max_query_len, d_templ, templ_len,
                           query_is_a, lut_a_host, lut_b_host, table_host, n_a, n_b, gap, d_score,
                           d_err_code, d_err_pos, (cudaStream_t)stream, global);
}

static int align_score_host(int global, const uint8_t *queries, const uint64_t *q_offsets, uint64_t n_queries,
                      const uint8_t *templ, uint64_t templ_len, int32_t query_is_a,
                      const int16_t *lut_a, const int16_t *lut_b, const int64_t *table, int32_t n_a,
                      int32_t n_b, int64_t gap, int64_t *score, int32_t *err_code, int64_t *err_pos) {
    int rc = ensure_device();
    if (rc != PG_OK) return rc;
    if (n_queries == 0) return PG_OK;
    if (!q_offsets || !score || !lut_a || !lut_b || !table) { set_error("null buffer"); return PG_ERR_ARG; }
    std::lock_guard<std::mutex> lk(g_mu);
    cudaStream_t st = g_streams[0];
    const uint64_t q0 = q_offsets[0], qbytes = q_offsets[n_queries] - q0;
    uint64_t maxq = 0;
    for (uint64_t i = 0; i < n_queries; ++i) maxq = std::max(maxq, q_offsets[i + 1] - q_offsets[i]);
    Tmp d_q(st), d_off(st), d_t(st), d_sc(st), d_ec(st), d_ep(st);
    if ((rc = d_q.alloc(qbytes)) || (rc = d_off.alloc((n_queries + 1) * 8)) || (rc = d_t.alloc(templ_len)) ||
        (rc = d_sc.alloc(n_queries * 8)) || (rc = d_ec.alloc(n_queries * 4)) || (rc = d_ep.alloc(n_queries * 8)))
        return rc;
    if (qbytes) PG_CUDA(cudaMemcpyAsync(d_q.p, queries + q0, qbytes, cudaMemcpyHostToDevice, st));
    PG_CUDA(cudaMemcpyAsync(d_off.p, q_offsets, (n_queries + 1) * 8, cudaMemcpyHostToDevice, st));
    if (templ_len) PG_CUDA(cudaMemcpyAsync(d_t.p, templ, templ_len, cudaMemcpyHostToDevice, st));
    rc = launch_sw_score(d_q.as<uint8_t>() - q0, d_off.as<uint64_t>(), n_queries, maxq, d_t.as<uint8_t>(), templ_len,
                         query_is_a, lut_a, lut_b, table, n_a, n_b, gap, d_sc.as<int64_t>(), d_ec.as<int32_t>(),
                         d_ep.as<int64_t>(), st, global);
    if (rc != PG_OK) return rc;
    PG_CUDA(cudaMemcpyAsync(score, d_sc.p, n_queries * 8, cudaMemcpyDeviceToHost, st));
    if (err_code) PG_CUDA(cudaMemcpyAsync(err_code, d_ec.p, n_queries * 4, cudaMemcpyDeviceToHost, st));
    if (err_pos) PG_CUDA(cudaMemcpyAsync(err_pos, d_ep.p, n_queries * 8, cudaMemcpyDeviceToHost, st));
    PG_CUDA(cudaStreamSynchronize(st));
    return PG_OK;
}

int pg_sw_score_batch_dev(const uint8_t *d_queries, const uint64_t *d_q_offsets, uint64_t n_queries,
                          uint64_t max_query_len, const uint8_t *d_templ, uint64_t templ_len,
                          int32_t query_is_a, const int16_t *lut_a_host, const int16_t *lut_b_host,
                          const int64_t *table_host, int32_t n_a, int32_t n_b, int64_t gap,
                          int64_t *d_score, int32_t *d_err_code, int64_t *d_err_pos, void *stream) {
    return align_score_dev(0, d_queries, d_q_offsets, n_queries, max_query_len, d_templ, templ_len, query_is_a, lut_a_host,
                           lut_b_host, table_host, n_a, n_b, gap, d_score, d_err_code, d_err_pos, stream);
}
int pg_nw_score_batch_dev(const uint8_t *d_queries, const uint64_t *d_q_offsets, uint64_t n_queries,
                          uint64_t max_query_len, const uint8_t *d_templ, uint64_t templ_len,
                          int32_t query_is_a, const int16_t *lut_a_host, const int16_t *lut_b_host,
                          const int64_t *table_host, int32_t n_a, int32_t n_b, int64_t gap,
                          int64_t *d_score, int32_t *d_err_code, int64_t *d_err_pos, void *stream) {
    return align_score_dev(1, d_queries, d_q_offsets, n_queries, max_query_len, d_templ, templ_len, query_is_a, lut_a_host,
                           lut_b_host, table_host, n_a, n_b, gap, d_score, d_err_code, d_err_pos, stream);
}
int pg_sw_score_batch(const uint8_t *queries, const uint64_t *q_offsets, uint64_t n_queries,
                      const uint8_t *templ, uint64_t templ_len, int32_t query_is_a,
                      const int16_t *lut_a, const int16_t *lut_b, const int64_t *table, int32_t n_a,
                      int32_t n_b, int64_t gap, int64_t *score, int32_t *err_code, int64_t *err_pos) {
    return align_score_host(0, queries, q_offsets, n_queries, templ, templ_len, query_is_a, lut_a, lut_b, table, n_a, n_b, gap,
                            score, err_code, err_pos);
}
int pg_nw_score_batch(const uint8_t *queries, const uint64_t *q_offsets, uint64_t n_queries,
                      const uint8_t *templ, uint64_t templ_len, int32_t query_is_a,
                      const int16_t *lut_a, const int16_t *lut_b, const int64_t *table, int32_t n_a,
                      int32_t n_b, int64_t gap, int64_t *score, int32_t *err_code, int64_t *err_pos) {
    return align_score_host(1, queries, q_offsets, n_queries, templ, templ_len, query_is_a, lut_a, lut_b, table, n_a, n_b, gap,
                            score, err_code, err_pos);
}

static int align_strings_host(int global, const uint8_t *queries, const uint64_t *q_offsets, uint64_t n_queries,
                      const uint8_t *templ, uint64_t templ_len, int32_t query_is_a,
                      const int16_t *lut_a, const int16_t *lut_b, const int64_t *table, int32_t n_a,
                      int32_t n_b, int64_t gap, int64_t *score, int32_t *err_code, int64_t *err_pos,
                      uint8_t *align_a, uint8_t *align_b, uint64_t out_stride, uint32_t *align_len,
                      int32_t *status) {
    int rc = ensure_device();
    if (rc != PG_OK) return rc;
    if (n_queries == 0) return PG_OK;
    if (!q_offsets || !score || !lut_a || !lut_b || !table || !align_a || !align_b || !align_len || n_a <= 0 || n_b <= 0) {
        set_error("null buffer");
        return PG_ERR_ARG;
    }
    std::lock_guard<std::mutex> lk(g_mu);
    cudaStream_t st = g_streams[0];
    const uint64_t q0 = q_offsets[0], qbytes = q_offsets[n_queries] - q0;
    uint64_t maxq = 0;
    for (uint64_t i = 0; i < n_queries; ++i) maxq = std::max(maxq, q_offsets[i + 1] - q_offsets[i]);
    Tmp d_q(st), d_off(st), d_t(st), d_sc(st), d_ec(st), d_ep(st), d_a(st), d_b(st), d_len(st), d_st(st);
    if ((rc = d_q.alloc(qbytes)) || (rc = d_off.alloc((n_queries + 1) * 8)) || (rc = d_t.alloc(templ_len)) ||
        (rc = d_sc.alloc(n_queries * 8)) || (rc = d_ec.alloc(n_queries * 4)) || (rc = d_ep.alloc(n_queries * 8)) ||
        (rc = d_a.alloc(n_queries * out_stride)) || (rc = d_b.alloc(n_queries * out_stride)) ||
        (rc = d_len.alloc(n_queries * 4)) || (rc = d_st.alloc(n_queries * 4)))
        return rc;
    if (qbytes) PG_CUDA(cudaMemcpyAsync(d_q.p, queries + q0, qbytes, cudaMemcpyHostToDevice, st));
    PG_CUDA(cudaMemcpyAsync(d_off.p, q_offsets, (n_queries + 1) * 8, cudaMemcpyHostToDevice, st));
    if (templ_len) PG_CUDA(cudaMemcpyAsync(d_t.p, templ, templ_len, cudaMemcpyHostToDevice, st));
    PG_CUDA(cudaMemsetAsync(d_len.p, 0, n_queries * 4, st));
    PG_CUDA(cudaMemsetAsync(d_st.p, 0, n_queries * 4, st));
    rc = launch_sw_align(d_q.as<uint8_t>() - q0, d_off.as<uint64_t>(), n_queries, maxq, d_t.as<uint8_t>(), templ_len, query_is_a,
                         lut_a, lut_b, table, n_a, n_b, gap, d_sc.as<int64_t>(), d_ec.as<int32_t>(), d_ep.as<int64_t>(),
                         d_a.as<uint8_t>(), d_b.as<uint8_t>(), out_stride, d_len.as<uint32_t>(), d_st.as<int32_t>(), st, global);
    if (rc != PG_OK) return rc;
    PG_CUDA(cudaMemcpyAsync(score, d_sc.p, n_queries * 8, cudaMemcpyDeviceToHost, st));
    if (err_code) PG_CUDA(cudaMemcpyAsync(err_code, d_ec.p, n_queries * 4, cudaMemcpyDeviceToHost, st));
    if (err_pos) PG_CUDA(cudaMemcpyAsync(err_pos, d_ep.p, n_queries * 8, cudaMemcpyDeviceToHost, st));
    PG_CUDA(cudaMemcpyAsync(align_a, d_a.p, n_queries * out_stride, cudaMemcpyDeviceToHost, st));
    PG_CUDA(cudaMemcpyAsync(align_b, d_b.p, n_queries * out_stride, cudaMemcpyDeviceToHost, st));
    PG_CUDA(cudaMemcpyAsync(align_len, d_len.p, n_queries * 4, cudaMemcpyDeviceToHost, st));
    if (status) PG_CUDA(cudaMemcpyAsync(status, d_st.p, n_queries * 4, cudaMemcpyDeviceToHost, st));
    PG_CUDA(cudaStreamSynchronize(st));
    return PG_OK;
}

int pg_sw_align_batch(const uint8_t *queries, const uint64_t *q_offsets, uint64_t n_queries,
                      const uint8_t *templ, uint64_t templ_len, int32_t query_is_a,
                      const int16_t *lut_a, const int16_t *lut_b, const int64_t *table, int32_t n_a,
                      int32_t n_b, int64_t gap, int64_t *score, int32_t *err_code, int64_t *err_pos,
                      uint8_t *align_a, uint8_t *align_b, uint64_t out_stride, uint32_t *align_len,
                      int32_t *status) {
    return align_strings_host(0, queries, q_offsets, n_queries, templ, templ_len, query_is_a, lut_a, lut_b, table, n_a, n_b, gap,
                              score, err_code, err_pos, align_a, align_b, out_stride, align_len, status);
}
int pg_nw_align_batch(const uint8_t *queries, const uint64_t *q_offsets, uint64_t n_queries,
                      const uint8_t *templ, uint64_t templ_len, int32_t query_is_a,
                      const int16_t *lut_a, const int16_t *lut_b, const int64_t *table, int32_t n_a,
                      int32_t n_b, int64_t gap, int64_t *score, int32_t *err_code, int64_t *err_pos,
                      uint8_t *align_a, uint8_t *align_b, uint64_t out_stride, uint32_t *align_len,
                      int32_t *status) {
    return align_strings_host(1, queries, q_offsets, n_queries, templ, templ_len, query_is_a, lut_a, lut_b, table, n_a, n_b, gap,
                              score, err_code, err_pos, align_a, align_b, out_stride, align_len, status);
}

// ---------------------------------------------------------------------------------
// SantaLucia Tm
// ---------------------------------------------------------------------------------
int pg_tm_batch_dev(const uint8_t *d_bases, const uint64_t *d_offsets, uint64_t n, double cp,
                    double na, double mg, double *d_tm, double *d_dh, double *d_ds, int32_t *d_status,
                    void *stream) {
    int rc = ensure_device();
    if (rc != PG_OK) return rc;
    return launch_tm(d_bases, d_offsets, n, cp, na, mg, d_tm, d_dh, d_ds, d_status, (cudaStream_t)stream);
}

int pg_tm_batch(const uint8_t *bases, const uint64_t *offsets, uint64_t n, double cp, double na,
                double mg, double *tm, double *dh, double *ds, int32_t *status) {
    int rc = ensure_device();
    if (rc != PG_OK) return rc;
    if (n == 0) return PG_OK;
    if (!offsets) { set_error("null offsets"); return PG_ERR_ARG; }
    std::lock_guard<std::mutex> lk(g_mu);
    cudaStream_t st = g_streams[0];
    const uint64_t b0 = offsets[0], nbytes = offsets[n] - b0;
    Tmp d_b(st), d_off(st), d_tm(st), d_dh(st), d_ds(st), d_st(st);
    if ((rc = d_b.alloc(nbytes)) || (rc = d_off.alloc((n + 1) * 8)) || (rc = d_tm.alloc(n * 8)) ||
        (rc = d_dh.alloc(n * 8)) || (rc = d_ds.alloc(n * 8)) || (rc = d_st.alloc(n * 4)))
        return rc;
    if (nbytes) PG_CUDA(cudaMemcpyAsync(d_b.p, bases + b0, nbytes, cudaMemcpyHostToDevice, st));
    PG_CUDA(cudaMemcpyAsync(d_off.p, offsets, (n + 1) * 8, cudaMemcpyHostToDevice, st));
    PG_CUDA(cudaMemsetAsync(d_tm.p, 0, n * 8, st));
    PG_CUDA(cudaMemsetAsync(d_dh.p, 0, n * 8, st));
    PG_CUDA(cudaMemsetAsync(d_ds.p, 0, n * 8, st));
    rc = launch_tm(d_b.as<uint8_t>() - b0, d_off.as<uint64_t>(), n, cp, na, mg, d_tm.as<double>(), d_dh.as<double>(),
                   d_ds.as<double>(), d_st.as<int32_t>(), st);
    if (rc != PG_OK) return rc;
    std::vector<int32_t> hst(n);
    if (tm) PG_CUDA(cudaMemcpyAsync(tm, d_tm.p, n * 8, cudaMemcpyDeviceToHost, st));
    if (dh) PG_CUDA(cudaMemcpyAsync(dh, d_dh.p, n * 8, cudaMemcpyDeviceToHost, st));
    if (ds) PG_CUDA(cudaMemcpyAsync(ds, d_ds.p, n * 8, cudaMemcpyDeviceToHost, st));
    PG_CUDA(cudaMemcpyAsync(hst.data(), d_st.p, n * 4, cudaMemcpyDeviceToHost, st));
    PG_CUDA(cudaStreamSynchronize(st));
    int worst = PG_OK;
    for (uint64_t i = 0; i < n; ++i) {
        if (status) status[i] = hst[i];
        if (hst[i] == PG_ITEM_PANIC) worst = PG_ERR_PANIC;
        else if (hst[i] == PG_ITEM_UNSUPPORTED && worst == PG_OK) worst = PG_ERR_UNSUPPORTED;
    }
    if (worst == PG_ERR_PANIC) set_error("at least one primer is empty: primers.SantaLucia panics (index out of range)");
    if (worst == PG_ERR_UNSUPPORTED) set_error("at least one primer holds a byte >= 0x80 (unsupported: strings.ToUpper would re-encode it)");
    return worst;
}

int pg_design_primers_batch(const uint8_t *bases, const uint64_t *offsets, uint64_t n,
                            double target_tm, uint32_t *fwd_len, uint32_t *rev_len, int32_t *status) {
    int rc = ensure_device();
    if (rc != PG_OK) return rc;
    if (n == 0) return PG_OK;
    if (!offsets || !fwd_len || !rev_len) { set_error("null buffer"); return PG_ERR_ARG; }
    std::lock_guard<std::mutex> lk(g_mu);
    cudaStream_t st = g_streams[0];
    const uint64_t b0 = offsets[0], nbytes = offsets[n] - b0;
    Tmp d_b(st), d_off(st), d_f(st), d_r(st), d_st(st);
    if ((rc = d_b.alloc(nbytes)) || (rc = d_off.alloc((n + 1) * 8)) || (rc = d_f.alloc(n * 4)) ||
        (rc = d_r.alloc(n * 4)) || (rc = d_st.alloc(n * 4)))
        return rc;
    if (nbytes) PG_CUDA(cudaMemcpyAsync(d_b.p, bases + b0, nbytes, cudaMemcpyHostToDevice, st));
    PG_CUDA(cudaMemcpyAsync(d_off.p, offsets, (n + 1) * 8, cudaMemcpyHostToDevice, st));
    rc = launch_design_primers(d_b.as<uint8_t>() - b0, d_off.as<uint64_t>(), n, target_tm, d_f.as<uint32_t>(),
                               d_r.as<uint32_t>(), d_st.as<int32_t>(), st);
    if (rc != PG_OK) return rc;
    std::vector<int32_t> hst(n);
    PG_CUDA(cudaMemcpyAsync(fwd_len, d_f.p, n * 4, cudaMemcpyDeviceToHost, st));
    PG_CUDA(cudaMemcpyAsync(rev_len, d_r.p, n * 4, cudaMemcpyDeviceToHost, st));
    PG_CUDA(cudaMemcpyAsync(hst.data(), d_st.p, n * 4, cudaMemcpyDeviceToHost, st));
    PG_CUDA(cudaStreamSynchronize(st));
    int worst = PG_OK;
    for (uint64_t i = 0; i < n; ++i) {
        if (status) status[i] = hst[i];
        if (hst[i] == PG_ITEM_PANIC) worst = PG_ERR_PANIC;
        else if (hst[i] == PG_ITEM_UNSUPPORTED && worst == PG_OK) worst = PG_ERR_UNSUPPORTED;
    }
    if (worst == PG_ERR_PANIC) set_error("at least one sequence is exhausted before the target Tm: pcr.DesignPrimers panics (slice bounds out of range)");
    if (worst == PG_ERR_UNSUPPORTED) set_error("at least one sequence holds a byte >= 0x80 (unsupported)");
    return worst;
}

int pg_pcr_minimal_primer_batch(const uint8_t *bases, const uint64_t *offsets, uint64_t n, double target_tm,
                                uint32_t *min_len, int32_t *status) {
    int rc = ensure_device();
    if (rc != PG_OK) return rc;
    if (n == 0) return PG_OK;
    if (!offsets || !min_len) { set_error("null buffer"); return PG_ERR_ARG; }
    std::lock_guard<std::mutex> lk(g_mu);
    cudaStream_t st = g_streams[0];
    const uint64_t b0 = offsets[0], nbytes = offsets[n] - b0;
    Tmp d_b(st), d_off(st), d_m(st), d_st(st);
    if ((rc = d_b.alloc(nbytes)) || (rc = d_off.alloc((n + 1) * 8)) || (rc = d_m.alloc(n * 4)) || (rc = d_st.alloc(n * 4))) return rc;
    if (nbytes) PG_CUDA(cudaMemcpyAsync(d_b.p, bases + b0, nbytes, cudaMemcpyHostToDevice, st));
    PG_CUDA(cudaMemcpyAsync(d_off.p, offsets, (n + 1) * 8, cudaMemcpyHostToDevice, st));
    rc = launch_minimal_primer(d_b.as<uint8_t>() - b0, d_off.as<uint64_t>(), n, target_tm, d_m.as<uint32_t>(), d_st.as<int32_t>(), st);
    if (rc != PG_OK) return rc;
    std::vector<int32_t> hst(n);
    PG_CUDA(cudaMemcpyAsync(min_len, d_m.p, n * 4, cudaMemcpyDeviceToHost, st));
    PG_CUDA(cudaMemcpyAsync(hst.data(), d_st.p, n * 4, cudaMemcpyDeviceToHost, st));
    PG_CUDA(cudaStreamSynchronize(st));
    int worst = PG_OK;
    for (uint64_t i = 0; i < n; ++i) {
        if (status) status[i] = hst[i];
        if (hst[i] == PG_ITEM_PANIC) worst = PG_ERR_PANIC;
        else if (hst[i] == PG_ITEM_UNSUPPORTED && worst == PG_OK) worst = PG_ERR_UNSUPPORTED;
    }
    if (worst == PG_ERR_PANIC) set_error("at least one primer is shorter than 7 nt: pcr.SimulateSimple panics (slice bounds out of range)");
    if (worst == PG_ERR_UNSUPPORTED) set_error("at least one primer holds a byte >= 0x80 (unsupported)");
    return worst;
}

int pg_find_sites_batch(const uint8_t *seqs, const uint64_t *seq_offsets, uint64_t n_seq, const uint8_t *patterns,
                        const uint64_t *pat_offsets, uint32_t n_pat, uint32_t flags, uint32_t *hit_seq, uint64_t *hit_pos,
                        uint32_t *hit_pat, uint64_t hits_cap, uint64_t *n_hits) {
    int rc = ensure_device();
    if (rc != PG_OK) return rc;
    if (!n_hits) { set_error("null buffer"); return PG_ERR_ARG; }
    *n_hits = 0;
    if (n_seq == 0 || n_pat == 0) return PG_OK;
    if (!seq_offsets || !pat_offsets || (hits_cap && (!hit_seq || !hit_pos || !hit_pat))) { set_error("null buffer"); return PG_ERR_ARG; }
    std::lock_guard<std::mutex> lk(g_mu);
    cudaStream_t st = g_streams[0];
    const uint64_t s0 = seq_offsets[0], sbytes = seq_offsets[n_seq] - s0, p0 = pat_offsets[0], pbytes = pat_offsets[n_pat] - p0;
    std::vector<uint64_t> poff(n_pat + 1);
    for (uint32_t q = 0; q <= n_pat; ++q) poff[q] = pat_offsets[q] - p0;  // the kernel stages patterns from offset 0
    Tmp d_s(st), d_soff(st), d_p(st), d_poff(st), d_hs(st), d_hp(st), d_hq(st), d_n(st);
    if ((rc = d_s.alloc(sbytes)) || (rc = d_soff.alloc((n_seq + 1) * 8)) || (rc = d_p.alloc(pbytes)) || (rc = d_poff.alloc((n_pat + 1) * 8)) ||
        (rc = d_hs.alloc(hits_cap * 4)) || (rc = d_hp.alloc(hits_cap * 8)) || (rc = d_hq.alloc(hits_cap * 4)) || (rc = d_n.alloc(8)))
        return rc;
    if (sbytes) PG_CUDA(cudaMemcpyAsync(d_s.p, seqs + s0, sbytes, cudaMemcpyHostToDevice, st));
    PG_CUDA(cudaMemcpyAsync(d_soff.p, seq_offsets, (n_seq + 1) * 8, cudaMemcpyHostToDevice, st));
    if (pbytes) PG_CUDA(cudaMemcpyAsync(d_p.p, patterns + p0, pbytes, cudaMemcpyHostToDevice, st));
    PG_CUDA(cudaMemcpyAsync(d_poff.p, poff.data(), (n_pat + 1) * 8, cudaMemcpyHostToDevice, st));
    PG_CUDA(cudaMemsetAsync(d_n.p, 0, 8, st));
    rc = launch_find_sites(d_s.as<uint8_t>() - s0, d_soff.as<uint64_t>(), n_seq, sbytes, d_p.as<uint8_t>(), d_poff.as<uint64_t>(), n_pat, pbytes,
                           flags, d_hs.as<uint32_t>(), d_hp.as<uint64_t>(), d_hq.as<uint32_t>(), hits_cap, d_n.as<unsigned long long>(), st);
    if (rc != PG_OK) return rc;
    unsigned long long found = 0;
    PG_CUDA(cudaMemcpyAsync(&found, d_n.p, 8, cudaMemcpyDeviceToHost, st));
    PG_CUDA(cudaStreamSynchronize(st));
    *n_hits = found;
    const uint64_t kept = std::min<uint64_t>(found, hits_cap);
    if (kept) {
        PG_CUDA(cudaMemcpyAsync(hit_seq, d_hs.p, kept * 4, cudaMemcpyDeviceToHost, st));
        PG_CUDA(cudaMemcpyAsync(hit_pos, d_hp.p, kept * 8, cudaMemcpyDeviceToHost, st));
        PG_CUDA(cudaMemcpyAsync(hit_pat, d_hq.p, kept * 4, cudaMemcpyDeviceToHost, st));
        PG_CUDA(cudaStreamSynchronize(st));
    }
    if (found > hits_cap) {
        set_error("hits_cap %llu < %llu occurrences", (unsigned long long)hits_cap, found);
        return PG_ERR_ARG;
    }
    return PG_OK;
}

int pg_fastq_ingest_dev(const uint8_t *d_text, uint64_t nbytes, uint8_t *d_bases, uint64_t bases_cap,
                        uint64_t *d_offsets, uint64_t records_cap, uint64_t *n_records,
                        uint64_t *total_bases, int32_t *err_code, uint64_t *err_line, void *stream) {
    int rc = ensure_device();
    if (rc != PG_OK) return rc;
    if (!n_records || !total_bases || !err_code || !err_line || !d_offsets) { set_error("null buffer"); return PG_ERR_ARG; }
    return launch_fastq_ingest(d_text, nbytes, d_bases, bases_cap, d_offsets, records_cap, n_records, total_bases, err_code,
                               err_line, (cudaStream_t)stream);
}

int pg_fastq_ingest(const uint8_t *text, uint64_t nbytes, uint8_t *bases, uint64_t bases_cap,
                    uint64_t *offsets, uint64_t records_cap, uint64_t *n_records,
                    uint64_t *total_bases, int32_t *err_code, uint64_t *err_line) {
    int rc = ensure_device();
    if (rc != PG_OK) return rc;
    if (!n_records || !total_bases || !err_code || !err_line || !offsets) { set_error("null buffer"); return PG_ERR_ARG; }
    std::lock_guard<std::mutex> lk(g_mu);
    cudaStream_t st = g_streams[0];
    Tmp d_text(st), d_bases(st), d_off(st);
    if ((rc = d_text.alloc(nbytes + 16)) || (rc = d_bases.alloc(bases_cap + 16)) || (rc = d_off.alloc((records_cap + 1) * 8)))
        return rc;
    if (nbytes) PG_CUDA(cudaMemcpyAsync(d_text.p, text, nbytes, cudaMemcpyHostToDevice, st));
    rc = launch_fastq_ingest(d_text.as<uint8_t>(), nbytes, d_bases.as<uint8_t>(), bases_cap, d_off.as<uint64_t>(), records_cap,
                             n_records, total_bases, err_code, err_line, st);
    if (rc != PG_OK) return rc;
    if (*total_bases) PG_CUDA(cudaMemcpyAsync(bases, d_bases.p, *total_bases, cudaMemcpyDeviceToHost, st));
    PG_CUDA(cudaMemcpyAsync(offsets, d_off.p, (*n_records + 1) * 8, cudaMemcpyDeviceToHost, st));
    PG_CUDA(cudaStreamSynchronize(st));
    return PG_OK;
}

int pg_fasta_ingest_dev(const uint8_t *d_text, uint64_t nbytes, uint32_t max_line_size, uint32_t flags,
                        uint8_t *d_bases, uint64_t bases_cap, uint64_t *d_offsets, uint8_t *d_names,
                        uint64_t names_cap, uint64_t *d_name_offsets, uint64_t records_cap, uint64_t *n_records,
                        uint64_t *total_bases, uint64_t *total_name_bytes, int32_t *err_code, uint64_t *err_line,
                        void *stream) {
    int rc = ensure_device();
    if (rc != PG_OK) return rc;
    if (!n_records || !total_bases || !total_name_bytes || !err_code || !err_line || !d_offsets) { set_error("null buffer"); return PG_ERR_ARG; }
    return launch_fasta_ingest(d_text, nbytes, max_line_size, flags, d_bases, bases_cap, d_offsets, d_names, names_cap, d_name_offsets,
                               records_cap, n_records, total_bases, total_name_bytes, err_code, err_line, (cudaStream_t)stream);
}

int pg_fasta_ingest(const uint8_t *text, uint64_t nbytes, uint32_t max_line_size, uint32_t flags, uint8_t *bases,
                    uint64_t bases_cap, uint64_t *offsets, uint8_t *names, uint64_t names_cap, uint64_t *name_offsets,
                    uint64_t records_cap, uint64_t *n_records, uint64_t *total_bases, uint64_t *total_name_bytes,
                    int32_t *err_code, uint64_t *err_line) {
    int rc = ensure_device();
    if (rc != PG_OK) return rc;
    if (!n_records || !total_bases || !total_name_bytes || !err_code || !err_line || !offsets) { set_error("null buffer"); return PG_ERR_ARG; }
    const bool want_names = names != nullptr || name_offsets != nullptr;
    if (want_names && (!names || !name_offsets)) { set_error("names and name_offsets must be given together"); return PG_ERR_ARG; }
    std::lock_guard<std::mutex> lk(g_mu);
    cudaStream_t st = g_streams[0];
    Tmp d_text(st), d_bases(st), d_off(st), d_names(st), d_noff(st);
    if ((rc = d_text.alloc(nbytes + 16)) || (rc = d_bases.alloc(bases_cap + 16)) || (rc = d_off.alloc((records_cap + 1) * 8)))
        return rc;
    if (want_names && ((rc = d_names.alloc(names_cap + 16)) || (rc = d_noff.alloc((records_cap + 1) * 8)))) return rc;
    if (nbytes) PG_CUDA(cudaMemcpyAsync(d_text.p, text, nbytes, cudaMemcpyHostToDevice, st));
    rc = launch_fasta_ingest(d_text.as<uint8_t>(), nbytes, max_line_size, flags, d_bases.as<uint8_t>(), bases_cap, d_off.as<uint64_t>(),
                             want_names ? d_names.as<uint8_t>() : nullptr, names_cap, want_names ? d_noff.as<uint64_t>() : nullptr,
                             records_cap, n_records, total_bases, total_name_bytes, err_code, err_line, st);
    if (rc != PG_OK) return rc;
    if (*total_bases) PG_CUDA(cudaMemcpyAsync(bases, d_bases.p, *total_bases, cudaMemcpyDeviceToHost, st));
    PG_CUDA(cudaMemcpyAsync(offsets, d_off.p, (*n_records + 1) * 8, cudaMemcpyDeviceToHost, st));
    if (want_names) {
        if (*total_name_bytes) PG_CUDA(cudaMemcpyAsync(names, d_names.p, *total_name_bytes, cudaMemcpyDeviceToHost, st));
        PG_CUDA(cudaMemcpyAsync(name_offsets, d_noff.p, (*n_records + 1) * 8, cudaMemcpyDeviceToHost, st));
    }
    PG_CUDA(cudaStreamSynchronize(st));
    return PG_OK;
}

// ---------------------------------------------------------------------------------
int pg_synth_reads_dev(uint8_t *d_bases, uint64_t first_read, uint64_t n_reads, uint32_t read_len,
                       uint64_t seed, int32_t kind, uint32_t family, void *stream) {
    int rc = ensure_device();
    if (rc != PG_OK) return rc;
    return launch_synth_reads(d_bases, first_read, n_reads, read_len, seed, kind, family, (cudaStream_t)stream);
}

}  // extern "C"
