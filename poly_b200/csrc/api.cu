// api.cu -- the C ABI of libpolyb200.so (include/poly_b200.h): argument checking,
// device selection, host<->device staging for the host-pointer entry points, and
// dispatch to the kernel launchers.  No compute happens on the host: when no sm_100
// device is usable every compute entry point fails with PG_ERR_NO_DEVICE.
#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <functional>
#include <cctype>
#include <sched.h>
#include <sys/syscall.h>
#include <unistd.h>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <string>
#include <vector>

#include "common.cuh"

namespace pg {

static thread_local char t_err[512] = "";
static thread_local const char *t_last_kernel = "";
static std::atomic<uint64_t> g_launches{0};

// grow-only device scratch used by the host-pointer entry points (under the context's mutex)
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

// grow-only pinned host staging (pageable caller buffers go through these; under the context's mutex)
struct PinnedScratch {
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return PG_OK;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocPortable);
        if (e != cudaSuccess) {
            set_error("cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
            cudaGetLastError();
            return PG_ERR_NOMEM;
        }
        cap = bytes;
        return PG_OK;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

// Fork-join helper threads for host-side copies between pageable caller memory and the pinned
// staging buffers: one memcpy thread moves ~10 GB/s, PCIe 5 x16 moves 55.
class CopyPool {
    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable cv_work_, cv_done_;
    const std::function<void(int)> *fn_ = nullptr;
    int parts_ = 0, next_ = 0, pending_ = 0;
    bool stop_ = false;

    void drain_parts(std::unique_lock<std::mutex> &lk) {
        while (fn_ && next_ < parts_) {
            const int i = next_++;
            const std::function<void(int)> *f = fn_;
            lk.unlock();
            (*f)(i);
            lk.lock();
            if (--pending_ == 0) cv_done_.notify_all();
        }
    }

public:
    CopyPool() = default;
    CopyPool(const CopyPool &) = delete;
    CopyPool &operator=(const CopyPool &) = delete;
    ~CopyPool() { stop(); }  // contexts are statics: joinable threads at exit would call std::terminate
    int threads() const { return (int)workers_.size() + 1; }
    void start(int helpers, int device) {
        if (!workers_.empty() || helpers <= 0) return;
        for (int t = 0; t < helpers; ++t)
            workers_.emplace_back([this, device] {
                pg_numa_bind_thread(device, nullptr);  // copies run next to the GPU's PCIe root
                std::unique_lock<std::mutex> lk(mu_);
                for (;;) {
                    cv_work_.wait(lk, [&] { return stop_ || (fn_ && next_ < parts_); });
                    if (stop_) return;
                    drain_parts(lk);
                }
            });
    }
    // f(0) .. f(parts-1), each exactly once, on the helpers and the calling thread; returns when all are done
    void run(int parts, const std::function<void(int)> &f) {
        if (parts <= 0) return;
        std::unique_lock<std::mutex> lk(mu_);
        fn_ = &f; parts_ = parts; next_ = 0; pending_ = parts;
        cv_work_.notify_all();
        drain_parts(lk);
        cv_done_.wait(lk, [&] { return pending_ == 0; });
        fn_ = nullptr;
    }
    void stop() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_work_.notify_all();
        for (std::thread &t : workers_) t.join();
        workers_.clear();
        stop_ = false;
    }
};

// One context per CUDA device of the process.  The library serves any number of devices from one
// process (pg_*_multi shard a batch over them; pg_thread_device binds a thread to one) as well as
// the one-process-per-GPU layout (pg_init(device) picks the process default).
constexpr int MAX_DEVICES = 32;
constexpr int N_SLOTS = 3;  // stream/buffer slots of the pipelined host path
struct DevCtx {
    int device = -1;
    int sms = 0;
    std::atomic<bool> ready{false};
    cudaStream_t streams[N_SLOTS] = {nullptr, nullptr, nullptr};
    Scratch in[N_SLOTS], out[N_SLOTS], aux[N_SLOTS], st[N_SLOTS];
    PinnedScratch pin_in[N_SLOTS], pin_out[N_SLOTS];
    CopyPool pool;
    std::mutex mu;   // serialises the host-pointer entry points on this device
    std::mutex fmu;  // guards func_smem
    std::unordered_map<const void *, size_t> func_smem;  // largest dynamic smem configured per kernel
};
static DevCtx g_ctx[MAX_DEVICES];
static std::mutex g_init_mu;                 // context creation / default-device changes
static std::atomic<int> g_default_device{-1};  // process default (pg_init), -1: not chosen yet
static thread_local int t_bound_device = -1;   // pg_thread_device override for this thread
static thread_local DevCtx *t_ctx = nullptr;   // context the current entry point runs on

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
int sm_count() { return t_ctx && t_ctx->sms > 0 ? t_ctx->sms : 148; }

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (device, kernel, size class): function
// attributes live in the device's context, so the cache is per device and lock-protected (several
// host threads may drive the same device through the *_dev entry points).
int func_smem(const void *fn, size_t bytes) {
    DevCtx *c = t_ctx;
    if (!c) { set_error("internal: no device context bound"); return PG_ERR_ARG; }
    std::lock_guard<std::mutex> lk(c->fmu);
    size_t &cur = c->func_smem[fn];
    if (bytes > cur) {
        PG_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        cur = bytes;
    }
    return PG_OK;
}

// create (once) the context of `device`; g_init_mu held
static int ctx_create_locked(int device) {
    DevCtx &c = g_ctx[device];
    if (c.ready.load(std::memory_order_acquire)) return PG_OK;
    cudaDeviceProp prop;
    PG_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("device %d is sm_%d%d; libpolyb200 is built for sm_100a only", device, prop.major, prop.minor);
        return PG_ERR_NO_DEVICE;
    }
    PG_CUDA(cudaSetDevice(device));
    for (int i = 0; i < N_SLOTS; ++i)
        if (!c.streams[i]) PG_CUDA(cudaStreamCreateWithFlags(&c.streams[i], cudaStreamNonBlocking));
    {   // Stream-ordered temporaries (cudaMallocAsync) are freed at the end of every call and most calls
        // synchronise: with the default release threshold of 0 the pool would hand its memory back to the
        // driver each time.  Keep up to 12 GiB cached (the 8 GiB pair tiles of the sparse distance path included) so that small calls do not pay a driver allocation.
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t keep = 12ull << 30, cur = 0;
            if (cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &cur) == cudaSuccess && cur < keep)
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    c.device = device;
    c.sms = prop.multiProcessorCount;
    c.ready.store(true, std::memory_order_release);
    return PG_OK;
}

static int device_count_checked(int *n) {
    cudaError_t e = cudaGetDeviceCount(n);
    if (e != cudaSuccess || *n == 0) {
        set_error("no CUDA device available (%s); libpolyb200 has no CPU fallback",
                  e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        cudaGetLastError();
        return PG_ERR_NO_DEVICE;
    }
    if (*n > MAX_DEVICES) *n = MAX_DEVICES;
    return PG_OK;
}

// make `device` usable and current on the calling thread; t_ctx = its context
static int use_device(int device) {
    int n = 0, rc = device_count_checked(&n);
    if (rc != PG_OK) return rc;
    if (device < 0 || device >= n) {
        set_error("device %d out of range (count %d)", device, n);
        return PG_ERR_ARG;
    }
    if (!g_ctx[device].ready.load(std::memory_order_acquire)) {
        std::lock_guard<std::mutex> lk(g_init_mu);
        if ((rc = ctx_create_locked(device)) != PG_OK) return rc;
    }
    PG_CUDA(cudaSetDevice(device));
    t_ctx = &g_ctx[device];
    return PG_OK;
}

// every entry point: the thread's bound device (pg_thread_device), else the process default
// (pg_init), else the current CUDA device / device 0
static int ensure_device() {
    int device = t_bound_device;
    if (device < 0) device = g_default_device.load(std::memory_order_acquire);
    if (device < 0) {
        int n = 0, rc = device_count_checked(&n);
        if (rc != PG_OK) return rc;
        std::lock_guard<std::mutex> lk(g_init_mu);
        device = g_default_device.load();
        if (device < 0) {
            if (cudaGetDevice(&device) != cudaSuccess || device >= n) device = 0;
            if ((rc = ctx_create_locked(device)) != PG_OK) return rc;
            g_default_device.store(device, std::memory_order_release);
        }
    }
    return use_device(device);
}

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
    if (device < 0) {
        g_default_device.store(-1);
        return ensure_device();
    }
    int rc = use_device(device);
    if (rc == PG_OK) g_default_device.store(device, std::memory_order_release);
    return rc;
}

int pg_thread_device(int device) {
    if (device < 0) { t_bound_device = -1; return PG_OK; }
    int rc = use_device(device);
    if (rc == PG_OK) t_bound_device = device;
    return rc;
}

int pg_shutdown(void) {
    std::lock_guard<std::mutex> lk(g_init_mu);
    for (int d = 0; d < MAX_DEVICES; ++d) {
        DevCtx &c = g_ctx[d];
        if (!c.ready.load()) continue;
        std::lock_guard<std::mutex> lk2(c.mu);
        cudaSetDevice(d);
        cudaDeviceSynchronize();
        for (int i = 0; i < N_SLOTS; ++i) {
            c.in[i].release(); c.out[i].release(); c.aux[i].release(); c.st[i].release();
            c.pin_in[i].release(); c.pin_out[i].release();
            if (c.streams[i]) { cudaStreamDestroy(c.streams[i]); c.streams[i] = nullptr; }
        }
        c.pool.stop();
        {
            std::lock_guard<std::mutex> lk3(c.fmu);
            c.func_smem.clear();
        }
        c.ready.store(false);
    }
    g_default_device.store(-1);
    t_ctx = nullptr;
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
    *sms = t_ctx->sms;
    return PG_OK;
}

// Put the calling thread on the CPUs of the NUMA node `device` hangs off, and prefer that node for
// the pages it allocates from now on (pinned buffers included): PCIe traffic of 8 GPUs otherwise
// funnels through one socket's memory controllers / the inter-socket link.  Best effort: every
// failure (no sysfs, container without the syscall, single-node host) leaves the thread as it was.
int pg_numa_bind_thread(int device, int *node_out) {
    if (node_out) *node_out = -1;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) { cudaGetLastError(); set_error("device %d out of range", device); return PG_ERR_ARG; }
    char bus[64] = "";
    if (cudaDeviceGetPCIBusId(bus, (int)sizeof bus, device) != cudaSuccess) { cudaGetLastError(); return PG_OK; }
    for (char *c = bus; *c; ++c) *c = (char)tolower((unsigned char)*c);
    char path[256];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
    int node = -1;
    if (FILE *f = fopen(path, "r")) { if (fscanf(f, "%d", &node) != 1) node = -1; fclose(f); }
    if (node < 0) return PG_OK;
    snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
    cpu_set_t want, cur, both;
    CPU_ZERO(&want);
    if (FILE *f = fopen(path, "r")) {
        int a, b;
        for (;;) {
            if (fscanf(f, "%d", &a) != 1) break;
            b = a;
            int ch = fgetc(f);
            if (ch == '-') { if (fscanf(f, "%d", &b) != 1) break; ch = fgetc(f); }
            for (int c = a; c <= b && c < CPU_SETSIZE; ++c) CPU_SET(c, &want);
            if (ch != ',') break;
        }
        fclose(f);
    }
    if (sched_getaffinity(0, sizeof cur, &cur) == 0) {
        CPU_AND(&both, &cur, &want);
        if (CPU_COUNT(&both) > 0) sched_setaffinity(0, sizeof both, &both);
    }
#ifdef SYS_set_mempolicy
    if (node < 1024) {
        unsigned long mask[16] = {0};
        mask[node / (8 * sizeof(unsigned long))] |= 1ul << (node % (8 * sizeof(unsigned long)));
        (void)syscall(SYS_set_mempolicy, 1 /* MPOL_PREFERRED */, mask, (unsigned long)(sizeof mask * 8));
    }
#endif
    if (node_out) *node_out = node;
    return PG_OK;
}

int pg_host_alloc(void **ptr, size_t bytes) {
    if (!ptr) return PG_ERR_ARG;
    int rc = ensure_device();
    if (rc != PG_OK) return rc;
    // portable: pinned for every device context of the process (the *_multi entry points)
    PG_CUDA(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocPortable));
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
static int check_dev_flags(uint32_t flags) {
    if (flags & PG_SKETCH_TAIL_KEEP) { set_error("PG_SKETCH_TAIL_KEEP applies to host buffers only"); return PG_ERR_ARG; }
    return PG_OK;
}

int pg_mash_sketch_uniform_dev(const uint8_t *d_bases, uint64_t n_reads, uint32_t read_len,
                               int32_t k, int32_t s, uint32_t flags, uint32_t *d_out,
                               uint64_t row_stride, int32_t *d_status, void *stream) {
    int rc = ensure_device();
    if (rc != PG_OK) return rc;
    if ((rc = check_ks(k, s)) != PG_OK) return rc;
    if ((rc = check_dev_flags(flags)) != PG_OK) return rc;
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
    if ((rc = check_dev_flags(flags)) != PG_OK) return rc;
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

// rows [row_offset, row_offset + n_local) of every destination buffer (compact rows of cnt words)
static int sketch_scatter_dev(const uint8_t *d_bases, uint64_t n_local, uint32_t read_len, int32_t k, int32_t s,
                              void *const *dst_ptrs, int32_t n_dst, int32_t self, uint64_t row_offset, cudaStream_t st) {
    int rc;
    if ((rc = check_ks(k, s)) != PG_OK) return rc;
    if (n_dst < 1 || n_dst > PG_MAX_PEERS || self < 0 || self >= n_dst || !dst_ptrs) {
        set_error("bad destination set (%d buffers, own index %d; at most %d)", n_dst, self, PG_MAX_PEERS);
        return PG_ERR_ARG;
    }
    const uint64_t cnt = std::min<uint64_t>(kmers_of(read_len, k), (uint64_t)s);
    SketchDst dst;
    dst.n = n_dst;
    for (int p = 0; p < n_dst; ++p) {
        if (!dst_ptrs[p]) { set_error("null gathered pointer for rank %d", p); return PG_ERR_ARG; }
        dst.ptr[p] = (uint32_t *)dst_ptrs[p] + row_offset * cnt;  // this rank's row block
    }
    return launch_sketch_uniform(d_bases, n_local, read_len, k, s, 0, dst.ptr[self], cnt, nullptr, st, &dst);
}

int pg_mash_sketch_uniform_gather_dev(const uint8_t *d_bases, uint64_t n_local, uint32_t read_len,
                                      int32_t k, int32_t s, void *const *gathered_ptrs,
                                      int32_t world, int32_t rank, void *stream) {
    int rc = ensure_device();
    if (rc != PG_OK) return rc;
    return sketch_scatter_dev(d_bases, n_local, read_len, k, s, gathered_ptrs, world, rank, (uint64_t)(rank < 0 ? 0 : rank) * n_local,
                              (cudaStream_t)stream);
}

int pg_mash_sketch_uniform_scatter_dev(const uint8_t *d_bases, uint64_t n_local, uint32_t read_len, int32_t k, int32_t s,
                                       void *const *dst_ptrs, int32_t n_dst, int32_t self, uint64_t row_offset, void *stream) {
    int rc = ensure_device();
    if (rc != PG_OK) return rc;
    return sketch_scatter_dev(d_bases, n_local, read_len, k, s, dst_ptrs, n_dst, self, row_offset, (cudaStream_t)stream);
}

// Pipelined host path shared by the uniform and ragged entry points: chunks of reads
// cycle through 3 stream/buffer slots (H2D, kernel, D2H overlap across slots).
static bool is_pageable(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
}

static int host_copy_helpers() {
    if (const char *e = getenv("PG_HOST_COPY_THREADS")) return std::max(0, atoi(e) - 1);
    cpu_set_t cur;
    int cores = (int)std::thread::hardware_concurrency();
    if (sched_getaffinity(0, sizeof cur, &cur) == 0) cores = std::min(cores, CPU_COUNT(&cur));
    return std::max(0, std::min(8, cores / 2) - 1);
}

static int sketch_host(const uint8_t *bases, const uint64_t *offsets, uint32_t ulen, uint64_t n_reads,
                       int32_t k, int32_t s, uint32_t flags, uint32_t *out, uint64_t row_stride,
                       uint32_t *count, int32_t *status) {
    int rc = ensure_device();
    if (rc != PG_OK) return rc;
    if ((rc = check_ks(k, s)) != PG_OK) return rc;
    if (n_reads == 0) return PG_OK;
    if (!bases || !out) { set_error("null buffer"); return PG_ERR_ARG; }
    DevCtx *cx = t_ctx;
    std::lock_guard<std::mutex> lk(cx->mu);
    // Pinned caller buffers (pg_host_alloc, cudaHostRegister) are DMA'd directly.  Pageable ones (a Go
    // []byte / []uint32) would be bounced by the driver at ~14 GB/s: they go through pinned staging
    // buffers instead, filled / emptied by a few host threads while the other slots' DMA runs; the
    // device then always produces compact rows and the zero tail of PG_SKETCH_PAD_ZERO rows is
    // written by the host threads, not shipped over PCIe.
    // PG_SKETCH_TAIL_KEEP: only the informative words of a row are written.  Ragged batches then always go through
    // the staging buffers, whose rows are copied out count[i] words at a time.
    const bool keep_tail = (flags & PG_SKETCH_TAIL_KEEP) != 0;
    flags &= ~PG_SKETCH_TAIL_KEEP;
    const bool stage_in = is_pageable(bases), stage_out = is_pageable(out) || (keep_tail && offsets != nullptr);
    if (stage_in || stage_out) cx->pool.start(host_copy_helpers(), cx->device);
    const int nthr = cx->pool.threads();
    const bool want_status = status != nullptr || s <= 1;
    const uint64_t target_bytes = (stage_in || stage_out) ? (32ull << 20) : (192ull << 20);  // input bytes per chunk
    const uint64_t target_out_bytes = 128ull << 20;                                          // staged output bytes per chunk
    bool any_panic = false;
    uint64_t r0 = 0;
    int slot = 0;
    struct Pending { uint64_t r0, nr, dstride; bool used; } pend[N_SLOTS] = {};
    std::vector<int32_t> st_slot[N_SLOTS];
    auto par_rows = [&](uint64_t nrows, const std::function<void(uint64_t, uint64_t)> &f) {
        const int parts = (int)std::min<uint64_t>((uint64_t)nthr, std::max<uint64_t>(1, nrows));
        cx->pool.run(parts, [&](int i) { f(nrows * (uint64_t)i / parts, nrows * (uint64_t)(i + 1) / parts); });
    };
    auto drain = [&](int sl) -> int {
        if (!pend[sl].used) return PG_OK;
        PG_CUDA(cudaStreamSynchronize(cx->streams[sl]));
        const Pending &pd = pend[sl];
        if (stage_out && pd.dstride) {  // staged rows -> caller rows (+ zero tail up to the caller's stride)
            const uint32_t *src = (const uint32_t *)cx->pin_out[sl].p;
            par_rows(pd.nr, [&](uint64_t a, uint64_t b) {
                if (pd.dstride == row_stride && !(keep_tail && offsets)) {
                    memcpy(out + (pd.r0 + a) * row_stride, src + a * pd.dstride, (b - a) * pd.dstride * 4);
                    return;
                }
                for (uint64_t i = a; i < b; ++i) {
                    uint32_t *dst = out + (pd.r0 + i) * row_stride;
                    if (keep_tail) {  // exactly the informative words of this row
                        const uint64_t r = pd.r0 + i;
                        const uint64_t ci = offsets ? std::min<uint64_t>(kmers_of(offsets[r + 1] - offsets[r], k), (uint64_t)s) : pd.dstride;
                        memcpy(dst, src + i * pd.dstride, ci * 4);
                        continue;
                    }
                    memcpy(dst, src + i * pd.dstride, pd.dstride * 4);
                    memset(dst + pd.dstride, 0, (row_stride - pd.dstride) * 4);
                }
            });
        } else if (row_stride > pd.dstride && !keep_tail) {  // direct DMA wrote dstride words per row: zero the rest of each row
            par_rows(pd.nr, [&](uint64_t a, uint64_t b) {
                for (uint64_t i = a; i < b; ++i) memset(out + (pd.r0 + i) * row_stride + pd.dstride, 0, (row_stride - pd.dstride) * 4);
            });
        }
        if (want_status) {
            for (uint64_t i = 0; i < pd.nr; ++i) {
                if (st_slot[sl][i] != PG_ITEM_OK) any_panic = true;
                if (status) status[pd.r0 + i] = st_slot[sl][i];
            }
        }
        pend[sl].used = false;
        return PG_OK;
    };
    auto body = [&]() -> int {
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
                    if (stage_out && (r1 - r0) * std::min<uint64_t>(kmers_of(maxlen, k), (uint64_t)s) * 4 >= target_out_bytes) break;
                }
                end = offsets[r1];
            } else {
                uint64_t per = std::max<uint64_t>(1, target_bytes / std::max<uint32_t>(ulen, 1));
                const uint64_t row_bytes = std::min<uint64_t>(kmers_of(ulen, k), (uint64_t)s) * 4;
                if (stage_out && row_bytes) per = std::min(per, std::max<uint64_t>(32, target_out_bytes / row_bytes));
                per = (per + 31) & ~31ull;  // keep chunk starts on K1 tile boundaries
                r1 = std::min(n_reads, r0 + per);
                beg = r0 * (uint64_t)ulen;
                end = r1 * (uint64_t)ulen;
                maxlen = minlen = ulen;
            }
            const uint64_t nr = r1 - r0;
            const bool uniform = maxlen == minlen && maxlen <= 0xffffffffull;
            const uint64_t cnt_max = std::min<uint64_t>(kmers_of(maxlen, k), (uint64_t)s);
            const uint64_t need_stride = (flags & PG_SKETCH_PAD_ZERO) ? (uint64_t)s : cnt_max;
            if (row_stride < need_stride) { set_error("row_stride %llu < %llu", (unsigned long long)row_stride, (unsigned long long)need_stride); return PG_ERR_ARG; }
            // staged output: compact device rows, the host threads pad; direct output: the device pads
            const uint32_t dev_flags = (stage_out || keep_tail) ? (flags & ~PG_SKETCH_PAD_ZERO) : flags;
            const uint64_t dev_stride = (stage_out || keep_tail) ? cnt_max : need_stride;

            int rc2;
            if ((rc2 = drain(slot)) != PG_OK) return rc2;
            cudaStream_t stx = cx->streams[slot];
            const uint64_t in_bytes = end - beg;
            if ((rc2 = cx->in[slot].reserve(in_bytes + 64)) != PG_OK) return rc2;
            if ((rc2 = cx->out[slot].reserve(std::max<uint64_t>(nr * dev_stride * 4, 16))) != PG_OK) return rc2;
            uint8_t *d_in = (uint8_t *)cx->in[slot].p;
            uint32_t *d_out = (uint32_t *)cx->out[slot].p;
            int32_t *d_status = nullptr;
            uint32_t *d_count = nullptr;
            uint64_t *d_off = nullptr;
            if (want_status) {
                if ((rc2 = cx->st[slot].reserve(nr * 4)) != PG_OK) return rc2;
                d_status = (int32_t *)cx->st[slot].p;
                PG_CUDA(cudaMemsetAsync(d_status, 0, nr * 4, stx));
                st_slot[slot].assign(nr, 0);
            }
            const uint8_t *h_src = bases + beg;
            if (stage_in && in_bytes) {
                if ((rc2 = cx->pin_in[slot].reserve(in_bytes)) != PG_OK) return rc2;
                uint8_t *pin = (uint8_t *)cx->pin_in[slot].p;
                par_rows((in_bytes + (1u << 20) - 1) >> 20, [&](uint64_t a, uint64_t b) {  // 1 MiB granules
                    const uint64_t lo = a << 20, hi = std::min<uint64_t>(in_bytes, b << 20);
                    if (hi > lo) memcpy(pin + lo, h_src + lo, hi - lo);
                });
                h_src = pin;
            }
            if (in_bytes) PG_CUDA(cudaMemcpyAsync(d_in, h_src, in_bytes, cudaMemcpyHostToDevice, stx));
            if (uniform) {
                rc2 = launch_sketch_uniform(d_in, nr, (uint32_t)maxlen, k, s, dev_flags, d_out, dev_stride, d_status, stx);
            } else {
                if ((rc2 = cx->aux[slot].reserve((nr + 1) * 8 + nr * 4)) != PG_OK) return rc2;
                d_off = (uint64_t *)cx->aux[slot].p;
                d_count = (uint32_t *)(d_off + nr + 1);
                PG_CUDA(cudaMemcpyAsync(d_off, offsets + r0, (nr + 1) * 8, cudaMemcpyHostToDevice, stx));
                // kernels index bases with absolute offsets: shift the base pointer
                rc2 = launch_sketch_ragged(d_in - beg, d_off, nr, maxlen, k, s, dev_flags, d_out, dev_stride, d_count, d_status, stx);
            }
            if (rc2 != PG_OK) return rc2;
            if (dev_stride) {
                if (stage_out) {
                    if ((rc2 = cx->pin_out[slot].reserve(nr * dev_stride * 4)) != PG_OK) return rc2;
                    PG_CUDA(cudaMemcpyAsync(cx->pin_out[slot].p, d_out, nr * dev_stride * 4, cudaMemcpyDeviceToHost, stx));
                } else if (row_stride == dev_stride) {
                    PG_CUDA(cudaMemcpyAsync(out + r0 * row_stride, d_out, nr * dev_stride * 4, cudaMemcpyDeviceToHost, stx));
                } else {
                    PG_CUDA(cudaMemcpy2DAsync(out + r0 * row_stride, row_stride * 4, d_out, dev_stride * 4, dev_stride * 4, nr, cudaMemcpyDeviceToHost, stx));
                }
            }
            if (count) {
                if (uniform) {
                    for (uint64_t i = 0; i < nr; ++i) count[r0 + i] = (uint32_t)cnt_max;
                } else {
                    PG_CUDA(cudaMemcpyAsync(count + r0, d_count, nr * 4, cudaMemcpyDeviceToHost, stx));
                }
            }
            if (want_status) PG_CUDA(cudaMemcpyAsync(st_slot[slot].data(), d_status, nr * 4, cudaMemcpyDeviceToHost, stx));
            pend[slot] = {r0, nr, dev_stride, true};
            slot = (slot + 1) % N_SLOTS;
            r0 = r1;
        }
        for (int i = 0; i < N_SLOTS; ++i) {
            const int rc2 = drain((slot + i) % N_SLOTS);  // oldest first
            if (rc2 != PG_OK) return rc2;
        }
        return PG_OK;
    };
    rc = body();
    if (rc != PG_OK) {  // nothing may still be copying into the caller's (or this frame's) buffers when we return
        for (int i = 0; i < N_SLOTS; ++i) cudaStreamSynchronize(cx->streams[i]);
        cudaGetLastError();
        return rc;
    }
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
// Single-process multi-GPU entry points (SURVEY.md 8b "multi-GPU variants taking a device
// count", 8e): one host thread per device, reads sharded in contiguous blocks, no data-path
// collective for Sketch; Sketch+Distance exchanges the finished sketches by peer stores issued
// from the sketch kernels themselves (in-process peer access, no IPC).
// ---------------------------------------------------------------------------------
namespace pg {
namespace {

struct WorkerResult {
    int rc = PG_OK;
    std::string msg;
};

// worst result of the workers, message copied to the caller's thread-local slot.  A reference panic
// (PG_ERR_PANIC: per-item statuses are still complete) ranks below real failures.
static int fold_results(const std::vector<WorkerResult> &res) {
    int worst = PG_OK;
    const std::string *msg = nullptr;
    for (const WorkerResult &r : res) {
        if (r.rc == PG_OK) continue;
        const bool soft = r.rc == PG_ERR_PANIC, worst_soft = worst == PG_ERR_PANIC || worst == PG_OK;
        if (worst == PG_OK || (!soft && worst_soft)) { worst = r.rc; msg = &r.msg; }
    }
    if (msg) set_error("%s", msg->c_str());
    return worst;
}

static int resolve_devices(const int32_t *devices, int32_t n_devices, std::vector<int> *out) {
    int n = 0, rc = device_count_checked(&n);
    if (rc != PG_OK) return rc;
    out->clear();
    if (!devices) {
        const int cnt = n_devices > 0 ? std::min<int>(n_devices, n) : n;
        for (int i = 0; i < cnt; ++i) out->push_back(i);
    } else {
        if (n_devices <= 0) { set_error("device list given with n_devices = %d", n_devices); return PG_ERR_ARG; }
        for (int i = 0; i < n_devices; ++i) {
            if (devices[i] < 0 || devices[i] >= n) { set_error("device %d out of range (count %d)", devices[i], n); return PG_ERR_ARG; }
            for (int j = 0; j < i; ++j)
                if (devices[j] == devices[i]) { set_error("device %d listed twice", devices[i]); return PG_ERR_ARG; }
            out->push_back(devices[i]);
        }
    }
    if (out->size() > (size_t)PG_MAX_PEERS) out->resize(PG_MAX_PEERS);
    return PG_OK;
}

// reusable barrier for the device threads of one call
class Rendezvous {
    std::mutex mu_;
    std::condition_variable cv_;
    int n_, waiting_ = 0;
    uint64_t gen_ = 0;
public:
    explicit Rendezvous(int n) : n_(n) {}
    void wait() {
        std::unique_lock<std::mutex> lk(mu_);
        const uint64_t g = gen_;
        if (++waiting_ == n_) { waiting_ = 0; ++gen_; cv_.notify_all(); }
        else cv_.wait(lk, [&] { return gen_ != g; });
    }
};

}  // namespace
}  // namespace pg

static int sketch_multi(const uint8_t *bases, const uint64_t *offsets, uint32_t ulen, uint64_t n_reads, int32_t k, int32_t s,
                        uint32_t flags, uint32_t *out, uint64_t row_stride, uint32_t *count, int32_t *status,
                        const int32_t *devices, int32_t n_devices) {
    int rc;
    if ((rc = check_ks(k, s)) != PG_OK) return rc;
    std::vector<int> dev;
    if ((rc = resolve_devices(devices, n_devices, &dev)) != PG_OK) return rc;
    if (n_reads == 0) return PG_OK;
    if (!bases || !out) { set_error("null buffer"); return PG_ERR_ARG; }
    const int P = (int)dev.size();
    // contiguous shards; boundaries on K1 tile boundaries (32 reads).  Ragged batches balance bytes.
    std::vector<uint64_t> cut(P + 1, n_reads);
    cut[0] = 0;
    if (!offsets) {
        uint64_t per = (n_reads + P - 1) / P;
        per = (per + 31) & ~31ull;
        for (int r = 1; r < P; ++r) cut[r] = std::min(n_reads, (uint64_t)r * per);
    } else {
        const uint64_t b0 = offsets[0], total = offsets[n_reads] - b0;
        for (int r = 1; r < P; ++r) {
            const uint64_t want = b0 + total / P * r;
            uint64_t i = std::lower_bound(offsets, offsets + n_reads + 1, want) - offsets;
            i = std::min<uint64_t>((i + 31) & ~31ull, n_reads);
            cut[r] = std::max(i, cut[r - 1]);
        }
    }
    std::vector<WorkerResult> res(P);
    std::vector<std::thread> th;
    for (int r = 0; r < P; ++r) {
        const uint64_t lo = cut[r], hi = cut[r + 1];
        if (hi <= lo) continue;
        th.emplace_back([=, &res] {
            t_bound_device = dev[r];  // this worker's entry points run on its device
            pg_numa_bind_thread(dev[r], nullptr);  // best effort: staging copies from the GPU's NUMA node
            const int wrc = offsets ? sketch_host(bases, offsets + lo, 0, hi - lo, k, s, flags, out + lo * row_stride, row_stride,
                                                  count ? count + lo : nullptr, status ? status + lo : nullptr)
                                    : sketch_host(bases + lo * (uint64_t)ulen, nullptr, ulen, hi - lo, k, s, flags, out + lo * row_stride,
                                                  row_stride, nullptr, status ? status + lo : nullptr);
            res[r].rc = wrc;
            if (wrc != PG_OK) res[r].msg = t_err;
        });
    }
    for (std::thread &t : th) t.join();
    return fold_results(res);
}

extern "C" int pg_mash_sketch_uniform_multi(const uint8_t *bases, uint64_t n_reads, uint32_t read_len, int32_t k, int32_t s,
                                            uint32_t flags, uint32_t *out, uint64_t row_stride, int32_t *status,
                                            const int32_t *devices, int32_t n_devices) {
    return sketch_multi(bases, nullptr, read_len, n_reads, k, s, flags, out, row_stride, nullptr, status, devices, n_devices);
}

extern "C" int pg_mash_sketch_batch_multi(const uint8_t *bases, const uint64_t *offsets, uint64_t n_reads, int32_t k, int32_t s,
                                          uint32_t flags, uint32_t *out, uint64_t row_stride, uint32_t *count, int32_t *status,
                                          const int32_t *devices, int32_t n_devices) {
    if (n_reads && !offsets) { set_error("null offsets"); return PG_ERR_ARG; }
    return sketch_multi(bases, offsets, 0, n_reads, k, s, flags, out, row_stride, count, status, devices, n_devices);
}

// Sketch + all-pairs distance over several devices of this process (cfg3 / cfg4 pipeline):
//   1. device r takes reads [lo_r, hi_r), copies them in and runs the sketch kernel, whose finished
//      tiles / rows are stored straight into the gathered buffer of EVERY device (peer access
//      enabled in-process): sketching and the all-gather are one kernel;
//   2. after a rendezvous every device holds all n sketches and computes row block r of the pair
//      matrix (receiver = row) with the same kernels as pg_mash_distance_block.
// Without peer access between two of the devices the exchange falls back to cudaMemcpyPeerAsync pulls.
extern "C" int pg_mash_sketch_distance_multi(const uint8_t *bases, uint64_t n_reads, uint32_t read_len, int32_t k, int32_t s,
                                             const int32_t *devices, int32_t n_devices, uint32_t *sketches, uint32_t *same,
                                             double *distance) {
    int rc;
    if ((rc = check_ks(k, s)) != PG_OK) return rc;
    std::vector<int> dev;
    if ((rc = resolve_devices(devices, n_devices, &dev)) != PG_OK) return rc;
    if (n_reads == 0) return PG_OK;
    if (!bases) { set_error("null buffer"); return PG_ERR_ARG; }
    if (s <= 0) { set_error("distance over sketches of size %d: the reference panics", s); return PG_ERR_PANIC; }
    const uint64_t n = n_reads, L = read_len;
    const uint64_t cnt = std::min<uint64_t>(kmers_of(L, k), (uint64_t)s);
    int P = (int)std::min<uint64_t>(dev.size(), (n + 31) / 32);
    dev.resize(P);
    std::vector<uint64_t> cut(P + 1, n);
    cut[0] = 0;
    {
        uint64_t per = (n + P - 1) / P;
        per = (per + 31) & ~31ull;
        for (int r = 1; r < P; ++r) cut[r] = std::min(n, (uint64_t)r * per);
    }
    // peer access available between every pair?
    bool fused = true;
    for (int a = 0; a < P && fused; ++a)
        for (int b = 0; b < P && fused; ++b) {
            int can = 1;
            if (a != b && cudaDeviceCanAccessPeer(&can, dev[a], dev[b]) != cudaSuccess) can = 0;
            fused = fused && can;
        }
    cudaGetLastError();
    std::vector<void *> gathered(P, nullptr);  // [n][cnt] compact rows on every device
    std::vector<WorkerResult> res(P);
    std::atomic<bool> failed{false};
    Rendezvous meet(P);
    auto worker = [&](int r) {
        WorkerResult &my = res[r];
        auto fail = [&](int code) { my.rc = code; my.msg = t_err; failed.store(true); };
        auto step = [&](auto &&fn) { if (!failed.load()) { const int c = fn(); if (c != PG_OK) fail(c); } };
        t_bound_device = dev[r];
        pg_numa_bind_thread(dev[r], nullptr);
        const uint64_t lo = cut[r], hi = cut[r + 1], nl = hi - lo;
        uint8_t *d_reads = nullptr;
        uint32_t *d_full = nullptr;  // [n][s] full Go arrays (== gathered when cnt == s)
        cudaStream_t st = nullptr;
        step([&]() -> int {
            int c = ensure_device();
            if (c != PG_OK) return c;
            st = t_ctx->streams[0];
            if (fused)
                for (int p = 0; p < P; ++p)
                    if (p != r) {
                        cudaError_t e = cudaDeviceEnablePeerAccess(dev[p], 0);
                        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cuda_fail(e, "cudaDeviceEnablePeerAccess", __FILE__, __LINE__);
                        cudaGetLastError();
                    }
            PG_CUDA(cudaMalloc(&gathered[r], std::max<uint64_t>(n * cnt * 4, 16)));
            PG_CUDA(cudaMalloc(&d_reads, std::max<uint64_t>(nl * L, 16) + 64));
            return PG_OK;
        });
        meet.wait();  // every buffer exists, peer access is on
        step([&]() -> int {
            if (nl) PG_CUDA(cudaMemcpyAsync(d_reads, bases + lo * L, nl * L, cudaMemcpyHostToDevice, st));
            int c;
            if (fused) {
                c = sketch_scatter_dev(d_reads, nl, read_len, k, s, gathered.data(), P, r, lo, st);
            } else {
                void *own = gathered[r];
                c = sketch_scatter_dev(d_reads, nl, read_len, k, s, &own, 1, 0, lo, st);
            }
            if (c != PG_OK) return c;
            PG_CUDA(cudaStreamSynchronize(st));
            return PG_OK;
        });
        meet.wait();  // all sketches computed (fused: and stored everywhere)
        step([&]() -> int {
            if (!fused)  // pull the other devices' row blocks
                for (int p = 0; p < P; ++p)
                    if (p != r && cut[p + 1] > cut[p])
                        PG_CUDA(cudaMemcpyPeerAsync((uint32_t *)gathered[r] + cut[p] * cnt, dev[r], (uint32_t *)gathered[p] + cut[p] * cnt, dev[p],
                                                    (cut[p + 1] - cut[p]) * cnt * 4, st));
            if (cnt == (uint64_t)s) {
                d_full = (uint32_t *)gathered[r];
            } else {  // the zero tail of a fresh Mash is materialised after the exchange
                PG_CUDA(cudaMalloc(&d_full, n * (uint64_t)s * 4));
                PG_CUDA(cudaMemsetAsync(d_full, 0, n * (uint64_t)s * 4, st));
                if (cnt) PG_CUDA(cudaMemcpy2DAsync(d_full, (size_t)s * 4, gathered[r], cnt * 4, cnt * 4, n, cudaMemcpyDeviceToDevice, st));
            }
            if (sketches && nl) PG_CUDA(cudaMemcpyAsync(sketches + lo * (uint64_t)s, d_full + lo * (uint64_t)s, nl * (uint64_t)s * 4, cudaMemcpyDeviceToHost, st));
            if ((same || distance) && nl) {
                DistancePlan plan;
                int c = distance_plan_create(d_full, n, s, st, &plan);
                const uint64_t rows_per = std::max<uint64_t>(8, ((64ull << 20) / n) & ~7ull);
                for (uint64_t rb = lo; rb < hi && c == PG_OK; rb += rows_per) {
                    const uint64_t re = std::min(hi, rb + rows_per);
                    Tmp d_same(st), d_dist(st);
                    if (same && (c = d_same.alloc((re - rb) * n * 4))) break;
                    if (distance && (c = d_dist.alloc((re - rb) * n * 8))) break;
                    c = distance_plan_rows(plan, rb, re, same ? d_same.as<uint32_t>() : nullptr, distance ? d_dist.as<double>() : nullptr, st);
                    if (c != PG_OK) break;
                    cudaError_t e = cudaSuccess;
                    if (same) e = cudaMemcpyAsync(same + rb * n, d_same.p, (re - rb) * n * 4, cudaMemcpyDeviceToHost, st);
                    if (e == cudaSuccess && distance) e = cudaMemcpyAsync(distance + rb * n, d_dist.p, (re - rb) * n * 8, cudaMemcpyDeviceToHost, st);
                    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
                    if (e != cudaSuccess) c = cuda_fail(e, "distance block copy", __FILE__, __LINE__);
                }
                distance_plan_destroy(plan, st);
                if (c != PG_OK) return c;
            }
            PG_CUDA(cudaStreamSynchronize(st));
            return PG_OK;
        });
        meet.wait();  // nobody reads a peer buffer any more
        if (t_ctx) {
            cudaSetDevice(dev[r]);
            if (d_full && d_full != gathered[r]) cudaFree(d_full);
            if (d_reads) cudaFree(d_reads);
            if (gathered[r]) cudaFree(gathered[r]);
            cudaGetLastError();
        }
    };
    std::vector<std::thread> th;
    for (int r = 0; r < P; ++r) th.emplace_back(worker, r);
    for (std::thread &t : th) t.join();
    return fold_results(res);
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
    std::lock_guard<std::mutex> lk(t_ctx->mu);
    cudaStream_t st = t_ctx->streams[0];
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
    std::lock_guard<std::mutex> lk(t_ctx->mu);
    cudaStream_t st = t_ctx->streams[0];
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

int pg_mash_distance_sparse_dev(const uint32_t *d_sketches, uint64_t n, int32_t s, uint64_t row_begin, uint64_t row_end, uint32_t flags,
                                uint32_t *d_pair_i, uint32_t *d_pair_j, uint32_t *d_pair_same, uint64_t pairs_cap, uint64_t *d_n_pairs,
                                void *stream) {
    int rc = ensure_device();
    if (rc != PG_OK) return rc;
    if (row_end > n || row_begin > row_end) { set_error("bad row range"); return PG_ERR_ARG; }
    if (!d_n_pairs || (pairs_cap && (!d_pair_i || !d_pair_j || !d_pair_same))) { set_error("null buffer"); return PG_ERR_ARG; }
    return launch_distance_sparse(d_sketches, n, s, row_begin, row_end, flags, d_pair_i, d_pair_j, d_pair_same, pairs_cap,
                                  (unsigned long long *)d_n_pairs, (cudaStream_t)stream);
}

int pg_mash_distance_sparse(const uint32_t *sketches, uint64_t n, int32_t s, uint64_t row_begin, uint64_t row_end, uint32_t flags,
                            uint32_t *pair_i, uint32_t *pair_j, uint32_t *pair_same, uint64_t pairs_cap, uint64_t *n_pairs) {
    int rc = ensure_device();
    if (rc != PG_OK) return rc;
    if (!n_pairs) { set_error("null buffer"); return PG_ERR_ARG; }
    *n_pairs = 0;
    if (row_end > n || row_begin > row_end) { set_error("bad row range"); return PG_ERR_ARG; }
    if (row_end == row_begin || n == 0) return PG_OK;
    if (s <= 0) { set_error("distance over sketches of size %d: the reference panics", s); return PG_ERR_PANIC; }
    if (pairs_cap && (!pair_i || !pair_j || !pair_same)) { set_error("null buffer"); return PG_ERR_ARG; }
    std::lock_guard<std::mutex> lk(t_ctx->mu);
    cudaStream_t st = t_ctx->streams[0];
    Tmp d_sk(st), d_i(st), d_j(st), d_c(st), d_n(st);
    if ((rc = d_sk.alloc(n * (uint64_t)s * 4)) || (rc = d_i.alloc(pairs_cap * 4)) || (rc = d_j.alloc(pairs_cap * 4)) ||
        (rc = d_c.alloc(pairs_cap * 4)) || (rc = d_n.alloc(8)))
        return rc;
    PG_CUDA(cudaMemcpyAsync(d_sk.p, sketches, n * (uint64_t)s * 4, cudaMemcpyHostToDevice, st));
    rc = launch_distance_sparse(d_sk.as<uint32_t>(), n, s, row_begin, row_end, flags, d_i.as<uint32_t>(), d_j.as<uint32_t>(), d_c.as<uint32_t>(),
                                pairs_cap, d_n.as<unsigned long long>(), st);
    if (rc != PG_OK) return rc;
    unsigned long long found = 0;
    PG_CUDA(cudaMemcpyAsync(&found, d_n.p, 8, cudaMemcpyDeviceToHost, st));
    PG_CUDA(cudaStreamSynchronize(st));
    *n_pairs = found;
    const uint64_t kept = std::min<uint64_t>(found, pairs_cap);
    if (kept) {
        PG_CUDA(cudaMemcpyAsync(pair_i, d_i.p, kept * 4, cudaMemcpyDeviceToHost, st));
        PG_CUDA(cudaMemcpyAsync(pair_j, d_j.p, kept * 4, cudaMemcpyDeviceToHost, st));
        PG_CUDA(cudaMemcpyAsync(pair_same, d_c.p, kept * 4, cudaMemcpyDeviceToHost, st));
        PG_CUDA(cudaStreamSynchronize(st));
    }
    if (found > pairs_cap) { set_error("pairs_cap %llu < %llu pairs", (unsigned long long)pairs_cap, found); return PG_ERR_ARG; }
    return PG_OK;
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
    std::lock_guard<std::mutex> lk(t_ctx->mu);
    cudaStream_t st = t_ctx->streams[0];
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
    std::lock_guard<std::mutex> lk(t_ctx->mu);
    cudaStream_t st = t_ctx->streams[0];
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
    std::lock_guard<std::mutex> lk(t_ctx->mu);
    cudaStream_t st = t_ctx->streams[0];
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
    std::lock_guard<std::mutex> lk(t_ctx->mu);
    cudaStream_t st = t_ctx->streams[0];
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
    std::lock_guard<std::mutex> lk(t_ctx->mu);
    cudaStream_t st = t_ctx->streams[0];
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
    std::lock_guard<std::mutex> lk(t_ctx->mu);
    cudaStream_t st = t_ctx->streams[0];
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

static int fastq_ingest_dev_impl(const uint8_t *d_text, uint64_t nbytes, uint8_t *d_bases, uint64_t bases_cap,
                                 uint64_t *d_offsets, uint64_t *d_spans, uint64_t records_cap, uint64_t *n_records,
                                 uint64_t *total_bases, int32_t *err_code, uint64_t *err_line, void *stream) {
    int rc = ensure_device();
    if (rc != PG_OK) return rc;
    if (!n_records || !total_bases || !err_code || !err_line || !d_offsets) { set_error("null buffer"); return PG_ERR_ARG; }
    return launch_fastq_ingest(d_text, nbytes, d_bases, bases_cap, d_offsets, records_cap, n_records, total_bases, err_code,
                               err_line, (cudaStream_t)stream, d_spans);
}

static int fastq_ingest_host_impl(const uint8_t *text, uint64_t nbytes, uint8_t *bases, uint64_t bases_cap,
                                  uint64_t *offsets, uint64_t *spans, uint64_t records_cap, uint64_t *n_records,
                                  uint64_t *total_bases, int32_t *err_code, uint64_t *err_line) {
    int rc = ensure_device();
    if (rc != PG_OK) return rc;
    if (!n_records || !total_bases || !err_code || !err_line || !offsets) { set_error("null buffer"); return PG_ERR_ARG; }
    std::lock_guard<std::mutex> lk(t_ctx->mu);
    cudaStream_t st = t_ctx->streams[0];
    Tmp d_text(st), d_bases(st), d_off(st), d_spans(st);
    if ((rc = d_text.alloc(nbytes + 16)) || (rc = d_bases.alloc(bases_cap + 16)) || (rc = d_off.alloc((records_cap + 1) * 8)))
        return rc;
    if (spans && (rc = d_spans.alloc((records_cap + 1) * 32))) return rc;
    if (nbytes) PG_CUDA(cudaMemcpyAsync(d_text.p, text, nbytes, cudaMemcpyHostToDevice, st));
    rc = launch_fastq_ingest(d_text.as<uint8_t>(), nbytes, d_bases.as<uint8_t>(), bases_cap, d_off.as<uint64_t>(), records_cap,
                             n_records, total_bases, err_code, err_line, st, spans ? d_spans.as<uint64_t>() : nullptr);
    if (rc != PG_OK) return rc;
    if (*total_bases) PG_CUDA(cudaMemcpyAsync(bases, d_bases.p, *total_bases, cudaMemcpyDeviceToHost, st));
    PG_CUDA(cudaMemcpyAsync(offsets, d_off.p, (*n_records + 1) * 8, cudaMemcpyDeviceToHost, st));
    if (spans && *n_records) PG_CUDA(cudaMemcpyAsync(spans, d_spans.p, *n_records * 32, cudaMemcpyDeviceToHost, st));
    PG_CUDA(cudaStreamSynchronize(st));
    return PG_OK;
}

int pg_fastq_ingest_dev(const uint8_t *d_text, uint64_t nbytes, uint8_t *d_bases, uint64_t bases_cap,
                        uint64_t *d_offsets, uint64_t records_cap, uint64_t *n_records,
                        uint64_t *total_bases, int32_t *err_code, uint64_t *err_line, void *stream) {
    return fastq_ingest_dev_impl(d_text, nbytes, d_bases, bases_cap, d_offsets, nullptr, records_cap, n_records, total_bases, err_code, err_line, stream);
}
int pg_fastq_ingest_records_dev(const uint8_t *d_text, uint64_t nbytes, uint8_t *d_bases, uint64_t bases_cap,
                                uint64_t *d_offsets, uint64_t *d_spans, uint64_t records_cap, uint64_t *n_records,
                                uint64_t *total_bases, int32_t *err_code, uint64_t *err_line, void *stream) {
    if (!d_spans) { set_error("null spans"); return PG_ERR_ARG; }
    return fastq_ingest_dev_impl(d_text, nbytes, d_bases, bases_cap, d_offsets, d_spans, records_cap, n_records, total_bases, err_code, err_line, stream);
}
int pg_fastq_ingest(const uint8_t *text, uint64_t nbytes, uint8_t *bases, uint64_t bases_cap,
                    uint64_t *offsets, uint64_t records_cap, uint64_t *n_records,
                    uint64_t *total_bases, int32_t *err_code, uint64_t *err_line) {
    return fastq_ingest_host_impl(text, nbytes, bases, bases_cap, offsets, nullptr, records_cap, n_records, total_bases, err_code, err_line);
}
int pg_fastq_ingest_records(const uint8_t *text, uint64_t nbytes, uint8_t *bases, uint64_t bases_cap,
                            uint64_t *offsets, uint64_t *spans, uint64_t records_cap, uint64_t *n_records,
                            uint64_t *total_bases, int32_t *err_code, uint64_t *err_line) {
    if (!spans) { set_error("null spans"); return PG_ERR_ARG; }
    return fastq_ingest_host_impl(text, nbytes, bases, bases_cap, offsets, spans, records_cap, n_records, total_bases, err_code, err_line);
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
    std::lock_guard<std::mutex> lk(t_ctx->mu);
    cudaStream_t st = t_ctx->streams[0];
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
