// common.cuh -- shared declarations of libpolyb200 (internal; the public surface is
// include/poly_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/poly_b200.h"

namespace pg {

// thread-local error string + helpers (api.cu)
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);
void note_launch(const char *kernel_name);

#define PG_CUDA(call)                                                     \
    do {                                                                  \
        cudaError_t _e = (call);                                          \
        if (_e != cudaSuccess) return ::pg::cuda_fail(_e, #call, __FILE__, __LINE__); \
    } while (0)

#define PG_LAUNCH_CHECK(name)                                             \
    do {                                                                  \
        ::pg::note_launch(name);                                          \
        cudaError_t _e = cudaGetLastError();                              \
        if (_e != cudaSuccess) return ::pg::cuda_fail(_e, name, __FILE__, __LINE__); \
    } while (0)

int sm_count();
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) cached per (device, kernel); thread-safe (api.cu)
int func_smem(const void *kernel, size_t bytes);

// Stream-ordered temporaries that are released together when the scope ends -- also on the early
// returns of PG_CUDA -- with cudaFreeAsync on the stream they were allocated on.
struct StreamScratch {
    cudaStream_t st;
    void *ptr[24];
    int n = 0;
    explicit StreamScratch(cudaStream_t s) : st(s) {}
    StreamScratch(const StreamScratch &) = delete;
    StreamScratch &operator=(const StreamScratch &) = delete;
    ~StreamScratch() {
        for (int i = 0; i < n; ++i) cudaFreeAsync(ptr[i], st);
    }
    template <typename T>
    cudaError_t alloc(T **p, uint64_t count) {  // at least one element, so the pointer is always valid
        if (n >= 24) return cudaErrorMemoryAllocation;
        cudaError_t e = cudaMallocAsync((void **)p, (count ? count : 1) * sizeof(T), st);
        if (e == cudaSuccess) ptr[n++] = *p;
        return e;
    }
    void adopt(void *p) {  // take ownership of a buffer allocated elsewhere on the same stream
        if (p && n < 24) ptr[n++] = p;
    }
};

// destinations of a sketch row block: the local buffer, or the gathered buffers of all ranks
struct SketchDst {
    uint32_t *ptr[PG_MAX_PEERS];
    int n;
};

// ---- kernel launchers (one per .cu) -------------------------------------------
// sketch_fill.cu
int launch_sketch_uniform(const uint8_t *d_bases, uint64_t n_reads, uint32_t read_len, int k, int s,
                          uint32_t flags, uint32_t *d_out, uint64_t row_stride, int32_t *d_status,
                          cudaStream_t st, const SketchDst *extra = nullptr);
int launch_sketch_ragged(const uint8_t *d_bases, const uint64_t *d_offsets, uint64_t n_reads,
                         uint64_t max_read_len, int k, int s, uint32_t flags, uint32_t *d_out,
                         uint64_t row_stride, uint32_t *d_count, int32_t *d_status, cudaStream_t st);
// sketch_select.cu  (reads with n >= s; uniform: read_len != 0 and d_offsets == nullptr)
int launch_sketch_select(const uint8_t *d_bases, const uint64_t *d_offsets, uint32_t read_len,
                         uint64_t n_reads, int k, int s, uint32_t flags, uint32_t *d_out,
                         uint64_t row_stride, uint32_t *d_count, int32_t *d_status, cudaStream_t st,
                         const SketchDst *extra = nullptr, uint64_t max_read_len = 0);
// sketch_select_large.cu: the select regime for s > 16384 or k > 1024 (hashes through global memory)
int launch_sketch_select_large(const uint8_t *d_bases, const uint64_t *d_offsets, uint32_t read_len, uint64_t n_reads, int k, int s,
                               uint32_t *d_out, uint64_t row_stride, uint32_t *d_count, int32_t *d_status, cudaStream_t st,
                               const SketchDst &ex);
// distance.cu
int launch_similarity_pairs(const uint32_t *d_sk, const uint64_t *d_off, uint64_t n_sk,
                            const uint32_t *d_a, const uint32_t *d_b, uint64_t n_pairs,
                            int64_t *d_same, double *d_sim, double *d_dist, int32_t *d_status,
                            cudaStream_t st);
int launch_distance_block(const uint32_t *d_sk, uint64_t n, int s, uint64_t row_begin,
                          uint64_t row_end, uint32_t *d_same, double *d_dist, cudaStream_t st);
int launch_distance_sparse(const uint32_t *d_sk, uint64_t n, int s, uint64_t row_begin, uint64_t row_end, uint32_t flags,
                           uint32_t *d_i, uint32_t *d_j, uint32_t *d_same, uint64_t cap, unsigned long long *d_n_pairs,
                           cudaStream_t st);
// distance_join.cu (ascending sketches only): bucketed (value, id) index, built once per sketch
// set and reused for every row block
struct JoinIndex {
    uint64_t *entries = nullptr, *start = nullptr;
    uint64_t nb = 0, n = 0;
    int s = 0;
    bool ok = false;
};
int join_build(const uint32_t *d_sk, uint64_t n, int s, cudaStream_t st, JoinIndex *ix);
int join_emit(const JoinIndex &ix, uint64_t row_begin, uint64_t row_end, uint32_t *d_same, double *d_dist, cudaStream_t st);
void join_free(JoinIndex &ix, cudaStream_t st);
// distance.cu: a plan = sortedness flags (+ join index when every sketch is ascending)
struct DistancePlan {
    const uint32_t *d_sk = nullptr;
    uint64_t n = 0;
    int s = 0;
    uint8_t *d_flag = nullptr;
    JoinIndex join;
};
int distance_plan_create(const uint32_t *d_sk, uint64_t n, int s, cudaStream_t st, DistancePlan *plan);
int distance_plan_rows(const DistancePlan &plan, uint64_t row_begin, uint64_t row_end, uint32_t *d_same, double *d_dist, cudaStream_t st);
void distance_plan_destroy(DistancePlan &plan, cudaStream_t st);
// sw_score.cu
int launch_sw_score(const uint8_t *d_q, const uint64_t *d_qoff, uint64_t nq, uint64_t max_qlen,
                    const uint8_t *d_t, uint64_t tlen, int query_is_a, const int16_t *lut_a,
                    const int16_t *lut_b, const int64_t *table, int n_a, int n_b, int64_t gap,
                    int64_t *d_score, int32_t *d_err, int64_t *d_errpos, cudaStream_t st, int global = 0);
// sw_align.cu (score + aligned strings; queries <= 64 symbols, 32-bit scores)
int launch_sw_align(const uint8_t *d_q, const uint64_t *d_qoff, uint64_t nq, uint64_t max_qlen,
                    const uint8_t *d_t, uint64_t tlen, int query_is_a, const int16_t *lut_a,
                    const int16_t *lut_b, const int64_t *table, int n_a, int n_b, int64_t gap,
                    int64_t *d_score, int32_t *d_err, int64_t *d_errpos, uint8_t *d_align_a,
                    uint8_t *d_align_b, uint64_t out_stride, uint32_t *d_len, int32_t *d_status,
                    cudaStream_t st, int global = 0);
// sw_align_long.cu (both strings longer than 64 symbols: full matrix in HBM)
int launch_sw_align_long(const uint8_t *d_q, const uint64_t *d_qoff, uint64_t nq, uint64_t max_qlen, const uint8_t *d_t,
                         uint64_t tlen, int query_is_a, const int16_t *lut_a, const int16_t *lut_b, const int64_t *table,
                         int n_a, int n_b, int64_t gap, const int64_t *d_score, const int32_t *d_err, uint8_t *d_align_a,
                         uint8_t *d_align_b, uint64_t out_stride, uint32_t *d_len, int32_t *d_status, cudaStream_t st,
                         int global);
// tm.cu
int launch_tm(const uint8_t *d_bases, const uint64_t *d_off, uint64_t n, double cp, double na,
              double mg, double *d_tm, double *d_dh, double *d_ds, int32_t *d_status,
              cudaStream_t st);
int launch_design_primers(const uint8_t *d_bases, const uint64_t *d_off, uint64_t n, double target,
                          uint32_t *d_fwd, uint32_t *d_rev, int32_t *d_status, cudaStream_t st);
int launch_minimal_primer(const uint8_t *d_bases, const uint64_t *d_off, uint64_t n, double target,
                          uint32_t *d_min_len, int32_t *d_status, cudaStream_t st);
// find_sites.cu
int launch_find_sites(const uint8_t *d_seqs, const uint64_t *d_seq_off, uint64_t n_seq, uint64_t total_bytes,
                      const uint8_t *d_pats, const uint64_t *d_pat_off, uint32_t n_pat, uint64_t pat_bytes, uint32_t flags,
                      uint32_t *d_hit_seq, uint64_t *d_hit_pos, uint32_t *d_hit_pat, uint64_t cap,
                      unsigned long long *d_n_hits, cudaStream_t st);
// fastq_ingest.cu
int launch_fastq_ingest(const uint8_t *d_text, uint64_t nbytes, uint8_t *d_bases, uint64_t bases_cap,
                        uint64_t *d_offsets, uint64_t records_cap, uint64_t *n_records,
                        uint64_t *total_bases, int32_t *err_code, uint64_t *err_line, cudaStream_t st,
                        uint64_t *d_spans = nullptr);
// fasta_ingest.cu
int launch_fasta_ingest(const uint8_t *d_text, uint64_t nbytes, uint32_t max_line_size, uint32_t flags,
                        uint8_t *d_bases, uint64_t bases_cap, uint64_t *d_offsets, uint8_t *d_names,
                        uint64_t names_cap, uint64_t *d_name_offsets, uint64_t records_cap, uint64_t *n_records,
                        uint64_t *total_bases, uint64_t *total_name_bytes, int32_t *err_code, uint64_t *err_line,
                        cudaStream_t st);
// synth.cu
int launch_synth_reads(uint8_t *d_bases, uint64_t first_read, uint64_t n_reads, uint32_t read_len,
                       uint64_t seed, int kind, uint32_t family, cudaStream_t st);

}  // namespace pg
