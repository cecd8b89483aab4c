/*
 * poly_b200.h -- C ABI of libpolyb200.so: the B200 (sm_100a) implementation of
 * bebop/poly's search/mash sketching hot path plus the two secondary batched
 * kernels (search/align.SmithWaterman score, primers.SantaLucia).
 *
 * The reference is pure Go and has NO FFI/plugin interface; the
 * drop-in boundary is the exported Go API of three leaf packages.  Each entry
 * point below names the reference function it replaces (file:line under
 * /root/reference); INTEGRATION.md shows the cgo binding for each.
 *
 * Conventions
 *   - plain C: pointers + sizes, no C++/torch types, no exceptions across the ABI;
 *   - every function returns PG_OK (0) or a PG_ERR_* code; pg_last_error() gives a
 *     thread-local message;
 *   - the caller owns every buffer; the library never retains a caller pointer
 *     after return (cgo pointer rule) and never frees caller memory;
 *   - functions without the _dev suffix take HOST pointers (pageable or pinned) and
 *     perform the host<->device copies themselves; *_dev functions take DEVICE
 *     pointers on the calling thread's device (pg_thread_device / pg_init) plus a
 *     CUDA stream (cudaStream_t cast to void*, NULL = the legacy default stream) and
 *     only enqueue work: the sketch, Smith-Waterman / Needleman-Wunsch score, Tm and
 *     synth *_dev entry points never wait for the device.  Exceptions, which
 *     synchronise the given stream before returning because a host decision depends
 *     on a device result: pg_mash_distance_block_dev / pg_mash_distance_sparse_dev
 *     (sortedness and bucket sizes choose the algorithm), pg_fastq_ingest_dev /
 *     pg_fasta_ingest_dev (they return counts and the error position by value), and
 *     pg_mash_sketch_batch_dev for ragged batches the threshold path does not take
 *     (s < 2, or more than 4 Mi row x item slots) when they hold few long sequences;
 *   - there is NO CPU fallback: without a usable sm_100 device every compute entry
 *     point fails with PG_ERR_NO_DEVICE.
 */
#ifndef POLY_B200_H
#define POLY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PG_OK 0
#define PG_ERR_CUDA 1        /* a CUDA runtime call failed (see pg_last_error)          */
#define PG_ERR_ARG 2         /* invalid argument                                         */
#define PG_ERR_NO_DEVICE 3   /* no CUDA device / not an sm_100 part                      */
#define PG_ERR_PANIC 4       /* >= 1 item hits an input on which the Go reference panics */
#define PG_ERR_UNSUPPORTED 5 /* >= 1 item is outside the supported domain                */
#define PG_ERR_NOMEM 6

/* per-item status codes (status arrays) */
#define PG_ITEM_OK 0
#define PG_ITEM_PANIC 1       /* reference panics (index out of range) on this item     */
#define PG_ITEM_UNSUPPORTED 2 /* e.g. byte >= 0x80 passed to SantaLucia (ToUpper)       */

/* ---- library / device management -------------------------------------------- */
int pg_version(void);
/* Choose the process-default CUDA device (one process per GPU layout).  Optional: the first
 * compute call initialises the current CUDA device (device 0 unless the caller changed it).
 * The library keeps one context per device and serves any number of devices from one process:
 *   - pg_thread_device(d) binds the CALLING THREAD's later calls to device d (-1: back to the
 *     process default) -- for hosts that drive several GPUs themselves (Go: after
 *     runtime.LockOSThread());
 *   - the *_multi entry points shard one batch over several devices on their own.
 * Entry points are safe to call concurrently from many threads: host-pointer entry points are
 * serialised per device, *_dev entry points only enqueue work on the caller's stream. */
int pg_init(int device);
int pg_thread_device(int device);
/* Move the calling thread to the CPUs of the NUMA node the device is attached to and prefer that
 * node for memory it allocates afterwards (call before pg_host_alloc).  Best effort; *node (may be
 * NULL) receives the node or -1 when nothing was done. */
int pg_numa_bind_thread(int device, int *node);
int pg_shutdown(void);
const char *pg_last_error(void);
int pg_device_count(int *count);
int pg_device_sm_count(int *sms);
/* Pinned host memory for callers that want full PCIe bandwidth (optional). */
int pg_host_alloc(void **ptr, size_t bytes);
int pg_host_free(void *ptr);
/* Device memory helpers for C/Go callers of the *_dev entry points. */
int pg_dev_alloc(void **dptr, size_t bytes);
int pg_dev_free(void *dptr);
int pg_memcpy_h2d(void *dst_dev, const void *src_host, size_t bytes, void *stream);
int pg_memcpy_d2h(void *dst_host, const void *src_dev, size_t bytes, void *stream);
int pg_stream_sync(void *stream);

/* ---- mash.Sketch -- replaces (*Mash).Sketch on a fresh mash.New(k, s) -----------
 * search/mash/mash.go:59-65 (New) and :68-104 (Sketch), one sketch per read.
 *
 * Read i is bases[offsets[i] .. offsets[i+1]) (raw bytes, any value; the hash is
 * MurmurHash3_x86_32(seed 0) of the k raw bytes, mash.go:74-76).  With
 * n_i = max(len_i - k, 0) hashes (mash.go:73 visits L-k windows, not L-k+1):
 *   n_i <  s : out row i = the n_i hashes in positional order   (mash.go:81-84)
 *   n_i >= s : out row i = ascending bottom-s multiset          (mash.go:87-102)
 * count[i] = min(n_i, s) informative words; row i starts at out + i*row_stride.
 * row_stride must be >= the largest count of the batch (>= s with PG_SKETCH_PAD_ZERO); words
 * [count[i], row_stride) of a row are written as zeros -- the zero tail a fresh Mash holds,
 * as far as the row reaches.  With PG_SKETCH_PAD_ZERO every row is a full Go array of s words.
 * status[i] (may be NULL) reports per-read PG_ITEM_PANIC for s in {0,1} inputs on
 * which mash.go:96-98 indexes Sketches[-1]; k < 0 or s < 0 is PG_ERR_ARG.
 * Reads may be of any length (a batch of few long sequences is cut into slices that are sketched
 * in parallel and merged) and k and s of any size: in the n_i >= s regime, sketches above 16384 words
 * or k above 1024 take a slower global-memory path instead of the shared-memory kernels.
 */
#define PG_SKETCH_PAD_ZERO 1u
/* Host-buffer entry points only: write nothing but the count[i] informative words of a row -- the literal
 * behaviour of (*Mash).Sketch, which never touches Sketches[n:s] in the n < s regime (mash.go:73-80).  A Go
 * caller that hands in a fresh make([]uint32, n*s) slab (zeroed by the runtime) gets full Sketches arrays
 * without the library writing 4*s bytes per read; a reused array keeps its old tail, as in the reference. */
#define PG_SKETCH_TAIL_KEEP 2u

int pg_mash_sketch_batch(const uint8_t *bases, const uint64_t *offsets, uint64_t n_reads, int32_t k,
                         int32_t s, uint32_t flags, uint32_t *out, uint64_t row_stride,
                         uint32_t *count, int32_t *status);
/* Same, for fixed-length reads stored back to back (read i at bases + i*read_len):
 * no offsets array, count is min(max(read_len-k,0), s) for every read. */
int pg_mash_sketch_uniform(const uint8_t *bases, uint64_t n_reads, uint32_t read_len, int32_t k,
                           int32_t s, uint32_t flags, uint32_t *out, uint64_t row_stride,
                           int32_t *status);
/* Device-resident variants (inputs already in HBM, outputs stay in HBM). */
int pg_mash_sketch_batch_dev(const uint8_t *d_bases, const uint64_t *d_offsets, uint64_t n_reads,
                             uint64_t max_read_len, int32_t k, int32_t s, uint32_t flags,
                             uint32_t *d_out, uint64_t row_stride, uint32_t *d_count,
                             int32_t *d_status, void *stream);
int pg_mash_sketch_uniform_dev(const uint8_t *d_bases, uint64_t n_reads, uint32_t read_len,
                               int32_t k, int32_t s, uint32_t flags, uint32_t *d_out,
                               uint64_t row_stride, int32_t *d_status, void *stream);
/* Fused sketch + all-gather over NVLink peer memory (one process per GPU).  Every rank owns a
 * "gathered" buffer of world*n_local rows; gathered_ptrs[p] is that buffer on rank p as seen from
 * THIS process (own pointer for p == rank, pg_ipc_import()ed peer mappings otherwise).  The
 * sketch kernel stores each finished tile / row straight into all of them at row offset
 * rank*n_local (TMA bulk stores to peer addresses in the fill regime), so the transfer overlaps
 * the hashing tile by tile and no separate collective pass runs.  Rows are compact
 * (row stride = min(max(read_len-k,0), s) words).  The caller must synchronise the stream and
 * barrier across ranks before reading its gathered buffer.  n_local MUST be the same on all
 * ranks (the row offset is rank*n_local; use the _scatter_dev variant for unequal shards); world <= 8. */
#define PG_MAX_PEERS 8
#define PG_IPC_HANDLE_BYTES 64
int pg_ipc_export(void *dptr, uint8_t handle[PG_IPC_HANDLE_BYTES]);
int pg_ipc_import(const uint8_t handle[PG_IPC_HANDLE_BYTES], void **dptr);
int pg_ipc_close(void *dptr);
int pg_mash_sketch_uniform_gather_dev(const uint8_t *d_bases, uint64_t n_local, uint32_t read_len,
                                      int32_t k, int32_t s, void *const *gathered_ptrs,
                                      int32_t world, int32_t rank, void *stream);
/* The same with an explicit row offset: the n_local sketches go to rows [row_offset, row_offset +
 * n_local) of each of the n_dst buffers (dst_ptrs[self] is the local one).  Shards may then differ
 * in size; any k, any n_local (rows the TMA path does not take are stored by the generic kernels). */
int pg_mash_sketch_uniform_scatter_dev(const uint8_t *d_bases, uint64_t n_local, uint32_t read_len,
                                       int32_t k, int32_t s, void *const *dst_ptrs, int32_t n_dst,
                                       int32_t self, uint64_t row_offset, void *stream);

/* ---- single-process multi-GPU (SURVEY.md 8b "multi-GPU variants taking a device count", 8e) ----
 * What a Go host calls to reach all GPUs of the box with ONE call: reads are independent
 * (mash.go:68-104 touches only its receiver), so the batch is cut into contiguous shards, one per
 * device, each driven by its own host thread through the pipelined host path -- no data-path
 * collective.  devices == NULL: the first n_devices visible devices (n_devices <= 0: all of them);
 * otherwise the listed ordinals.  Arguments and results are exactly those of pg_mash_sketch_uniform /
 * pg_mash_sketch_batch; host buffers may be pageable or pinned (pg_host_alloc). */
int pg_mash_sketch_uniform_multi(const uint8_t *bases, uint64_t n_reads, uint32_t read_len, int32_t k,
                                 int32_t s, uint32_t flags, uint32_t *out, uint64_t row_stride,
                                 int32_t *status, const int32_t *devices, int32_t n_devices);
int pg_mash_sketch_batch_multi(const uint8_t *bases, const uint64_t *offsets, uint64_t n_reads,
                               int32_t k, int32_t s, uint32_t flags, uint32_t *out,
                               uint64_t row_stride, uint32_t *count, int32_t *status,
                               const int32_t *devices, int32_t n_devices);
/* mash.Sketch of every read followed by all-pairs Similarity/Distance (mash.go:68-104, 107-140) on
 * several devices: device r sketches its shard and the sketch kernel itself stores every finished
 * tile / row into the gathered buffer of EVERY device (in-process peer access over NVLink: the
 * all-gather is fused into the kernel), then computes row block r of the pair matrix.  Outputs (host,
 * each may be NULL): sketches n x s (full Go arrays, zero tail included), same n x n uint32,
 * distance n x n double; receiver = row.  Without peer access the exchange uses peer copies. */
int pg_mash_sketch_distance_multi(const uint8_t *bases, uint64_t n_reads, uint32_t read_len, int32_t k,
                                  int32_t s, const int32_t *devices, int32_t n_devices,
                                  uint32_t *sketches, uint32_t *same, double *distance);

/* Name of the kernel the last uniform call on this thread dispatched to and how many
 * kernels it launched (bench.py's gpu_launches evidence). */
const char *pg_last_kernel(void);
uint64_t pg_launch_count(void);

/* ---- (*Mash).Similarity / Distance -- search/mash/mash.go:107-140 ----------------
 * Sketch j is words sketches[sk_offsets[j] .. sk_offsets[j+1]) and its length is its
 * SketchSize (the full array a Go Mash holds, zero tail included).  For pair p the
 * receiver is pair_a[p], the argument pair_b[p].  Literal reference semantics:
 * larger/smaller by SketchSize with the receiver "larger" on ties (:109-115), the
 * range early-out (:117-119), the two-pointer walk (:121-132) -- exact also for
 * unsorted (n < s) sketches.  same[p] = matching count, similarity[p] =
 * same/smaller.SketchSize (:134), distance[p] = 1 - similarity (:139).  Any of
 * same/similarity/distance may be NULL.  A sketch of size 0 is PG_ITEM_PANIC.
 */
int pg_mash_similarity_pairs(const uint32_t *sketches, const uint64_t *sk_offsets,
                             uint64_t n_sketches, const uint32_t *pair_a, const uint32_t *pair_b,
                             uint64_t n_pairs, int64_t *same, double *similarity,
                             double *distance, int32_t *status);
int pg_mash_similarity_pairs_dev(const uint32_t *d_sketches, const uint64_t *d_sk_offsets,
                                 uint64_t n_sketches, const uint32_t *d_pair_a,
                                 const uint32_t *d_pair_b, uint64_t n_pairs, int64_t *d_same,
                                 double *d_similarity, double *d_distance, int32_t *d_status,
                                 void *stream);
/* All pairs of a row block against a set of equal-size sketches: rows
 * [row_begin, row_end) of the n x n matrix, receiver = row, argument = column.
 * sketches is n x s (full Go arrays).  same is (row_end-row_begin) x n uint32,
 * distance likewise double (either may be NULL).  Same literal semantics. */
int pg_mash_distance_block(const uint32_t *sketches, uint64_t n, int32_t s, uint64_t row_begin,
                           uint64_t row_end, uint32_t *same, double *distance);
int pg_mash_distance_block_dev(const uint32_t *d_sketches, uint64_t n, int32_t s,
                               uint64_t row_begin, uint64_t row_end, uint32_t *d_same,
                               double *d_distance, void *stream);

/* The same row block as (i, j, same) triples of the pairs that share at least one hash: every ordered
 * pair (i, j), row_begin <= i < row_end, j != i (with PG_PAIRS_UPPER: j > i only), whose matching count is
 * non-zero, in no particular order.  Pairs that are not listed have same == 0 (Similarity 0, Distance 1,
 * mash.go:134,139); the diagonal is never listed.  For n sketches this returns O(related pairs) instead of
 * the n x rows matrix (cfg3, j > i: 2.8e7 triples = 334 MB instead of 40 GB).  *n_pairs receives the number of
 * qualifying pairs; PG_ERR_ARG (with the first pairs_cap stored) if it exceeds pairs_cap.  The _dev variant
 * takes a DEVICE counter d_n_pairs (which may exceed pairs_cap) and never blocks on the result. */
#define PG_PAIRS_UPPER 1u
int pg_mash_distance_sparse(const uint32_t *sketches, uint64_t n, int32_t s, uint64_t row_begin,
                            uint64_t row_end, uint32_t flags, uint32_t *pair_i, uint32_t *pair_j,
                            uint32_t *pair_same, uint64_t pairs_cap, uint64_t *n_pairs);
int pg_mash_distance_sparse_dev(const uint32_t *d_sketches, uint64_t n, int32_t s, uint64_t row_begin,
                                uint64_t row_end, uint32_t flags, uint32_t *d_pair_i, uint32_t *d_pair_j,
                                uint32_t *d_pair_same, uint64_t pairs_cap, uint64_t *d_n_pairs,
                                void *stream);

/* ---- align.SmithWaterman score -- search/align/align.go:171-203 -------------------
 * One template against n queries.  query_is_a != 0: stringA = query (outer loop),
 * stringB = template; else swapped.  lut_a/lut_b: byte -> index into the first /
 * second alphabet, -1 = not in alphabet (alphabet/alphabet.go:35-41); table is
 * n_a x n_b row-major (search/align/matrix/matrix.go:28-38); gap is ADDED
 * (align.go:193-194).  Per query: score (align.go:197-201 running max),
 * err_code 0 / 1 (symbol of stringA) / 2 (symbol of stringB) and err_pos (byte
 * index in that string) of the first failing cell in the reference's row-major
 * visiting order (align.go:188-191); on error score is 0.
 */
int pg_sw_score_batch(const uint8_t *queries, const uint64_t *q_offsets, uint64_t n_queries,
                      const uint8_t *templ, uint64_t templ_len, int32_t query_is_a,
                      const int16_t *lut_a, const int16_t *lut_b, const int64_t *table,
                      int32_t n_a, int32_t n_b, int64_t gap, int64_t *score, int32_t *err_code,
                      int64_t *err_pos);
int pg_sw_score_batch_dev(const uint8_t *d_queries, const uint64_t *d_q_offsets, uint64_t n_queries,
                          uint64_t max_query_len, const uint8_t *d_templ, uint64_t templ_len,
                          int32_t query_is_a, const int16_t *lut_a_host, const int16_t *lut_b_host,
                          const int64_t *table_host, int32_t n_a, int32_t n_b, int64_t gap,
                          int64_t *d_score, int32_t *d_err_code, int64_t *d_err_pos, void *stream);

/* ---- align.SmithWaterman with the aligned strings -- search/align/align.go:171-232 in full ----
 * (SURVEY.md 8f.1).  As pg_sw_score_batch, plus for every query the two aligned strings of the
 * reference's traceback (first maximum in row-major order, diagonal > up > left, align.go:205-229):
 * align_a / align_b hold n_queries rows of out_stride bytes, align_len[i] the common length.
 * status[i]: PG_ITEM_OK, or PG_ITEM_UNSUPPORTED if the alignment is longer than out_stride
 * (align_len[i] then reports the needed length).  Limits: queries of <= 64 symbols and scores
 * that fit 32 bits (PG_ERR_UNSUPPORTED otherwise). */
int pg_sw_align_batch(const uint8_t *queries, const uint64_t *q_offsets, uint64_t n_queries,
                      const uint8_t *templ, uint64_t templ_len, int32_t query_is_a,
                      const int16_t *lut_a, const int16_t *lut_b, const int64_t *table,
                      int32_t n_a, int32_t n_b, int64_t gap, int64_t *score, int32_t *err_code,
                      int64_t *err_pos, uint8_t *align_a, uint8_t *align_b, uint64_t out_stride,
                      uint32_t *align_len, int32_t *status);

/* align.NeedlemanWunsch in full (align.go:100-166): same interface; the traceback starts at
 * (len(a), len(b)) and runs while both indices are positive (align.go:141), so a leading rest of
 * one string is not emitted, exactly as in the reference. */
int pg_nw_align_batch(const uint8_t *queries, const uint64_t *q_offsets, uint64_t n_queries,
                      const uint8_t *templ, uint64_t templ_len, int32_t query_is_a,
                      const int16_t *lut_a, const int16_t *lut_b, const int64_t *table,
                      int32_t n_a, int32_t n_b, int64_t gap, int64_t *score, int32_t *err_code,
                      int64_t *err_pos, uint8_t *align_a, uint8_t *align_b, uint64_t out_stride,
                      uint32_t *align_len, int32_t *status);

/* ---- align.NeedlemanWunsch score -- search/align/align.go:100-134 (fill) and :166 --------
 * (a "next" row of SURVEY.md 8f).  Same arguments and error semantics as the Smith-Waterman
 * entry points; score = matrix[len(a)][len(b)] of the global alignment (gap ramps on the first
 * row and column, no zero floor). */
int pg_nw_score_batch(const uint8_t *queries, const uint64_t *q_offsets, uint64_t n_queries,
                      const uint8_t *templ, uint64_t templ_len, int32_t query_is_a,
                      const int16_t *lut_a, const int16_t *lut_b, const int64_t *table,
                      int32_t n_a, int32_t n_b, int64_t gap, int64_t *score, int32_t *err_code,
                      int64_t *err_pos);
int pg_nw_score_batch_dev(const uint8_t *d_queries, const uint64_t *d_q_offsets, uint64_t n_queries,
                          uint64_t max_query_len, const uint8_t *d_templ, uint64_t templ_len,
                          int32_t query_is_a, const int16_t *lut_a_host, const int16_t *lut_b_host,
                          const int64_t *table_host, int32_t n_a, int32_t n_b, int64_t gap,
                          int64_t *d_score, int32_t *d_err_code, int64_t *d_err_pos, void *stream);

/* ---- primers.SantaLucia / MeltingTemp -- primers/primers.go:70-105,121-128 --------
 * tm/dh/ds may each be NULL.  status: PG_ITEM_PANIC for an empty primer
 * (primers.go:89), PG_ITEM_UNSUPPORTED for a byte >= 0x80 (strings.ToUpper would
 * re-encode it, primers.go:71).  MeltingTemp == cp 500e-9, na 50e-3, mg 0. */
int pg_tm_batch(const uint8_t *bases, const uint64_t *offsets, uint64_t n, double cp, double na,
                double mg, double *tm, double *dh, double *ds, int32_t *status);
int pg_tm_batch_dev(const uint8_t *d_bases, const uint64_t *d_offsets, uint64_t n, double cp,
                    double na, double mg, double *d_tm, double *d_dh, double *d_ds,
                    int32_t *d_status, void *stream);

/* ---- pcr.DesignPrimersWithOverhangs core -- primers/pcr/pcr.go:44-60 ---------
 * For each sequence: fwd_len = length of the forward primer (shortest prefix of >= 15 nt of the
 * upper-cased sequence with MeltingTemp >= target_tm, pcr.go:46-49), rev_len = length of the
 * reverse primer (reverse complement of the shortest such suffix, pcr.go:50-53).  Every candidate
 * is evaluated exactly as primers.MeltingTemp would.  status: PG_ITEM_PANIC where the reference
 * slices out of range (sequence shorter than 15 nt or exhausted before the target is reached),
 * PG_ITEM_UNSUPPORTED for bytes >= 0x80.  The primer strings themselves (overhang + primer,
 * pcr.go:55-59) are assembled by the caller from the lengths. */
int pg_design_primers_batch(const uint8_t *bases, const uint64_t *offsets, uint64_t n,
                            double target_tm, uint32_t *fwd_len, uint32_t *rev_len, int32_t *status);

/* ---- pcr.SimulateSimple building blocks -- primers/pcr/pcr.go:73-169 ---------
 * pg_pcr_minimal_primer_batch: the minimal-primer loop of pcr.go:93-100 for every primer:
 * min_len = the longest 3' suffix of >= 7 nt (minimalPrimerLength, pcr.go:35) whose MeltingTemp is still below target_tm (the
 * reference's loop keeps the last length that FAILED the test), 0 when the 7-mer already reaches
 * it, the primer's length when even the whole primer stays below (the reference then ignores the
 * primer, pcr.go:103).  status: PG_ITEM_PANIC for primers shorter than 7 nt, PG_ITEM_UNSUPPORTED
 * for bytes >= 0x80.
 * pg_find_sites_batch: every (possibly overlapping) exact occurrence of every pattern in every
 * sequence -- what suffixarray.Lookup(pattern, -1) returns at pcr.go:110,113; empty patterns have
 * none.  flags & PG_SITES_UPPER compares the ASCII-upper-cased sequence bytes (pcr.go:82).  Hits
 * come back unordered as (sequence index, position inside it, pattern index); *n_hits is the
 * number found, PG_ERR_ARG (with the first hits_cap stored) if it exceeds hits_cap. */
#define PG_SITES_UPPER 1u
int pg_pcr_minimal_primer_batch(const uint8_t *bases, const uint64_t *offsets, uint64_t n,
                                double target_tm, uint32_t *min_len, int32_t *status);
int pg_find_sites_batch(const uint8_t *seqs, const uint64_t *seq_offsets, uint64_t n_seq,
                        const uint8_t *patterns, const uint64_t *pat_offsets, uint32_t n_pat,
                        uint32_t flags, uint32_t *hit_seq, uint64_t *hit_pos, uint32_t *hit_pat,
                        uint64_t hits_cap, uint64_t *n_hits);

/* ---- FASTQ ingest -- io/fastq Parser.ParseNext / ParseN, io/fastq/fastq.go:88-99,117-214 -------
 * (the step before the hot path.)  Parses a whole FASTQ text buffer on the GPU
 * into the dense bases + offsets layout the sketch entry points take.  Strict 4-line records;
 * parsing stops at the first record the reference rejects and the records before it are returned
 * together with the error, as ParseN does: *err_code 0 none, 1 a line of a record is not newline
 * terminated (EOF inside a record), 2 empty sequence, 3 empty quality, 4 no '@', 5 the reference
 * panics (empty identifier line, or an optional datum without '='), 6 line longer than the 64 KiB
 * reader of fastq.Parse; *err_line = 1-based number of the line the reference stops at.
 * Returns PG_ERR_ARG when a capacity is too small (*n_records / *total_bases hold the needs).
 * The _dev variant takes device buffers (text, bases, offsets) and host scalars. */
int pg_fastq_ingest(const uint8_t *text, uint64_t nbytes, uint8_t *bases, uint64_t bases_cap,
                    uint64_t *offsets, uint64_t records_cap, uint64_t *n_records,
                    uint64_t *total_bases, int32_t *err_code, uint64_t *err_line);
int pg_fastq_ingest_dev(const uint8_t *d_text, uint64_t nbytes, uint8_t *d_bases, uint64_t bases_cap,
                        uint64_t *d_offsets, uint64_t records_cap, uint64_t *n_records,
                        uint64_t *total_bases, int32_t *err_code, uint64_t *err_line, void *stream);

/* The same plus what fastq.Fastq holds besides the sequence (io/fastq/fastq.go:46-51): for record i
 * spans[4i .. 4i+3] = {begin, length of the identifier line (with its '@', without the newline),
 * begin, length of the quality line} as offsets into `text`.  The caller cuts Identifier
 * (strings.Split(line, " ")[0][1:], fastq.go:158), Optionals (the "key=value" tokens, fastq.go:159-165)
 * and Quality (fastq.go:199) out of the text it already holds: no second copy of the text is made. */
int pg_fastq_ingest_records(const uint8_t *text, uint64_t nbytes, uint8_t *bases, uint64_t bases_cap,
                            uint64_t *offsets, uint64_t *spans, uint64_t records_cap, uint64_t *n_records,
                            uint64_t *total_bases, int32_t *err_code, uint64_t *err_line);
int pg_fastq_ingest_records_dev(const uint8_t *d_text, uint64_t nbytes, uint8_t *d_bases, uint64_t bases_cap,
                                uint64_t *d_offsets, uint64_t *d_spans, uint64_t records_cap, uint64_t *n_records,
                                uint64_t *total_bases, int32_t *err_code, uint64_t *err_line, void *stream);

/* ---- FASTA ingest -- fasta.Parse = NewParser(r, maxLineSize).ParseAll(),
 * io/fasta/fasta.go:72-77,96-118,149-243 -----------------------------------------
 * Parses a whole FASTA text buffer on the GPU into dense sequences + offsets and (optionally:
 * pass names, name_offsets both non-NULL) dense names + name_offsets, n_records + 1 offsets each.
 * max_line_size as NewParser's (fasta.Parse uses 65536; values < 16 behave as 16 like
 * bufio.NewReaderSize).  Semantics are ParseNext's, quirks included: lines of length <= 1 and
 * ';' lines are skipped, a '>' line directly after a name line is sequence text, a record
 * without a trailing newline at the end of the text is dropped without an error, and parsing
 * stops at the first error with the records before it returned: *err_code 0 none,
 * 1 "did not find fasta start '>'", 2 "empty fasta sequence", 3 "line too large for buffer"
 * (>= max_line_size content bytes), 4 bufio.ErrBufferFull (over-long ';' line inside a record);
 * *err_line = the line number the reference's message prints.
 * flags & PG_FASTA_BUFIO_ALIAS: also reproduce the reference's use of the bufio line slice after
 * Peek(1) (fasta.go:192): a line whose newline is the last byte of a full reader buffer is seen
 * with its bytes replaced by the text one buffer further on, exactly as fasta.Parse over a
 * strings.Reader / bytes.Reader / *os.File yields.  Without the flag lines are taken as written.
 * Returns PG_ERR_ARG when a capacity is too small (*n_records, *total_bases, *total_name_bytes
 * then hold the needs; nbytes bounds both byte counts, the newline count + 1 the records). */
#define PG_FASTA_BUFIO_ALIAS 1u
int pg_fasta_ingest(const uint8_t *text, uint64_t nbytes, uint32_t max_line_size, uint32_t flags,
                    uint8_t *bases, uint64_t bases_cap, uint64_t *offsets, uint8_t *names,
                    uint64_t names_cap, uint64_t *name_offsets, uint64_t records_cap,
                    uint64_t *n_records, uint64_t *total_bases, uint64_t *total_name_bytes,
                    int32_t *err_code, uint64_t *err_line);
int pg_fasta_ingest_dev(const uint8_t *d_text, uint64_t nbytes, uint32_t max_line_size, uint32_t flags,
                        uint8_t *d_bases, uint64_t bases_cap, uint64_t *d_offsets, uint8_t *d_names,
                        uint64_t names_cap, uint64_t *d_name_offsets, uint64_t records_cap,
                        uint64_t *n_records, uint64_t *total_bases, uint64_t *total_name_bytes,
                        int32_t *err_code, uint64_t *err_line, void *stream);

/* ---- synthetic workloads (bench/test tooling, not a reference API) ---
 * kind 0: independent reads  base(i,j) = code(seed, i*L + j)
 * kind 1: family reads       (family = reads per template, 1/64 substitutions)
 * Writes n_reads*read_len bytes for reads [first_read, first_read+n_reads). */
int pg_synth_reads_dev(uint8_t *d_bases, uint64_t first_read, uint64_t n_reads, uint32_t read_len,
                       uint64_t seed, int32_t kind, uint32_t family, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* POLY_B200_H */
