/*
 * oracle/poly_oracle.h -- CPU restatement of the bebop/poly hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: it
 * may be imported/linked/executed only by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs, and only as the checker (or
 * the timed CPU baseline) -- never as a fallback for the CUDA path.
 *
 * Parity pinning: the Go reference cannot be executed in this image (no Go
 * toolchain).  The restatement is pinned by (1) MurmurHash3_x86_32 known-answer
 * vectors (the algorithm lives in github.com/spaolacci/murmur3 v1.1.0, go.mod:12,
 * which is not vendored under /root/reference), (2) a replay of every assertion
 * in the reference's own tests for this path (search/mash/mash_test.go:9-62,
 * search/align/align_test.go:139-292, search/align/example_test.go:82,110,
 * primers/primers_test.go:29-84, primers/pcr/example_test.go:36,46,54) and
 * (3) the golden checksums of SURVEY.md section 8c/8d.  See tests/test_oracle_*.py.
 *
 * All citations are file:line under /root/reference.
 */
#ifndef POLY_ORACLE_H
#define POLY_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define PO_OK 0
#define PO_PANIC (-1)       /* the Go reference would panic on this input        */
#define PO_UNSUPPORTED (-2) /* outside the restated domain (non-ASCII ToUpper)   */

/* github.com/spaolacci/murmur3 v1.1.0 Sum32WithSeed == MurmurHash3_x86_32. */
uint32_t po_murmur3_32(const uint8_t *data, size_t len, uint32_t seed);

/* search/mash/mash.go:68-104, literally (full re-sort on every qualifying insert).
 * `sketches` has length s and is updated in place exactly as the receiver is. */
int po_mash_sketch_faithful(const uint8_t *seq, int64_t len, int k, int s, uint32_t *sketches);
/* Closed form of the same function: identical results. */
int po_mash_sketch_closed(const uint8_t *seq, int64_t len, int k, int s, uint32_t *sketches);

/* search/mash/mash.go:107-135.  `a` is the receiver.  Writes the matching count. */
int po_mash_similarity(const uint32_t *a, int sa, const uint32_t *b, int sb,
                       int64_t *same, double *similarity);
/* search/mash/mash.go:138-140. */
int po_mash_distance(const uint32_t *a, int sa, const uint32_t *b, int sb, double *distance);

/* Batch driver used for CPU-baseline timing: one fresh zeroed 4*s-byte sketch per
 * read (as mash.New does, mash.go:59-65), reads in [offsets[i], offsets[i+1]).
 * variant 0 = faithful, 1 = closed form.  out is n*s words (padded layout).
 * If out == NULL the sketches are allocated, computed and freed per read (pure
 * timing mode; a digest of the per-read word sums is returned through checksum). */
int po_mash_sketch_batch(const uint8_t *bases, const uint64_t *offsets, uint64_t n, int k, int s,
                         int variant, int nthreads, uint32_t *out, uint64_t *checksum);

/* search/align/align.go:171-203 (score + first-max position; traceback is out of
 * scope).  lut_a / lut_b map a byte to its index in the first / second alphabet
 * (-1 = "Symbol not in alphabet", alphabet/alphabet.go:35-41); table is
 * n_a x n_b row-major (search/align/matrix/matrix.go:28-38).
 * err_code: 0 none, 1 symbol of a not in first alphabet, 2 symbol of b not in
 * second alphabet; err_pos = index of the offending byte in that string. */
int po_sw_score(const uint8_t *a, int64_t la, const uint8_t *b, int64_t lb,
                const int16_t *lut_a, const int16_t *lut_b, const int64_t *table, int n_b,
                int64_t gap, int64_t *score, int64_t *max_row, int64_t *max_col,
                int32_t *err_code, int64_t *err_pos);

/* search/align/align.go:171-232 in full (score + the two aligned strings). */
int po_sw_align(const uint8_t *a, int64_t la, const uint8_t *b, int64_t lb, const int16_t *lut_a,
                const int16_t *lut_b, const int64_t *table, int n_b, int64_t gap, int64_t *score,
                uint8_t *out_a, uint8_t *out_b, int64_t cap, int64_t *out_len, int32_t *err_code,
                int64_t *err_pos);

/* search/align/align.go:100-166 in full (score + the two aligned strings). */
int po_nw_align(const uint8_t *a, int64_t la, const uint8_t *b, int64_t lb, const int16_t *lut_a,
                const int16_t *lut_b, const int64_t *table, int n_b, int64_t gap, int64_t *score,
                uint8_t *out_a, uint8_t *out_b, int64_t cap, int64_t *out_len, int32_t *err_code,
                int64_t *err_pos);

/* search/align/align.go:100-166, score only. */
int po_nw_score(const uint8_t *a, int64_t la, const uint8_t *b, int64_t lb,
                const int16_t *lut_a, const int16_t *lut_b, const int64_t *table, int n_b,
                int64_t gap, int64_t *score, int32_t *err_code, int64_t *err_pos);

/* io/fastq/fastq.go:88-99,117-214 over a whole buffer (see the definition for the error codes). */
int po_fastq_parse(const uint8_t *text, uint64_t n, uint64_t *seq_start, uint64_t *seq_len, uint64_t cap,
                   uint64_t *n_records, int32_t *err_code, uint64_t *err_line);

/* io/fasta/fasta.go:72-77,96-118,149-243 (fasta.Parse) over a whole buffer, with the bufio.Reader
 * underneath modelled literally; alias=1 keeps the reference's use of `line` after Peek(1)
 * (fasta.go:192), alias=0 copies the line first.  See fasta_oracle.c for the error codes. */
int po_fasta_parse(const uint8_t *text, uint64_t n, uint32_t max_line_size, int alias, uint8_t *seq,
                   uint64_t seq_cap, uint64_t *seq_off, uint8_t *name, uint64_t name_cap,
                   uint64_t *name_off, uint64_t rec_cap, uint64_t *n_records, int32_t *err_code,
                   uint64_t *err_line);

/* transform/transform.go:15-23,78-109. out has room for len bytes. */
void po_reverse_complement(const uint8_t *seq, int64_t len, uint8_t *out);

/* primers/primers.go:70-105.  Returns PO_PANIC for the empty string
 * (primers.go:89 indexes sequence[len-1]) and PO_UNSUPPORTED if a byte >= 0x80
 * is present (strings.ToUpper would re-encode it). */
int po_santalucia(const uint8_t *seq, int64_t len, double cp, double na, double mg,
                  double *tm, double *dh, double *ds);
/* primers/primers.go:121-128. */
int po_melting_temp(const uint8_t *seq, int64_t len, double *tm);

/* ---- batch drivers (batch_drivers.c): a static parallel-for of per-item calls, for bench.py's
 * CPU legs only.  checksum (may be NULL) receives a digest so that the work cannot be elided. */
int po_sw_score_batch(const uint8_t *queries, const uint64_t *offsets, uint64_t n, const uint8_t *templ, int64_t templ_len,
                      const int16_t *lut_a, const int16_t *lut_b, const int64_t *table, int n_b, int64_t gap, int nthreads,
                      int64_t *score, uint64_t *checksum);
int po_melting_temp_batch(const uint8_t *bases, const uint64_t *offsets, uint64_t n, int nthreads, double *tm, uint64_t *checksum);
int po_mash_similarity_block(const uint32_t *sk, uint64_t n, int s, uint64_t row_lo, uint64_t row_hi, uint64_t col_lo,
                             uint64_t col_hi, int nthreads, uint32_t *same, uint64_t *checksum);

#ifdef __cplusplus
}
#endif
#endif
