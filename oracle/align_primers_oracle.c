/*
 * oracle/align_primers_oracle.c -- CPU restatement of search/align (score part)
 * and primers.SantaLucia (TEST INFRASTRUCTURE ONLY, see poly_oracle.h).
 * Citations are file:line under /root/reference.
 */
#include "poly_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

static inline int64_t max64(int64_t a, int64_t b) { return a > b ? a : b; } /* align.go:243-248 */

/* scoring.Score(a,b) -> SubstitutionMatrix.Score: FirstAlphabet.Encode(a) is tried
 * before SecondAlphabet.Encode(b) (search/align/matrix/matrix.go:28-38). */
static inline int cell_score(uint8_t ca, uint8_t cb, const int16_t *lut_a, const int16_t *lut_b,
                             const int64_t *table, int n_b, int64_t *out) {
    int ia = lut_a[ca];
    if (ia < 0) return 1;
    int ib = lut_b[cb];
    if (ib < 0) return 2;
    *out = table[(size_t)ia * (size_t)n_b + (size_t)ib];
    return 0;
}

/* search/align/align.go:171-203.  Two rolling rows instead of the full matrix
 * (the traceback that needs it is out of scope); same visiting order, same
 * strict-> running maximum. */
int po_sw_score(const uint8_t *a, int64_t la, const uint8_t *b, int64_t lb, const int16_t *lut_a,
                const int16_t *lut_b, const int64_t *table, int n_b, int64_t gap, int64_t *score,
                int64_t *max_row, int64_t *max_col, int32_t *err_code, int64_t *err_pos) {
    int64_t *prev = (int64_t *)calloc((size_t)lb + 1, sizeof(int64_t));
    int64_t *cur = (int64_t *)calloc((size_t)lb + 1, sizeof(int64_t));
    if (!prev || !cur) { free(prev); free(cur); return -100; }
    int64_t best = 0, br = 0, bc = 0; /* align.go:181-183 */
    *err_code = 0;
    *err_pos = -1;
    for (int64_t i = 1; i <= la; i++) {            /* align.go:186 */
        cur[0] = 0;
        for (int64_t j = 1; j <= lb; j++) {        /* align.go:187 */
            int64_t m = 0;
            int e = cell_score(a[i - 1], b[j - 1], lut_a, lut_b, table, n_b, &m); /* :188 */
            if (e) {                               /* align.go:189-191 -> (0,"","",err) */
                *err_code = e;
                *err_pos = e == 1 ? i - 1 : j - 1;
                *score = 0;
                if (max_row) *max_row = 0;
                if (max_col) *max_col = 0;
                free(prev); free(cur);
                return PO_OK;
            }
            int64_t diag = prev[j - 1] + m;        /* align.go:192 */
            int64_t up = prev[j] + gap;            /* align.go:193 */
            int64_t left = cur[j - 1] + gap;       /* align.go:194 */
            int64_t v = max64(0, max64(diag, max64(up, left))); /* align.go:195 */
            cur[j] = v;
            if (v > best) { best = v; br = i; bc = j; }         /* align.go:197-201 */
        }
        int64_t *t = prev; prev = cur; cur = t;
    }
    *score = best;
    if (max_row) *max_row = br;
    if (max_col) *max_col = bc;
    free(prev); free(cur);
    return PO_OK;
}

/* search/align/align.go:100-134 (fill) and :166 (returned score). */
int po_nw_score(const uint8_t *a, int64_t la, const uint8_t *b, int64_t lb, const int16_t *lut_a,
                const int16_t *lut_b, const int64_t *table, int n_b, int64_t gap, int64_t *score,
                int32_t *err_code, int64_t *err_pos) {
    int64_t *prev = (int64_t *)calloc((size_t)lb + 1, sizeof(int64_t));
    int64_t *cur = (int64_t *)calloc((size_t)lb + 1, sizeof(int64_t));
    if (!prev || !cur) { free(prev); free(cur); return -100; }
    *err_code = 0;
    *err_pos = -1;
    for (int64_t j = 1; j <= lb; j++) prev[j] = prev[j - 1] + gap; /* align.go:120-122 */
    for (int64_t i = 1; i <= la; i++) {
        cur[0] = prev[0] + gap;                                    /* align.go:115-117 */
        for (int64_t j = 1; j <= lb; j++) {
            int64_t m = 0;
            int e = cell_score(a[i - 1], b[j - 1], lut_a, lut_b, table, n_b, &m);
            if (e) {
                *err_code = e;
                *err_pos = e == 1 ? i - 1 : j - 1;
                *score = 0;
                free(prev); free(cur);
                return PO_OK;
            }
            cur[j] = max64(prev[j - 1] + m, max64(prev[j] + gap, cur[j - 1] + gap)); /* :132-135 */
        }
        int64_t *t = prev; prev = cur; cur = t;
    }
    *score = prev[lb];
    free(prev); free(cur);
    return PO_OK;
}

/* transform/transform.go:78-109: complementTable; unlisted bytes map to 0. */
static uint8_t comp_table[256];
static int comp_ready = 0;
static void comp_init(void) {
    if (comp_ready) return;
    const char *from = "ABCDGHKMNRSTVWYabcdghkmnrstvwy";
    const char *to = "TVGHCDMKNYSABWRtvghcdmknysabwr";
    memset(comp_table, 0, sizeof comp_table);
    for (int i = 0; from[i]; i++) comp_table[(uint8_t)from[i]] = (uint8_t)to[i];
    comp_ready = 1;
}

/* transform/transform.go:15-23. */
void po_reverse_complement(const uint8_t *seq, int64_t len, uint8_t *out) {
    comp_init();
    for (int64_t i = 0; i < len; i++) out[i] = comp_table[seq[len - i - 1]];
}

/* primers/primers.go:42-59: nearest-neighbour table; absent keys read as {0,0}
 * (Go map zero value, primers.go:98). */
static int nn_lookup(uint8_t x, uint8_t y, double *H, double *S) {
    static const struct { char a, b; double H, S; } nn[16] = {
        {'A', 'A', -7.6, -21.3}, {'T', 'T', -7.6, -21.3}, {'A', 'T', -7.2, -20.4},
        {'T', 'A', -7.2, -21.3}, {'C', 'A', -8.5, -22.7}, {'T', 'G', -8.5, -22.7},
        {'G', 'T', -8.4, -22.4}, {'A', 'C', -8.4, -22.4}, {'C', 'T', -7.8, -21.0},
        {'A', 'G', -7.8, -21.0}, {'G', 'A', -8.2, -22.2}, {'T', 'C', -8.2, -22.2},
        {'C', 'G', -10.6, -27.2}, {'G', 'C', -9.8, -24.4}, {'G', 'G', -8.0, -19.9},
        {'C', 'C', -8.0, -19.9}};
    for (int i = 0; i < 16; i++)
        if ((uint8_t)nn[i].a == x && (uint8_t)nn[i].b == y) {
            *H = nn[i].H;
            *S = nn[i].S;
            return 1;
        }
    *H = 0.0;
    *S = 0.0;
    return 0;
}

/* primers/primers.go:70-105, statement for statement (f64 accumulation order kept). */
int po_santalucia(const uint8_t *seq_in, int64_t len, double cp, double na, double mg, double *tm,
                  double *dh_out, double *ds_out) {
    if (len <= 0) return PO_PANIC; /* primers.go:89: sequence[len(sequence)-1] */
    uint8_t *seq = (uint8_t *)malloc((size_t)len);
    uint8_t *rc = (uint8_t *)malloc((size_t)len);
    if (!seq || !rc) { free(seq); free(rc); return -100; }
    for (int64_t i = 0; i < len; i++) { /* strings.ToUpper, primers.go:71 (ASCII domain) */
        uint8_t c = seq_in[i];
        if (c >= 0x80) { free(seq); free(rc); return PO_UNSUPPORTED; }
        seq[i] = (c >= 'a' && c <= 'z') ? (uint8_t)(c - 32) : c;
    }
    const double gas_constant = 1.9872; /* primers.go:73 */
    double symmetry;
    double dH = 0.0, dS = 0.0;
    dH += 0.2;  /* primers.go:78 */
    dS += -5.7; /* primers.go:79 */
    po_reverse_complement(seq, len, rc);
    if (memcmp(seq, rc, (size_t)len) == 0) { /* primers.go:81-87 */
        dH += 0.0;
        dS += -1.4;
        symmetry = 1;
    } else {
        symmetry = 4;
    }
    if (seq[len - 1] == 'A' || seq[len - 1] == 'T') { /* primers.go:89-92 */
        dH += 2.2;
        dS += 6.9;
    }
    double salt_effect = na + (mg * 140);                   /* primers.go:94 */
    dS += (0.368 * (double)(len - 1) * log(salt_effect));   /* primers.go:95 */
    for (int64_t i = 0; i + 1 < len; i++) {                 /* primers.go:97-101 */
        double H, S;
        nn_lookup(seq[i], seq[i + 1], &H, &S);
        dH += H;
        dS += S;
    }
    *tm = dH * 1000 / (dS + gas_constant * log(cp / symmetry)) - 273.15; /* primers.go:103 */
    if (dh_out) *dh_out = dH;
    if (ds_out) *ds_out = dS;
    free(seq);
    free(rc);
    return PO_OK;
}

/* primers/primers.go:121-128. */
int po_melting_temp(const uint8_t *seq, int64_t len, double *tm) {
    return po_santalucia(seq, len, 500e-9, 50e-3, 0.0, tm, NULL, NULL);
}

/* search/align/align.go:171-232 in full: fill, first-max position, traceback (diag > up > left
 * preference, strings built by prepending).  out_a/out_b receive the aligned strings (no NUL),
 * *out_len their common length; cap = capacity of each buffer (la + lb always suffices).
 * Only used to pin the traceback kernel; O(la*lb) memory. */
int po_sw_align(const uint8_t *a, int64_t la, const uint8_t *b, int64_t lb, const int16_t *lut_a,
                const int16_t *lut_b, const int64_t *table, int n_b, int64_t gap, int64_t *score,
                uint8_t *out_a, uint8_t *out_b, int64_t cap, int64_t *out_len, int32_t *err_code,
                int64_t *err_pos) {
    const int64_t W = lb + 1;
    int64_t *m = (int64_t *)calloc((size_t)((la + 1) * W), sizeof(int64_t)); /* align.go:175-178 */
    if (!m) return -100;
    int64_t best = 0, br = 0, bc = 0;
    *err_code = 0; *err_pos = -1; *out_len = 0;
    for (int64_t i = 1; i <= la; i++)
        for (int64_t j = 1; j <= lb; j++) {
            int64_t s = 0;
            int e = cell_score(a[i - 1], b[j - 1], lut_a, lut_b, table, n_b, &s);
            if (e) { *err_code = e; *err_pos = e == 1 ? i - 1 : j - 1; *score = 0; free(m); return PO_OK; }
            int64_t v = max64(0, max64(m[(i - 1) * W + j - 1] + s, max64(m[(i - 1) * W + j] + gap, m[i * W + j - 1] + gap)));
            m[i * W + j] = v;
            if (v > best) { best = v; br = i; bc = j; }            /* align.go:197-201 */
        }
    /* traceback, align.go:205-229: build reversed, then reverse (== prepending) */
    int64_t n = 0, i = br, j = bc;
    while (m[i * W + j] > 0) {
        int64_t s = 0;
        cell_score(a[i - 1], b[j - 1], lut_a, lut_b, table, n_b, &s);
        if (n >= cap) { free(m); return -101; }
        if (m[i * W + j] == m[(i - 1) * W + j - 1] + s) { out_a[n] = a[i - 1]; out_b[n] = b[j - 1]; i--; j--; }
        else if (m[i * W + j] == m[(i - 1) * W + j] + gap) { out_a[n] = a[i - 1]; out_b[n] = '-'; i--; }
        else if (m[i * W + j] == m[i * W + j - 1] + gap) { out_a[n] = '-'; out_b[n] = b[j - 1]; j--; }
        else { free(m); return -102; } /* the reference would spin forever; cannot happen for a max of the three */
        n++;
    }
    for (int64_t t = 0; t < n / 2; t++) {
        uint8_t x = out_a[t]; out_a[t] = out_a[n - 1 - t]; out_a[n - 1 - t] = x;
        x = out_b[t]; out_b[t] = out_b[n - 1 - t]; out_b[n - 1 - t] = x;
    }
    *out_len = n;
    *score = best;
    free(m);
    return PO_OK;
}


/* io/fastq/fastq.go:88-99 (ParseN) over :117-214 (ParseNext) for a whole in-memory buffer, with the
 * 64 KiB reader of Parse/Read (fastq.go:54-58).  Records before the first rejected one are
 * returned; err_code: 0 none, 1 line not newline terminated (io.EOF inside a record), 2 empty
 * sequence (:179-181), 3 empty quality (:197-199), 4 no '@' (:203-205), 5 the reference panics
 * (empty identifier line :160, or an optional datum without '=' :166-169), 6 line longer than the
 * reader buffer (bufio.ErrBufferFull).  err_line = parser.line when the reference returns. */
int po_fastq_parse(const uint8_t *text, uint64_t n, uint64_t *seq_start, uint64_t *seq_len, uint64_t cap,
                   uint64_t *n_records, int32_t *err_code, uint64_t *err_line) {
    uint64_t pos = 0, line = 0, count = 0;
    *err_code = 0; *err_line = 0;
    while (pos < n) {                                   /* Peek(1) succeeds, fastq.go:118 */
        uint64_t lb[4], le[4];
        int e = 0, no_at = 0;
        for (int l = 0; l < 4 && !e; l++) {
            uint64_t q = pos;
            while (q < n && text[q] != '\n') q++;
            line++;                                     /* parser.line++ precedes the error check */
            if (q >= n) { e = (n - pos >= 2 * 32 * 1024) ? 6 : 1; break; } /* ReadSlice: the 64 KiB buffer fills before EOF is seen -> ErrBufferFull, else io.EOF */
            if (q + 1 - pos > 2 * 32 * 1024) { e = 6; break; }
            lb[l] = pos; le[l] = q; pos = q + 1;
            if (l == 0) {
                if (le[0] == lb[0]) { e = 5; break; }   /* string(line)[0] on "" */
                no_at = text[lb[0]] != '@';
                uint64_t token = 0; int has_eq = 0;
                for (uint64_t p = lb[0]; p <= le[0]; p++) {
                    uint8_t c = p < le[0] ? text[p] : (uint8_t)' ';
                    if (c == ' ') { if (token >= 1 && !has_eq) { e = 5; break; } token++; has_eq = 0; }
                    else if (c == '=') has_eq = 1;
                }
            } else if (l == 1) {
                if (le[1] == lb[1]) e = 2;
            } else if (l == 3) {
                if (le[3] == lb[3]) e = 3;
            }
        }
        if (!e && no_at) e = 4;
        if (e) { *err_code = e; *err_line = line; break; }
        if (count < cap) { seq_start[count] = lb[1]; seq_len[count] = le[1] - lb[1]; }
        count++;
    }
    *n_records = count;
    return PO_OK;
}

/* search/align/align.go:100-166 in full: fill, traceback from (la, lb) while BOTH indices are
 * positive (align.go:141: the loop stops as soon as one string is exhausted, the leading rest of
 * the other string is not emitted), diagonal > up > else-left (align.go:146-159), strings built by
 * appending and reversed at the end (align.go:162-164). */
int po_nw_align(const uint8_t *a, int64_t la, const uint8_t *b, int64_t lb, const int16_t *lut_a,
                const int16_t *lut_b, const int64_t *table, int n_b, int64_t gap, int64_t *score,
                uint8_t *out_a, uint8_t *out_b, int64_t cap, int64_t *out_len, int32_t *err_code,
                int64_t *err_pos) {
    const int64_t W = lb + 1;
    int64_t *m = (int64_t *)calloc((size_t)((la + 1) * W), sizeof(int64_t));
    if (!m) return -100;
    *err_code = 0; *err_pos = -1; *out_len = 0;
    for (int64_t i = 1; i <= la; i++) m[i * W] = m[(i - 1) * W] + gap;   /* align.go:115-117 */
    for (int64_t j = 1; j <= lb; j++) m[j] = m[j - 1] + gap;             /* align.go:120-122 */
    for (int64_t i = 1; i <= la; i++)
        for (int64_t j = 1; j <= lb; j++) {
            int64_t s = 0;
            int e = cell_score(a[i - 1], b[j - 1], lut_a, lut_b, table, n_b, &s);
            if (e) { *err_code = e; *err_pos = e == 1 ? i - 1 : j - 1; *score = 0; free(m); return PO_OK; }
            m[i * W + j] = max64(m[(i - 1) * W + j - 1] + s, max64(m[(i - 1) * W + j] + gap, m[i * W + j - 1] + gap));
        }
    int64_t n = 0, i = la, j = lb;
    while (i > 0 && j > 0) {                                             /* align.go:141 */
        int64_t s = 0;
        cell_score(a[i - 1], b[j - 1], lut_a, lut_b, table, n_b, &s);
        if (n >= cap) { free(m); return -101; }
        if (m[i * W + j] == m[(i - 1) * W + j - 1] + s) { out_a[n] = a[i - 1]; out_b[n] = b[j - 1]; i--; j--; }
        else if (m[i * W + j] == m[(i - 1) * W + j] + gap) { out_a[n] = a[i - 1]; out_b[n] = '-'; i--; }
        else { out_a[n] = '-'; out_b[n] = b[j - 1]; j--; }
        n++;
    }
    for (int64_t t = 0; t < n / 2; t++) {
        uint8_t x = out_a[t]; out_a[t] = out_a[n - 1 - t]; out_a[n - 1 - t] = x;
        x = out_b[t]; out_b[t] = out_b[n - 1 - t]; out_b[n - 1 - t] = x;
    }
    *out_len = n;
    *score = m[la * W + lb];
    free(m);
    return PO_OK;
}
