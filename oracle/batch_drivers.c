/*
 * oracle/batch_drivers.c -- parallel-for drivers over the restated functions, used ONLY by
 * bench.py's cpu_baseline / --impl reference legs (TEST INFRASTRUCTURE, see poly_oracle.h).
 * The reference has no batch API and no goroutines on this path: a user loops over the items and
 * calls the function once per item, so that is what each worker does here, over a static split of
 * the items across `nthreads` host threads.
 */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "poly_oracle.h"

typedef struct {
    int kind; /* 0 = SW score, 1 = Tm, 2 = similarity rows */
    uint64_t lo, hi;
    /* SW / Tm */
    const uint8_t *items;
    const uint64_t *offsets;
    const uint8_t *templ;
    int64_t templ_len;
    const int16_t *lut_a, *lut_b;
    const int64_t *table;
    int n_b;
    int64_t gap;
    int64_t *score;
    double *tm;
    /* similarity */
    const uint32_t *sk;
    uint64_t n;
    int s;
    uint64_t col_lo, col_hi;
    uint32_t *same;
    uint64_t acc;
    int rc;
} drv_job;

static void *drv_worker(void *p) {
    drv_job *j = (drv_job *)p;
    j->rc = PO_OK;
    j->acc = 0;
    for (uint64_t i = j->lo; i < j->hi; i++) {
        if (j->kind == 0) { /* align.SmithWaterman(primer, template, scoring), align.go:171-203 */
            int64_t sc = 0, mr = 0, mc = 0, ep = 0;
            int32_t ec = 0;
            int rc = po_sw_score(j->items + j->offsets[i], (int64_t)(j->offsets[i + 1] - j->offsets[i]), j->templ, j->templ_len,
                                 j->lut_a, j->lut_b, j->table, j->n_b, j->gap, &sc, &mr, &mc, &ec, &ep);
            if (rc != PO_OK) j->rc = rc;
            if (j->score) j->score[i] = sc;
            j->acc += (uint64_t)sc;
        } else if (j->kind == 1) { /* primers.MeltingTemp, primers.go:121-128 */
            double tm = 0;
            int rc = po_melting_temp(j->items + j->offsets[i], (int64_t)(j->offsets[i + 1] - j->offsets[i]), &tm);
            if (rc != PO_OK) j->rc = rc;
            if (j->tm) j->tm[i] = tm;
            j->acc += (uint64_t)(tm > 0);
        } else { /* row i: a.Similarity(b) for every column b, mash.go:107-135 */
            for (uint64_t c = j->col_lo; c < j->col_hi; c++) {
                int64_t same = 0;
                double sim = 0;
                int rc = po_mash_similarity(j->sk + i * (uint64_t)j->s, j->s, j->sk + c * (uint64_t)j->s, j->s, &same, &sim);
                if (rc != PO_OK) j->rc = rc;
                if (j->same) j->same[(i - j->lo) * (j->col_hi - j->col_lo) + (c - j->col_lo)] = (uint32_t)same;
                j->acc += (uint64_t)same;
            }
        }
    }
    return NULL;
}

static int drv_run(drv_job proto, uint64_t lo, uint64_t hi, int nthreads, uint64_t *acc_out) {
    const uint64_t n = hi - lo;
    if (nthreads < 1) nthreads = 1;
    if ((uint64_t)nthreads > n && n > 0) nthreads = (int)n;
    drv_job *jobs = (drv_job *)calloc((size_t)nthreads, sizeof(drv_job));
    pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
    if (!jobs || !th) return -100;
    for (int t = 0; t < nthreads; t++) {
        jobs[t] = proto;
        jobs[t].lo = lo + n * (uint64_t)t / (uint64_t)nthreads;
        jobs[t].hi = lo + n * (uint64_t)(t + 1) / (uint64_t)nthreads;
        if (proto.kind == 2 && proto.same) /* each worker writes its own row range */
            jobs[t].same = proto.same + (jobs[t].lo - lo) * (proto.col_hi - proto.col_lo);
        if (nthreads == 1) drv_worker(&jobs[t]);
        else pthread_create(&th[t], NULL, drv_worker, &jobs[t]);
    }
    int rc = PO_OK;
    uint64_t acc = 0;
    for (int t = 0; t < nthreads; t++) {
        if (nthreads > 1) pthread_join(th[t], NULL);
        if (jobs[t].rc != PO_OK) rc = jobs[t].rc;
        acc += jobs[t].acc;
    }
    if (acc_out) *acc_out = acc;
    free(jobs);
    free(th);
    return rc;
}

int po_sw_score_batch(const uint8_t *queries, const uint64_t *offsets, uint64_t n, const uint8_t *templ, int64_t templ_len,
                      const int16_t *lut_a, const int16_t *lut_b, const int64_t *table, int n_b, int64_t gap, int nthreads,
                      int64_t *score, uint64_t *checksum) {
    drv_job j;
    memset(&j, 0, sizeof j);
    j.kind = 0; j.items = queries; j.offsets = offsets; j.templ = templ; j.templ_len = templ_len;
    j.lut_a = lut_a; j.lut_b = lut_b; j.table = table; j.n_b = n_b; j.gap = gap; j.score = score;
    return drv_run(j, 0, n, nthreads, checksum);
}

int po_melting_temp_batch(const uint8_t *bases, const uint64_t *offsets, uint64_t n, int nthreads, double *tm, uint64_t *checksum) {
    drv_job j;
    memset(&j, 0, sizeof j);
    j.kind = 1; j.items = bases; j.offsets = offsets; j.tm = tm;
    return drv_run(j, 0, n, nthreads, checksum);
}

/* rows [row_lo,row_hi) x columns [col_lo,col_hi) of the matching-count matrix over n sketches of size s */
int po_mash_similarity_block(const uint32_t *sk, uint64_t n, int s, uint64_t row_lo, uint64_t row_hi, uint64_t col_lo,
                             uint64_t col_hi, int nthreads, uint32_t *same, uint64_t *checksum) {
    if (row_hi > n || col_hi > n) return -101;
    drv_job j;
    memset(&j, 0, sizeof j);
    j.kind = 2; j.sk = sk; j.n = n; j.s = s; j.col_lo = col_lo; j.col_hi = col_hi; j.same = same;
    return drv_run(j, row_lo, row_hi, nthreads, checksum);
}
