/*
 * oracle/mash_oracle.c -- CPU restatement of search/mash (TEST INFRASTRUCTURE ONLY,
 * see poly_oracle.h).  Citations are file:line under /root/reference.
 */
#include "poly_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ---- MurmurHash3_x86_32 (github.com/spaolacci/murmur3 v1.1.0, call site
 * search/mash/mash.go:76).  Published algorithm (Appleby, SMHasher): c1/c2 block
 * mix, 1-3 byte little-endian tail, xor length, fmix32. ---- */
static inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

uint32_t po_murmur3_32(const uint8_t *data, size_t len, uint32_t seed) {
    const uint32_t c1 = 0xcc9e2d51u, c2 = 0x1b873593u;
    uint32_t h = seed;
    size_t nblocks = len / 4;
    for (size_t i = 0; i < nblocks; i++) {
        uint32_t k = (uint32_t)data[4 * i] | ((uint32_t)data[4 * i + 1] << 8) |
                     ((uint32_t)data[4 * i + 2] << 16) | ((uint32_t)data[4 * i + 3] << 24);
        k *= c1;
        k = rotl32(k, 15);
        k *= c2;
        h ^= k;
        h = rotl32(h, 13);
        h = h * 5u + 0xe6546b64u;
    }
    const uint8_t *tail = data + 4 * nblocks;
    uint32_t k1 = 0;
    switch (len & 3) {
    case 3: k1 ^= (uint32_t)tail[2] << 16; /* fallthrough */
    case 2: k1 ^= (uint32_t)tail[1] << 8;  /* fallthrough */
    case 1:
        k1 ^= (uint32_t)tail[0];
        k1 *= c1;
        k1 = rotl32(k1, 15);
        k1 *= c2;
        h ^= k1;
    }
    h ^= (uint32_t)len;
    h ^= h >> 16;
    h *= 0x85ebca6bu;
    h ^= h >> 13;
    h *= 0xc2b2ae35u;
    h ^= h >> 16;
    return h;
}

static int cmp_u32(const void *a, const void *b) {
    uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
    return (x > y) - (x < y);
}

/* search/mash/mash.go:68-104, branch for branch.  sort.Slice over uint32 values
 * is order-equivalent to qsort (no payload, so stability is unobservable). */
int po_mash_sketch_faithful(const uint8_t *seq, int64_t len, int k, int s, uint32_t *sk) {
    if (s < 0) return PO_PANIC;                 /* make([]uint32, s) panics, mash.go:63 */
    int64_t max_shifted = (int64_t)s - 1;       /* mash.go:70 */
    for (int64_t i = 0; i < len - (int64_t)k; i++) { /* mash.go:73: L-k windows, not L-k+1 */
        if (k < 0) return PO_PANIC;             /* sequence[i:i+k] with k<0: slice bounds */
        uint32_t h = po_murmur3_32(seq + i, (size_t)k, 0); /* mash.go:74-76 */
        if (i < max_shifted) {                  /* mash.go:81-84 */
            sk[i] = h;
            continue;
        }
        if (i == max_shifted) {                 /* mash.go:87-92 */
            sk[max_shifted] = h;
            qsort(sk, (size_t)s, sizeof(uint32_t), cmp_u32);
            continue;
        }
        /* mash.go:96: i > max_shifted && Sketches[max_shifted] > hash */
        if (max_shifted < 0) return PO_PANIC;   /* s == 0: Sketches[-1] */
        if (sk[max_shifted] > h) {
            sk[max_shifted] = h;
            if (max_shifted - 1 < 0) return PO_PANIC; /* s == 1: Sketches[-1], mash.go:98 */
            if (h < sk[max_shifted - 1])        /* mash.go:98-100 */
                qsort(sk, (size_t)s, sizeof(uint32_t), cmp_u32);
            continue;
        }
    }
    return PO_OK;
}

/* Closed form: n = max(len-k,0) hashes; n >= s -> ascending
 * bottom-s multiset; n < s -> first n slots positional, the rest untouched. */
int po_mash_sketch_closed(const uint8_t *seq, int64_t len, int k, int s, uint32_t *sk) {
    if (s < 0) return PO_PANIC;
    int64_t n = len - (int64_t)k;
    if (n <= 0) return PO_OK;
    if (k < 0) return PO_PANIC;
    if (s == 0) return PO_PANIC;
    if (n < s) {
        for (int64_t i = 0; i < n; i++) sk[i] = po_murmur3_32(seq + i, (size_t)k, 0);
        return PO_OK;
    }
    uint32_t *h = (uint32_t *)malloc((size_t)n * sizeof(uint32_t));
    if (!h) return -100;
    for (int64_t i = 0; i < n; i++) h[i] = po_murmur3_32(seq + i, (size_t)k, 0);
    if (s == 1) {
        /* mash.go:96-98: the first later hash below Sketches[0] indexes Sketches[-1]. */
        for (int64_t i = 1; i < n; i++)
            if (h[i] < h[0]) {
                sk[0] = h[i];
                free(h);
                return PO_PANIC;
            }
        sk[0] = h[0];
        free(h);
        return PO_OK;
    }
    qsort(h, (size_t)n, sizeof(uint32_t), cmp_u32);
    memcpy(sk, h, (size_t)s * sizeof(uint32_t));
    free(h);
    return PO_OK;
}

/* search/mash/mash.go:107-135. */
int po_mash_similarity(const uint32_t *a, int sa, const uint32_t *b, int sb, int64_t *same_out,
                       double *similarity) {
    const uint32_t *larger = a, *smaller = b; /* mash.go:109-110 */
    int sl = sa, ss = sb;
    if (sa < sb) {                            /* mash.go:112-115 */
        larger = b; sl = sb;
        smaller = a; ss = sa;
    }
    if (sl < 1 || ss < 1) return PO_PANIC;    /* Sketches[SketchSize-1] / Sketches[0] */
    int64_t same = 0;
    if (larger[sl - 1] < smaller[0] || smaller[ss - 1] < larger[0]) { /* mash.go:117-119 */
        if (same_out) *same_out = 0;
        if (similarity) *similarity = 0.0;
        return PO_OK;
    }
    int si = 0, li = 0;
    while (si < ss && li < sl) {              /* mash.go:121-132 */
        if (smaller[si] == larger[li]) {
            same++; si++; li++;
        } else if (smaller[si] < larger[li]) {
            si++;
        } else {
            li++;
        }
    }
    if (same_out) *same_out = same;
    if (similarity) *similarity = (double)same / (double)ss; /* mash.go:134 */
    return PO_OK;
}

int po_mash_distance(const uint32_t *a, int sa, const uint32_t *b, int sb, double *distance) {
    double sim = 0.0;
    int rc = po_mash_similarity(a, sa, b, sb, NULL, &sim);
    if (rc != PO_OK) return rc;
    *distance = 1 - sim;                      /* mash.go:139 */
    return PO_OK;
}

/* ---- batch driver for CPU-baseline timing (static parallel-for over reads; the
 * reference itself has no goroutines on this path, SURVEY "READ THIS FIRST") ---- */
typedef struct {
    const uint8_t *bases;
    const uint64_t *offsets;
    uint64_t lo, hi;
    int k, s, variant;
    uint32_t *out;
    uint64_t fnv;
    int rc;
} batch_job;

static void *batch_worker(void *p) {
    batch_job *j = (batch_job *)p;
    uint64_t fnv = 0xcbf29ce484222325ull;
    j->rc = PO_OK;
    for (uint64_t i = j->lo; i < j->hi; i++) {
        uint32_t *sk;
        if (j->out) {
            sk = j->out + i * (uint64_t)j->s;
            memset(sk, 0, (size_t)j->s * 4);
        } else {
            sk = (uint32_t *)calloc((size_t)(j->s > 0 ? j->s : 1), 4); /* mash.New, mash.go:59-65 */
            if (!sk) { j->rc = -100; return NULL; }
        }
        const uint8_t *seq = j->bases + j->offsets[i];
        int64_t len = (int64_t)(j->offsets[i + 1] - j->offsets[i]);
        int rc = j->variant == 0 ? po_mash_sketch_faithful(seq, len, j->k, j->s, sk)
                                 : po_mash_sketch_closed(seq, len, j->k, j->s, sk);
        if (rc != PO_OK) j->rc = rc;
        if (!j->out) { /* touch two words so the work is not optimised away (no per-word digest: the reference does none) */
            const uint64_t acc = j->s > 0 ? (uint64_t)sk[0] ^ ((uint64_t)sk[j->s - 1] << 32) : 0;
            fnv = (fnv ^ acc) * 0x100000001b3ull;
            free(sk);
        }
    }
    j->fnv = fnv;
    return NULL;
}

int po_mash_sketch_batch(const uint8_t *bases, const uint64_t *offsets, uint64_t n, int k, int s,
                         int variant, int nthreads, uint32_t *out, uint64_t *checksum) {
    if (nthreads < 1) nthreads = 1;
    if ((uint64_t)nthreads > n && n > 0) nthreads = (int)n;
    batch_job *jobs = (batch_job *)calloc((size_t)nthreads, sizeof(batch_job));
    pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
    if (!jobs || !th) return -100;
    for (int t = 0; t < nthreads; t++) {
        jobs[t] = (batch_job){bases, offsets, n * (uint64_t)t / (uint64_t)nthreads,
                              n * (uint64_t)(t + 1) / (uint64_t)nthreads, k, s, variant, out, 0, 0};
        if (nthreads == 1) batch_worker(&jobs[t]);
        else pthread_create(&th[t], NULL, batch_worker, &jobs[t]);
    }
    int rc = PO_OK;
    uint64_t x = 0;
    for (int t = 0; t < nthreads; t++) {
        if (nthreads > 1) pthread_join(th[t], NULL);
        if (jobs[t].rc != PO_OK) rc = jobs[t].rc;
        x ^= jobs[t].fnv * (uint64_t)(2 * t + 1);
    }
    if (checksum) *checksum = x;
    free(jobs);
    free(th);
    return rc;
}
