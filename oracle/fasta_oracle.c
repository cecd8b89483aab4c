/*
 * oracle/fasta_oracle.c -- CPU restatement of fasta.Parse (io/fasta/fasta.go:72-77 ->
 * Parser.ParseAll / ParseN :96-118 -> Parser.ParseNext :149-243) over an in-memory text.
 *
 * TEST INFRASTRUCTURE ONLY (see poly_oracle.h).
 *
 * The parser is driven by a bufio.Reader (fasta.go:84,91).  Its observable behaviour depends
 * on that reader in two ways that are restated here literally instead of being idealised:
 *
 *   1. bufio.ErrBufferFull: a line whose content is >= the buffer size stops the parse
 *      (fasta.go:178-180; buffer = max(16, maxLineSize), 64 KiB for Parse, fasta.go:74).
 *   2. `line` (the slice ReadSlice returned) aliases the reader's buffer and the parser calls
 *      Peek(1) (fasta.go:192) BEFORE it uses `line` (fasta.go:197,206,208,216).  When the line's
 *      newline was the last buffered byte, Peek refills the buffer from offset 0 and the bytes
 *      `line` points at are replaced by file content one buffer further on.  With a reader that
 *      fills every Read completely (strings.Reader, bytes.Reader, *os.File on a regular file) the
 *      refill points are a deterministic function of the text, so the outcome is too.
 *      `alias` = 1 restates this; `alias` = 0 is the same parser with `line` copied out before
 *      the Peek (what the code means to do).
 *
 * The bufio model below follows the Go 1.21 standard library (go.mod:3) functions
 * (*Reader).fill, ReadSlice, Peek; the reader underneath is strings.Reader (Read copies
 * min(len(p), remaining) bytes, then returns 0, io.EOF).
 */
#include <stdlib.h>
#include <string.h>

#include "poly_oracle.h"

enum { E_NONE = 0, E_EOF = 1, E_FULL = 2 };

typedef struct {
    const uint8_t *src;
    uint64_t n, pos; /* strings.Reader */
    uint8_t *buf;
    uint64_t cap, r, w;
    int err;
} bufio_t;

static void bufio_fill(bufio_t *b) {
    if (b->r > 0) {
        memmove(b->buf, b->buf + b->r, b->w - b->r);
        b->w -= b->r;
        b->r = 0;
    }
    /* one Read: strings.Reader never returns (0, nil) */
    if (b->pos >= b->n) {
        b->err = E_EOF;
        return;
    }
    uint64_t room = b->cap - b->w, left = b->n - b->pos, k = room < left ? room : left;
    memcpy(b->buf + b->w, b->src + b->pos, k);
    b->pos += k;
    b->w += k;
}

static int bufio_read_err(bufio_t *b) {
    int e = b->err;
    b->err = E_NONE;
    return e;
}

/* returns the line as (offset into buf, length); *err as ReadSlice's error */
static void bufio_read_slice(bufio_t *b, uint64_t *off, uint64_t *len, int *err) {
    uint64_t s = 0;
    for (;;) {
        const uint8_t *p = b->w > b->r + s ? memchr(b->buf + b->r + s, '\n', b->w - b->r - s) : NULL;
        if (p) {
            uint64_t i = (uint64_t)(p - (b->buf + b->r));
            *off = b->r; *len = i + 1; *err = E_NONE;
            b->r += i + 1;
            return;
        }
        if (b->err != E_NONE) {
            *off = b->r; *len = b->w - b->r;
            b->r = b->w;
            *err = bufio_read_err(b);
            return;
        }
        if (b->w - b->r >= b->cap) {
            b->r = b->w;
            *off = 0; *len = b->cap; *err = E_FULL;
            return;
        }
        s = b->w - b->r;
        bufio_fill(b);
    }
}

/* Peek(1): returns 1 and *c when a byte is available, else 0 (error consumed) */
static int bufio_peek1(bufio_t *b, uint8_t *c) {
    while (b->w - b->r < 1 && b->w - b->r < b->cap && b->err == E_NONE) bufio_fill(b);
    if (b->w - b->r < 1) {
        (void)bufio_read_err(b);
        return 0;
    }
    *c = b->buf[b->r];
    return 1;
}

/*
 * Error codes (po_fasta_parse and pg_fasta_ingest share them):
 *   0 none
 *   1 "did not find fasta start '>'"           fasta.go:226-228 (when the wrapped error is not EOF)
 *   2 "empty fasta sequence for %q"            fasta.go:229-232
 *   3 "line %d too large for buffer"           fasta.go:178-180
 *   4 bufio.ErrBufferFull returned as is       fasta.go:172-176,243 (over-long ';' line inside a record)
 * *err_line is the line number the message prints (parser.line, or parser.line+1 for code 3;
 * parser.line for code 4, which prints none).
 * Records are appended to seq/name (dense bytes) with seq_off/name_off (n+1 entries each).
 */
int po_fasta_parse(const uint8_t *text, uint64_t n, uint32_t max_line_size, int alias, uint8_t *seq,
                   uint64_t seq_cap, uint64_t *seq_off, uint8_t *name, uint64_t name_cap,
                   uint64_t *name_off, uint64_t rec_cap, uint64_t *n_records, int32_t *err_code,
                   uint64_t *err_line) {
    bufio_t b;
    memset(&b, 0, sizeof b);
    b.src = text;
    b.n = n;
    b.cap = max_line_size < 16 ? 16 : max_line_size; /* bufio.NewReaderSize minimum */
    b.buf = (uint8_t *)malloc(b.cap);
    uint8_t *copy = (uint8_t *)malloc(b.cap);
    if (!b.buf || !copy) { free(b.buf); free(copy); return PO_UNSUPPORTED; }
    uint64_t nrec = 0, seq_len = 0, name_len = 0, parser_line = 0;
    int rc = PO_OK;
    *err_code = 0;
    *err_line = 0;
    seq_off[0] = 0;
    name_off[0] = 0;
    for (;;) { /* ParseN: one ParseNext per iteration (fasta.go:104-116) */
        uint8_t c;
        if (!bufio_peek1(&b, &c)) break; /* fasta.go:150-153: EOF, not an error for ParseN */
        int looking = 1, err = E_NONE;
        const uint64_t seq_begin = seq_len, name_begin = name_len;
        uint64_t cur_name_len = 0;
        int fatal = 0; /* 3: line too large */
        for (;;) {
            uint64_t off, len;
            bufio_read_slice(&b, &off, &len, &err);
            const uint8_t *line = b.buf + off;
            const int skippable = len <= 1 || line[0] == ';'; /* fasta.go:168 */
            parser_line++;
            if (err != E_NONE) { /* fasta.go:172-188 */
                if (skippable) {
                    if (err == E_EOF) err = E_NONE;
                    break;
                } else if (err == E_FULL) {
                    fatal = 3;
                    break;
                }
                if (seq_len + len > seq_cap) { rc = PO_UNSUPPORTED; goto done; }
                memcpy(seq + seq_len, line, len);
                seq_len += len;
                break;
            }
            len -= 1; /* fasta.go:191 */
            if (!alias) { memcpy(copy, line, len); line = copy; }
            uint8_t pk = 0;
            const int have = bufio_peek1(&b, &pk); /* fasta.go:192 -- may overwrite what `line` aliases */
            if (!looking && have && pk == '>') {
                if (!skippable) {
                    if (seq_len + len > seq_cap) { rc = PO_UNSUPPORTED; goto done; }
                    memcpy(seq + seq_len, line, len);
                    seq_len += len;
                }
                break;
            } else if (skippable) {
                continue;
            }
            if (looking) {
                if (line[0] == '>') {
                    cur_name_len = len - 1;
                    if (name_begin + cur_name_len > name_cap) { rc = PO_UNSUPPORTED; goto done; }
                    memcpy(name + name_begin, line + 1, cur_name_len);
                    looking = 0;
                }
                continue;
            }
            if (seq_len + len > seq_cap) { rc = PO_UNSUPPORTED; goto done; }
            memcpy(seq + seq_len, line, len);
            seq_len += len;
        }
        /* fasta.go:225-243 + ParseN's handling of the returned error */
        int stop = 0;
        if (fatal) {
            *err_code = 3; *err_line = parser_line + 1; stop = 1;
        } else if (looking) {
            if (err != E_EOF) { *err_code = 1; *err_line = parser_line; }
            stop = 1; /* wraps err: EOF is swallowed by ParseN */
        } else if (seq_len == seq_begin) {
            *err_code = 2; *err_line = parser_line; stop = 1;
        } else if (err != E_NONE) {
            if (err == E_FULL) { *err_code = 4; *err_line = parser_line; }
            stop = 1; /* fasta returned WITH an error: ParseN drops it */
        }
        if (stop) {
            seq_len = seq_begin;
            name_len = name_begin;
            break;
        }
        if (nrec >= rec_cap) { rc = PO_UNSUPPORTED; goto done; }
        name_len = name_begin + cur_name_len;
        nrec++;
        seq_off[nrec] = seq_len;
        name_off[nrec] = name_len;
    }
done:
    *n_records = nrec;
    free(b.buf);
    free(copy);
    return rc;
}
