// fasta_ingest.cu -- the step BEFORE the hot path: FASTA text -> dense
// sequences + offsets (and names + offsets), on the GPU.
//
// Mirrors fasta.Parse = NewParser(r, maxLineSize).ParseAll() of
// /root/reference/io/fasta/fasta.go:72-77,96-118,149-243 for a whole in-memory text.  ParseNext is a
// two-state machine over lines (looking for a name / inside a record) whose transitions depend on
// the line itself and on one peeked byte of the next line (fasta.go:192-219):
//
//     from LOOKING:  a non-skippable line whose first byte is '>'      -> INSIDE  (this is the name)
//     from INSIDE :  the byte after this line's newline is '>'         -> LOOKING (record ends here)
//
// so the state BEFORE every line is the exclusive scan of per-line transition functions under
// function composition -- done as a parallel ordered scan instead of a serial walk.  Everything
// else follows from that state: header lines, appended lines (INSIDE and not skippable), record
// ends.  Sequences are the concatenation of the appended lines, so the dense output is one
// exclusive scan of appended lengths plus a copy.
//
// The reference reads through a bufio.Reader, which has two observable effects restated here
// (DESIGN.md "FASTA ingest"):
//   * a line with >= max(16, maxLineSize) content bytes stops the parse (bufio.ErrBufferFull);
//   * PG_FASTA_BUFIO_ALIAS: `line` aliases the reader's buffer and is used after Peek(1); when the
//     line's newline is the last byte of a full buffer, Peek refills the buffer and the line's
//     bytes are replaced by the text one buffer size further on (clipped to what the refill read).
//     With a reader that fills every Read (strings.Reader, bytes.Reader, *os.File) the refill
//     points are a chain F' = 1 + (last newline < F + B): the successor of every line start is
//     computed in parallel, one thread then follows the chain from line 0 (nbytes / B dependent
//     loads) and flags the affected lines; their "effective" bytes are used for the name test, the name
//     and the appended sequence exactly as the reference would.
#include "text_scan.cuh"

namespace pg {

namespace {

enum : int32_t { FA_OK = 0, FA_ERR_NO_START = 1, FA_ERR_EMPTY_SEQ = 2, FA_ERR_LINE_TOO_LONG = 3, FA_ERR_BUFFER_FULL = 4 };

// per-line flag bits
constexpr uint8_t LF_SKIP = 1, LF_NAME = 2, LF_PEEK_GT = 4, LF_CORRUPT = 8;
constexpr uint32_t ST_LOOKING = 0, ST_INSIDE = 1;

// transition functions over {LOOKING, INSIDE}: bit s of the code = image of state s
struct StateOp {
    using In = uint8_t;  // per-line flags
    using T = uint32_t;
    __device__ static T identity() { return 0b10u; }
    __device__ static T lift(In f) {
        const uint32_t from_looking = (f & LF_NAME) ? ST_INSIDE : ST_LOOKING;
        const uint32_t from_inside = (f & LF_PEEK_GT) ? ST_LOOKING : ST_INSIDE;
        return from_looking | (from_inside << 1);
    }
    __device__ static T combine(T first, T then) {
        return ((then >> (first & 1u)) & 1u) | (((then >> ((first >> 1) & 1u)) & 1u) << 1);
    }
};

struct LineView {
    uint64_t b, e;  // content [b, e), newline at e
};
__device__ __forceinline__ LineView line_view(const uint64_t *__restrict__ nl, uint64_t i) {
    LineView v;
    v.b = i == 0 ? 0 : nl[i - 1] + 1;
    v.e = nl[i];
    return v;
}

// byte j of line i as the reference sees it after Peek(1) (fasta.go:192)
__device__ __forceinline__ uint8_t effective_byte(const uint8_t *__restrict__ text, uint64_t nbytes, uint64_t bufsz,
                                                  const LineView &v, bool corrupt, uint64_t j) {
    if (corrupt) {
        const uint64_t refill_base = v.e + 1;                        // file offset the refilled buffer starts at
        const uint64_t refill_len = min(bufsz, nbytes - refill_base);  // bytes the refill read
        const uint64_t pos = v.b - (refill_base - bufsz) + j;        // position of the byte inside the buffer
        if (pos < refill_len) return __ldg(text + refill_base + pos);
    }
    return __ldg(text + v.b + j);
}

// bufio refill chain (alias mode).  A buffer that starts at line i holds text[b_i, b_i + B); the
// next buffer starts at line next(i) = number of newlines below b_i + B.  next() of EVERY line is
// computed in parallel (one binary search each); bit 31 marks "the newline of line next(i) - 1 is
// the last byte of the buffer" (that line is seen corrupted), next(i) == i ends the chain (no full
// buffer with text behind it, or a line longer than the buffer: the parse stops there).
constexpr uint32_t NEXT_LAST_BYTE = 0x80000000u;

__global__ void __launch_bounds__(256)
refill_next_kernel(const uint64_t *__restrict__ nl, uint64_t n_lines, uint64_t nbytes, uint64_t bufsz,
                   uint32_t *__restrict__ next) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_lines) return;
    const uint64_t start = i == 0 ? 0 : nl[i - 1] + 1;
    uint32_t v = (uint32_t)i;
    if (start + bufsz < nbytes) {  // a full buffer with text left behind it
        const uint64_t end = start + bufsz;
        uint64_t lo = i, hi = min(i + bufsz, n_lines);  // number of newlines below `end`
        while (lo < hi) {
            const uint64_t mid = (lo + hi) >> 1;
            if (nl[mid] < end) lo = mid + 1; else hi = mid;
        }
        if (lo > i) v = (uint32_t)lo | (nl[lo - 1] == end - 1 ? NEXT_LAST_BYTE : 0u);
    }
    next[i] = v;
}

// one thread follows the chain from line 0: one dependent load per refill
__global__ void refill_chain_kernel(const uint32_t *__restrict__ next, uint64_t n_lines, uint8_t *__restrict__ corrupt) {
    uint32_t idx = 0;
    while (idx < n_lines) {
        const uint32_t v = next[idx], j = v & ~NEXT_LAST_BYTE;
        if (j == idx) break;
        if (v & NEXT_LAST_BYTE) corrupt[j - 1] = 1;
        idx = j;
    }
}

__global__ void __launch_bounds__(256)
line_flags_kernel(const uint8_t *__restrict__ text, uint64_t nbytes, const uint64_t *__restrict__ nl, uint64_t n_lines,
                  uint64_t bufsz, const uint8_t *__restrict__ corrupt, uint8_t *__restrict__ flags,
                  unsigned long long *__restrict__ first_long) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_lines) return;
    const LineView v = line_view(nl, i);
    const uint64_t len = v.e - v.b;
    const bool skip = len == 0 || __ldg(text + v.b) == ';';  // fasta.go:168 (before the Peek)
    const bool cor = corrupt && corrupt[i];
    uint8_t f = 0;
    if (skip) f |= LF_SKIP;
    if (cor) f |= LF_CORRUPT;
    if (!skip && effective_byte(text, nbytes, bufsz, v, cor, 0) == '>') f |= LF_NAME;  // fasta.go:206 (after it)
    if (v.e + 1 < nbytes && __ldg(text + v.e + 1) == '>') f |= LF_PEEK_GT;            // fasta.go:192-193
    flags[i] = f;
    if (len >= bufsz) atomicMin(first_long, (unsigned long long)i);
}

// with the state before each line known: header / appended length per line
__global__ void __launch_bounds__(256)
line_mark_kernel(const uint64_t *__restrict__ nl, uint64_t n_lines, const uint8_t *__restrict__ flags,
                 const uint32_t *__restrict__ state_ex, uint32_t *__restrict__ is_hdr, uint32_t *__restrict__ app_len) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_lines) return;
    const uint32_t st = state_ex[i] & 1u;  // prefix function applied to LOOKING
    const uint8_t f = flags[i];
    const LineView v = line_view(nl, i);
    is_hdr[i] = st == ST_LOOKING && (f & LF_NAME);
    app_len[i] = (st == ST_INSIDE && !(f & LF_SKIP)) ? (uint32_t)(v.e - v.b) : 0u;
}

__global__ void __launch_bounds__(256)
record_kernel(const uint64_t *__restrict__ nl, uint64_t n_lines, const uint8_t *__restrict__ flags,
              const uint32_t *__restrict__ state_ex, const uint32_t *__restrict__ is_hdr,
              const unsigned long long *__restrict__ hdr_ex, const unsigned long long *__restrict__ app_ex,
              unsigned long long *__restrict__ seq_off, uint32_t *__restrict__ name_len, uint64_t *__restrict__ end_line) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n_lines) return;
    if (i == n_lines) {
        seq_off[hdr_ex[n_lines]] = app_ex[n_lines];
        return;
    }
    if (is_hdr[i]) {
        const LineView v = line_view(nl, i);
        const unsigned long long m = hdr_ex[i];
        seq_off[m] = app_ex[i];
        name_len[m] = (uint32_t)(v.e - v.b - 1);
    }
    if ((state_ex[i] & 1u) == ST_INSIDE && (flags[i] & LF_PEEK_GT)) end_line[hdr_ex[i + 1] - 1] = i;  // record ends here
}

__global__ void __launch_bounds__(256)
empty_record_kernel(const unsigned long long *__restrict__ seq_off, uint64_t n_check, unsigned long long *__restrict__ first_empty) {
    const uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m < n_check && seq_off[m + 1] == seq_off[m]) atomicMin(first_empty, (unsigned long long)m);
}

// Copy: appended content -> bases, header content -> names (records < n_ok only).  A warp takes 32
// consecutive lines: lane l gathers the metadata of line i0 + l (coalesced loads), then the warp
// copies the 32 lines one after the other, the per-line source / destination / length broadcast by
// shuffles -- lines are ~60 bytes, so the metadata (7 words per line) costs as much as the text.
__global__ void __launch_bounds__(256)
copy_lines_kernel(const uint8_t *__restrict__ text, uint64_t nbytes, uint64_t bufsz, const uint64_t *__restrict__ nl,
                  uint64_t n_lines, const uint8_t *__restrict__ flags, const uint32_t *__restrict__ is_hdr,
                  const uint32_t *__restrict__ app_len, const unsigned long long *__restrict__ hdr_ex,
                  const unsigned long long *__restrict__ app_ex, const unsigned long long *__restrict__ name_off,
                  uint64_t n_ok, uint8_t *__restrict__ bases, uint8_t *__restrict__ names) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint64_t n_groups = (n_lines + 31) / 32;
    for (uint64_t g = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < n_groups; g += warps) {
        const uint64_t i = g * 32 + lane;
        // my line: source offset, destination pointer, length, and whether it needs the slow path
        uint64_t src = 0, len = 0;
        uint8_t *dst = nullptr;
        bool cor = false, hdr = false;
        if (i < n_lines) {
            const unsigned long long hc = hdr_ex[i + 1];  // headers up to and including this line
            if (hc != 0 && hc - 1 < n_ok) {
                const LineView v = line_view(nl, i);
                cor = flags[i] & LF_CORRUPT;
                hdr = is_hdr[i];
                if (hdr) {
                    if (names) { dst = names + name_off[hc - 1]; src = v.b + 1; len = v.e - v.b - 1; }
                } else if (app_len[i]) {
                    dst = bases + app_ex[i]; src = v.b; len = app_len[i];
                }
            }
        }
        uint32_t todo = __ballot_sync(0xffffffffu, len != 0);
        while (todo) {
            const int l = __ffs(todo) - 1;
            todo &= todo - 1;
            const uint64_t s_src = __shfl_sync(0xffffffffu, src, l), s_len = __shfl_sync(0xffffffffu, len, l);
            uint8_t *s_dst = reinterpret_cast<uint8_t *>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(dst), l));
            const bool s_cor = __shfl_sync(0xffffffffu, (int)cor, l);
            if (!s_cor) {
                for (uint64_t j = lane; j < s_len; j += 32) s_dst[j] = __ldg(text + s_src + j);
            } else {  // bytes as the reference sees them after the reader refill (rare)
                const bool s_hdr = __shfl_sync(0xffffffffu, (int)hdr, l);
                const LineView v = line_view(nl, g * 32 + l);
                for (uint64_t j = lane; j < s_len; j += 32) s_dst[j] = effective_byte(text, nbytes, bufsz, v, true, j + (s_hdr ? 1 : 0));
            }
        }
    }
}

template <typename T>
int fetch(T *host, const T *dev, cudaStream_t st) {
    PG_CUDA(cudaMemcpyAsync(host, dev, sizeof(T), cudaMemcpyDeviceToHost, st));
    PG_CUDA(cudaStreamSynchronize(st));
    return PG_OK;
}
#define FA_TRY(expr) do { const int _rc = (expr); if (_rc != PG_OK) return _rc; } while (0)

}  // namespace

int launch_fasta_ingest(const uint8_t *d_text, uint64_t nbytes, uint32_t max_line_size, uint32_t flags,
                        uint8_t *d_bases, uint64_t bases_cap, uint64_t *d_offsets, uint8_t *d_names,
                        uint64_t names_cap, uint64_t *d_name_offsets, uint64_t records_cap, uint64_t *n_records,
                        uint64_t *total_bases, uint64_t *total_name_bytes, int32_t *err_code, uint64_t *err_line,
                        cudaStream_t st) {
    *n_records = 0; *total_bases = 0; *total_name_bytes = 0; *err_code = FA_OK; *err_line = 0;
    const bool want_names = d_names != nullptr || d_name_offsets != nullptr;
    if (want_names && (!d_names || !d_name_offsets)) {
        set_error("names and name_offsets must be given together");
        return PG_ERR_ARG;
    }
    const uint64_t bufsz = max_line_size < 16 ? 16 : max_line_size;  // bufio.NewReaderSize minimum
    if (nbytes == 0) {
        PG_CUDA(cudaMemsetAsync(d_offsets, 0, 8, st));
        if (want_names) PG_CUDA(cudaMemsetAsync(d_name_offsets, 0, 8, st));
        return PG_OK;
    }
    StreamScratch tmp(st);
    uint64_t *d_nl = nullptr;
    uint64_t n_lines = 0, last_plus1 = 0;
    {
        const int rc0 = text::newline_positions(d_text, nbytes, &d_nl, &n_lines, &last_plus1, st);
        tmp.adopt(d_nl);
        if (rc0 != PG_OK) return rc0;
    }
    const uint64_t frag_len = nbytes - last_plus1;  // bytes after the last newline
    uint8_t frag_first = 0;
    if (frag_len) FA_TRY(fetch(&frag_first, d_text + last_plus1, st));

    uint8_t *d_flags = nullptr, *d_corrupt = nullptr;
    uint32_t *d_state = nullptr, *d_hdr = nullptr, *d_app = nullptr, *d_name_len = nullptr;
    unsigned long long *d_hdr_ex = nullptr, *d_app_ex = nullptr, *d_seq_off = nullptr, *d_name_off = nullptr, *d_scalars = nullptr;
    uint64_t *d_end_line = nullptr;
    PG_CUDA(tmp.alloc(&d_scalars, 2));  // [0] first over-long line, [1] first empty record
    PG_CUDA(cudaMemsetAsync(d_scalars, 0xff, 16, st));
    PG_CUDA(tmp.alloc(&d_flags, n_lines));
    PG_CUDA(tmp.alloc(&d_state, n_lines + 1));
    PG_CUDA(tmp.alloc(&d_hdr, n_lines));
    PG_CUDA(tmp.alloc(&d_app, n_lines));
    PG_CUDA(tmp.alloc(&d_hdr_ex, n_lines + 1));
    PG_CUDA(tmp.alloc(&d_app_ex, n_lines + 1));
    const unsigned line_blocks = (unsigned)((n_lines + 255) / 256);
    if ((flags & PG_FASTA_BUFIO_ALIAS) && n_lines && nbytes > bufsz) {
        if (n_lines >= 0x7fffffffull) {
            set_error("PG_FASTA_BUFIO_ALIAS supports fewer than 2^31 lines");
            return PG_ERR_UNSUPPORTED;
        }
        uint32_t *d_next = nullptr;
        PG_CUDA(tmp.alloc(&d_corrupt, n_lines));
        PG_CUDA(tmp.alloc(&d_next, n_lines));
        PG_CUDA(cudaMemsetAsync(d_corrupt, 0, n_lines, st));
        refill_next_kernel<<<line_blocks, 256, 0, st>>>(d_nl, n_lines, nbytes, bufsz, d_next);
        PG_LAUNCH_CHECK("refill_next_kernel");
        refill_chain_kernel<<<1, 1, 0, st>>>(d_next, n_lines, d_corrupt);
        PG_LAUNCH_CHECK("refill_chain_kernel");
    }
    if (n_lines) {
        line_flags_kernel<<<line_blocks, 256, 0, st>>>(d_text, nbytes, d_nl, n_lines, bufsz, d_corrupt, d_flags, d_scalars);
        PG_LAUNCH_CHECK("line_flags_kernel");
    }
    FA_TRY(text::device_scan<StateOp>(d_flags, n_lines, d_state, st));
    if (n_lines) {
        line_mark_kernel<<<line_blocks, 256, 0, st>>>(d_nl, n_lines, d_flags, d_state, d_hdr, d_app);
        PG_LAUNCH_CHECK("line_mark_kernel");
    }
    FA_TRY(text::device_scan<text::SumOp<uint32_t>>(d_hdr, n_lines, d_hdr_ex, st));
    FA_TRY(text::device_scan<text::SumOp<uint32_t>>(d_app, n_lines, d_app_ex, st));
    unsigned long long n_hdr = 0, total_app = 0, first_long = ~0ull;
    uint32_t end_fn = 0;
    FA_TRY(fetch(&n_hdr, d_hdr_ex + n_lines, st));
    FA_TRY(fetch(&total_app, d_app_ex + n_lines, st));
    FA_TRY(fetch(&end_fn, d_state + n_lines, st));
    FA_TRY(fetch(&first_long, d_scalars, st));
    if (frag_len >= bufsz && first_long == ~0ull) first_long = n_lines;  // the unterminated tail is a line too

    PG_CUDA(tmp.alloc(&d_seq_off, n_hdr + 1));
    PG_CUDA(tmp.alloc(&d_name_len, n_hdr));
    PG_CUDA(tmp.alloc(&d_name_off, n_hdr + 1));
    PG_CUDA(tmp.alloc(&d_end_line, n_hdr));
    record_kernel<<<(unsigned)((n_lines + 1 + 255) / 256), 256, 0, st>>>(d_nl, n_lines, d_flags, d_state, d_hdr, d_hdr_ex, d_app_ex,
                                                                        d_seq_off, d_name_len, d_end_line);
    PG_LAUNCH_CHECK("record_kernel");

    // ---- where does the reference stop?  (fasta.go:225-243 + ParseN :104-116) ----
    uint64_t n_check;      // records that ran to their end: each must be non-empty
    int32_t tail_code = FA_OK;
    uint64_t tail_line = 0, n_ok;
    if (first_long != ~0ull) {  // parse stops on this line
        uint32_t st_before = end_fn & 1u;
        unsigned long long hdr_before = n_hdr, app_before = total_app;
        uint8_t first_byte = frag_first;
        if (first_long < n_lines) {
            uint32_t fn = 0;
            uint64_t prev_nl = 0;
            FA_TRY(fetch(&fn, d_state + first_long, st));
            FA_TRY(fetch(&hdr_before, d_hdr_ex + first_long, st));
            FA_TRY(fetch(&app_before, d_app_ex + first_long, st));
            if (first_long) FA_TRY(fetch(&prev_nl, d_nl + first_long - 1, st));
            FA_TRY(fetch(&first_byte, d_text + (first_long ? prev_nl + 1 : 0), st));
            st_before = fn & 1u;
        }
        n_check = hdr_before - (st_before == ST_INSIDE ? 1 : 0);
        n_ok = n_check;
        if (first_byte != ';') {  // fasta.go:178-180
            tail_code = FA_ERR_LINE_TOO_LONG;
            tail_line = first_long + 2;
        } else {  // skippable: `break` with ErrBufferFull (fasta.go:172-176)
            tail_line = first_long + 1;
            if (st_before == ST_LOOKING) {
                tail_code = FA_ERR_NO_START;
            } else {
                unsigned long long rec_begin = 0;
                FA_TRY(fetch(&rec_begin, d_seq_off + n_check, st));
                tail_code = app_before == rec_begin ? FA_ERR_EMPTY_SEQ : FA_ERR_BUFFER_FULL;
            }
        }
    } else {  // the parse reaches the end of the text
        const bool frag_skippable = frag_len <= 1 || frag_first == ';';
        const uint32_t st_end = end_fn & 1u;
        if (st_end == ST_LOOKING) {
            n_check = n_ok = n_hdr;
            if (frag_skippable) {  // err == nil -> "did not find fasta start" is reported
                tail_code = FA_ERR_NO_START;
                tail_line = n_lines + 1;
            }  // else the error wraps io.EOF and ParseN swallows it
        } else {
            n_check = n_hdr - 1;  // the last record is still open
            unsigned long long rec_begin = 0;
            FA_TRY(fetch(&rec_begin, d_seq_off + n_check, st));
            if (!frag_skippable) {
                n_ok = n_check;  // fasta returned together with io.EOF: dropped by ParseN
            } else if (total_app == rec_begin) {
                n_ok = n_check;
                tail_code = FA_ERR_EMPTY_SEQ;
                tail_line = n_lines + 1;
            } else {
                n_ok = n_hdr;
            }
        }
    }
    if (n_check) {
        empty_record_kernel<<<(unsigned)((n_check + 255) / 256), 256, 0, st>>>(d_seq_off, n_check, d_scalars + 1);
        PG_LAUNCH_CHECK("empty_record_kernel");
        unsigned long long first_empty = ~0ull;
        FA_TRY(fetch(&first_empty, d_scalars + 1, st));
        if (first_empty != ~0ull) {
            uint64_t el = 0;
            FA_TRY(fetch(&el, d_end_line + first_empty, st));
            n_ok = first_empty;
            tail_code = FA_ERR_EMPTY_SEQ;
            tail_line = el + 1;
        }
    }
    *err_code = tail_code;
    *err_line = tail_line;
    *n_records = n_ok;

    unsigned long long seq_total = 0, name_total = 0;
    FA_TRY(fetch(&seq_total, d_seq_off + n_ok, st));
    *total_bases = seq_total;
    if (want_names) {
        FA_TRY(text::device_scan<text::SumOp<uint32_t>>(d_name_len, n_ok, d_name_off, st));
        FA_TRY(fetch(&name_total, d_name_off + n_ok, st));
        *total_name_bytes = name_total;
    }
    if (n_ok > records_cap || seq_total > bases_cap || (want_names && name_total > names_cap)) {
        set_error("capacity: need %llu records, %llu sequence bytes, %llu name bytes", (unsigned long long)n_ok, seq_total, name_total);
        return PG_ERR_ARG;
    }
    PG_CUDA(cudaMemcpyAsync(d_offsets, d_seq_off, (n_ok + 1) * 8, cudaMemcpyDeviceToDevice, st));
    if (want_names) PG_CUDA(cudaMemcpyAsync(d_name_offsets, d_name_off, (n_ok + 1) * 8, cudaMemcpyDeviceToDevice, st));
    if (n_ok && n_lines) {
        const unsigned blocks = (unsigned)std::min<uint64_t>((n_lines + 255) / 256, (uint64_t)sm_count() * 16);
        copy_lines_kernel<<<blocks, 256, 0, st>>>(d_text, nbytes, bufsz, d_nl, n_lines, d_flags, d_hdr, d_app, d_hdr_ex, d_app_ex,
                                                 d_name_off, n_ok, d_bases, want_names ? d_names : nullptr);
        PG_LAUNCH_CHECK("copy_lines_kernel");
    }
    PG_CUDA(cudaStreamSynchronize(st));
    return PG_OK;
}

}  // namespace pg
