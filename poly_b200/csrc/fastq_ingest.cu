// fastq_ingest.cu -- the step BEFORE the hot path: FASTQ text -> the dense
// bases + offsets layout the sketch entry points take, on the GPU.
//
// Mirrors Parser.ParseNext / ParseN of /root/reference/io/fastq/fastq.go:117-214,88-99 for a whole
// buffer: strict 4-line records (identifier, sequence, "+", quality), every line newline
// terminated; parsing stops at the first record the reference would reject and the records before
// it are returned together with the error (as ParseN does).  Only what the sketch path needs is
// materialised (the sequences); identifiers / optionals / qualities are validated, not copied.
//
//   1. count '\n' per 4 KB block, exclusive scan of the block counts, write newline positions
//   2. one thread per record: locate its 4 lines, apply the reference's checks in its order
//   3. exclusive scan of the sequence lengths of the valid prefix -> offsets
//   4. one warp per record copies its sequence bytes into the dense buffer
#include "text_scan.cuh"

namespace pg {

namespace {

constexpr uint64_t FQ_MAX_LINE = 2 * 32 * 1024;  // fastq.go:56 maxLineSize of Parse / Read

// error codes (per record and overall)
enum : int32_t {
    FQ_OK = 0,
    FQ_ERR_EOF = 1,        // a line of the record is not newline terminated ("unexepcted EOF encountered")
    FQ_ERR_EMPTY_SEQ = 2,  // fastq.go:179-181
    FQ_ERR_EMPTY_QUAL = 3, // fastq.go:197-199
    FQ_ERR_NO_AT = 4,      // fastq.go:203-205
    FQ_ERR_PANIC = 5,      // empty identifier line (string(line)[0]) or an optional without '=' (fastq.go:160-170)
    FQ_ERR_LINE_TOO_LONG = 6  // bufio.ErrBufferFull with the 64 KiB reader of Parse / Read
};

// One thread per candidate record r: lines 4r .. 4r+3 (a line ends at nl[i]); the last candidate
// may be incomplete (text after the last newline, or fewer than 4 lines left).
__global__ void __launch_bounds__(256)
record_kernel(const uint8_t *__restrict__ text, uint64_t n, const uint64_t *__restrict__ nl, uint64_t n_lines,
              uint64_t n_cand, uint32_t *__restrict__ seq_len, uint64_t *__restrict__ seq_start,
              int32_t *__restrict__ rec_err, uint32_t *__restrict__ rec_err_line, unsigned long long *__restrict__ first_bad,
              uint64_t *__restrict__ spans) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_cand) return;
    auto line_begin = [&](uint64_t li) -> uint64_t { return li == 0 ? 0 : nl[li - 1] + 1; };
    int32_t err = FQ_OK;
    uint32_t err_line = 0;
    uint64_t s_beg = 0, s_len = 0, id_beg = 0, id_len = 0, q_beg = 0, q_len = 0;
    bool no_at = false;
    for (int l = 0; l < 4 && err == FQ_OK; ++l) {
        const uint64_t li = 4 * r + l;
        if (li >= n_lines) {  // no newline left: ReadSlice fills its buffer first -- 64 KiB without a delimiter is
            // bufio.ErrBufferFull, less is io.EOF (with or without trailing bytes; the empty tail included:
            // Peek succeeded earlier only if bytes remain)
            const uint64_t tail = n - line_begin(li);
            err = tail >= FQ_MAX_LINE ? FQ_ERR_LINE_TOO_LONG : FQ_ERR_EOF;
            err_line = l + 1;
            break;
        }
        const uint64_t b = line_begin(li), e = nl[li];  // [b, e) without the newline
        if (e + 1 - b > FQ_MAX_LINE) { err = FQ_ERR_LINE_TOO_LONG; err_line = l + 1; break; }
        if (l == 0) {
            if (e == b) { err = FQ_ERR_PANIC; err_line = 1; break; }  // string(line)[0] on ""
            no_at = __ldg(text + b) != '@';
            id_beg = b;
            id_len = e - b;
            // strings.Split(line, " ")[1:]: every datum (also an empty one) must hold '=' (optionalSplits[1])
            uint32_t token = 0;
            bool has_eq = false;
            for (uint64_t p = b; p <= e; ++p) {
                const uint8_t c = p < e ? __ldg(text + p) : (uint8_t)' ';  // sentinel ends the last token
                if (c == ' ') {
                    if (token >= 1 && !has_eq) { err = FQ_ERR_PANIC; err_line = 1; break; }
                    ++token;
                    has_eq = false;
                } else if (c == '=') {
                    has_eq = true;
                }
            }
        } else if (l == 1) {
            if (e == b) { err = FQ_ERR_EMPTY_SEQ; err_line = 2; break; }
            s_beg = b;
            s_len = e - b;
        } else if (l == 3) {
            if (e == b) { err = FQ_ERR_EMPTY_QUAL; err_line = 4; break; }
            q_beg = b;
            q_len = e - b;
        }
    }
    if (err == FQ_OK && no_at) { err = FQ_ERR_NO_AT; err_line = 4; }
    seq_len[r] = err == FQ_OK ? (uint32_t)s_len : 0u;
    seq_start[r] = s_beg;
    rec_err[r] = err;
    rec_err_line[r] = err_line;
    if (spans) {  // identifier line (with its '@', without the newline) and quality line of the record, as text spans
        spans[4 * r + 0] = id_beg; spans[4 * r + 1] = id_len;
        spans[4 * r + 2] = q_beg;  spans[4 * r + 3] = q_len;
    }
    if (err != FQ_OK) atomicMin(first_bad, (unsigned long long)r);
}

__global__ void __launch_bounds__(256)
copy_sequences_kernel(const uint8_t *__restrict__ text, const uint64_t *__restrict__ seq_start,
                      const uint64_t *__restrict__ offsets, uint64_t n_rec, uint8_t *__restrict__ bases) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t r = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_rec; r += warps) {
        const uint8_t *src = text + seq_start[r];
        uint8_t *dst = bases + offsets[r];
        const uint64_t len = offsets[r + 1] - offsets[r];
        for (uint64_t i = lane; i < len; i += 32) dst[i] = __ldg(src + i);
    }
}

}  // namespace

// Device-resident ingest.  d_bases needs room for `bases_cap` bytes, d_offsets for records_cap + 1
// entries.  Host outputs: *n_records, *total_bases, *err_code / *err_line (1-based line number in
// the file of the line the reference stops at).  PG_ERR_ARG if a capacity is too small
// (*n_records / *total_bases then hold the required sizes).
int launch_fastq_ingest(const uint8_t *d_text, uint64_t nbytes, uint8_t *d_bases, uint64_t bases_cap,
                        uint64_t *d_offsets, uint64_t records_cap, uint64_t *n_records,
                        uint64_t *total_bases, int32_t *err_code, uint64_t *err_line, cudaStream_t st, uint64_t *d_spans) {
    *n_records = 0; *total_bases = 0; *err_code = FQ_OK; *err_line = 0;
    if (nbytes == 0) {
        if (records_cap + 1 >= 1 && d_offsets) PG_CUDA(cudaMemsetAsync(d_offsets, 0, 8, st));
        return PG_OK;
    }
    StreamScratch tmp(st);
    uint64_t *d_nl = nullptr;
    uint64_t n_lines = 0, last_nl_plus1 = 0;
    {
        const int rc0 = text::newline_positions(d_text, nbytes, &d_nl, &n_lines, &last_nl_plus1, st);
        tmp.adopt(d_nl);
        if (rc0 != PG_OK) return rc0;
    }
    const bool trailing = last_nl_plus1 < nbytes;  // bytes after the last newline
    const uint64_t n_cand = n_lines / 4 + ((n_lines % 4 != 0 || trailing) ? 1 : 0);
    int rc = PG_OK;
    uint32_t *d_len = nullptr, *d_eline = nullptr;
    uint64_t *d_sstart = nullptr, *d_off_tmp = nullptr, *d_span_tmp = nullptr;
    int32_t *d_err = nullptr;
    unsigned long long *d_first = nullptr;
    if (n_cand) {
        PG_CUDA(tmp.alloc(&d_len, n_cand));
        PG_CUDA(tmp.alloc(&d_eline, n_cand));
        PG_CUDA(tmp.alloc(&d_sstart, n_cand));
        PG_CUDA(tmp.alloc(&d_err, n_cand));
        PG_CUDA(tmp.alloc(&d_first, 1));
        PG_CUDA(cudaMemsetAsync(d_first, 0xff, 8, st));
        if (d_spans) PG_CUDA(tmp.alloc(&d_span_tmp, 4 * n_cand));
        record_kernel<<<(unsigned)((n_cand + 255) / 256), 256, 0, st>>>(d_text, nbytes, d_nl, n_lines, n_cand, d_len, d_sstart,
                                                                       d_err, d_eline, d_first, d_span_tmp);
        note_launch("record_kernel");
        unsigned long long first_bad = ~0ull;
        PG_CUDA(cudaMemcpyAsync(&first_bad, d_first, 8, cudaMemcpyDeviceToHost, st));
        PG_CUDA(cudaStreamSynchronize(st));
        uint64_t n_ok = n_cand;
        if (first_bad != ~0ull) {
            n_ok = first_bad;
            int32_t e = 0;
            uint32_t el = 0;
            PG_CUDA(cudaMemcpyAsync(&e, d_err + first_bad, 4, cudaMemcpyDeviceToHost, st));
            PG_CUDA(cudaMemcpyAsync(&el, d_eline + first_bad, 4, cudaMemcpyDeviceToHost, st));
            PG_CUDA(cudaStreamSynchronize(st));
            *err_code = e;
            *err_line = 4 * first_bad + el;
        }
        *n_records = n_ok;
        if (n_ok > records_cap) {
            set_error("records_cap %llu < %llu records", (unsigned long long)records_cap, (unsigned long long)n_ok);
            rc = PG_ERR_ARG;
        } else {
            PG_CUDA(tmp.alloc(&d_off_tmp, n_ok + 1));
            {
                const int rc1 = text::device_scan<text::SumOp<uint32_t>>(d_len, n_ok, (unsigned long long *)d_off_tmp, st);
                if (rc1 != PG_OK) return rc1;
            }
            PG_CUDA(cudaMemcpyAsync(total_bases, d_off_tmp + n_ok, 8, cudaMemcpyDeviceToHost, st));
            PG_CUDA(cudaStreamSynchronize(st));
            if (*total_bases > bases_cap) {
                set_error("bases_cap %llu < %llu bases", (unsigned long long)bases_cap, (unsigned long long)*total_bases);
                rc = PG_ERR_ARG;
            } else {
                PG_CUDA(cudaMemcpyAsync(d_offsets, d_off_tmp, (n_ok + 1) * 8, cudaMemcpyDeviceToDevice, st));
                if (d_spans && n_ok) PG_CUDA(cudaMemcpyAsync(d_spans, d_span_tmp, n_ok * 32, cudaMemcpyDeviceToDevice, st));
                if (n_ok) {
                    const unsigned blocks = (unsigned)std::min<uint64_t>((n_ok + 7) / 8, (uint64_t)sm_count() * 16);
                    copy_sequences_kernel<<<blocks, 256, 0, st>>>(d_text, d_sstart, d_off_tmp, n_ok, d_bases);
                    note_launch("copy_sequences_kernel");
                }
            }
        }
    } else {
        PG_CUDA(cudaMemsetAsync(d_offsets, 0, 8, st));
    }
    cudaError_t e = cudaGetLastError();
    PG_CUDA(cudaStreamSynchronize(st));
    if (e != cudaSuccess) return cuda_fail(e, "fastq ingest", __FILE__, __LINE__);
    return rc;
}

}  // namespace pg
