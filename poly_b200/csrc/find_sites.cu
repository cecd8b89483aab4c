// find_sites.cu -- every occurrence of every pattern in every sequence (exact match, overlapping
// occurrences included): the role index/suffixarray's Lookup(pattern, -1) plays in
// pcr.SimulateSimple, /root/reference/primers/pcr/pcr.go:87,110-115 (primer binding sites of the
// minimal primers and of their reverse complements).  SURVEY.md 8f.3.
//
// One thread per text position; the patterns (a primer list: a few kB) are staged in shared memory
// and tried in turn, the first byte rejecting 3 of 4 candidates.  Hits are appended through one
// atomic counter; their order is unspecified (Lookup's is too -- the caller sorts, pcr.go:127-128).
// PG_SITES_UPPER compares the upper-cased sequence byte (pcr.go:82 strings.ToUpper).
#include "common.cuh"

namespace pg {

namespace {

constexpr int FS_THREADS = 256;
constexpr uint32_t FS_SMEM_PATTERN_BYTES = 32 * 1024;

__device__ __forceinline__ uint8_t fs_upper(uint8_t c) { return (c >= 'a' && c <= 'z') ? (uint8_t)(c - 32) : c; }

template <bool SMEM>
__global__ void __launch_bounds__(FS_THREADS)
find_sites_kernel(const uint8_t *__restrict__ seqs, const uint64_t *__restrict__ seq_off, uint64_t n_seq,
                  const uint8_t *__restrict__ pats, const uint64_t *__restrict__ pat_off, uint32_t n_pat, uint32_t flags,
                  uint32_t *__restrict__ hit_seq, uint64_t *__restrict__ hit_pos, uint32_t *__restrict__ hit_pat, uint64_t cap,
                  unsigned long long *__restrict__ n_hits) {
    extern __shared__ uint8_t s_pat[];
    const uint64_t pat_bytes = pat_off[n_pat];
    if (SMEM) {
        for (uint64_t j = threadIdx.x; j < pat_bytes; j += FS_THREADS) s_pat[j] = pats[j];
        __syncthreads();
    }
    const uint8_t *pp = SMEM ? s_pat : pats;
    const uint64_t total = seq_off[n_seq];
    const bool up = flags & PG_SITES_UPPER;
    for (uint64_t g = seq_off[0] + (uint64_t)blockIdx.x * FS_THREADS + threadIdx.x; g < total; g += (uint64_t)gridDim.x * FS_THREADS) {
        uint64_t lo = 0, hi = n_seq;  // sequence holding byte g: last q with seq_off[q] <= g
        while (hi - lo > 1) {
            const uint64_t mid = (lo + hi) >> 1;
            if (seq_off[mid] <= g) lo = mid; else hi = mid;
        }
        const uint64_t room = seq_off[lo + 1] - g;  // bytes left in this sequence from here
        uint8_t c0 = __ldg(seqs + g);
        if (up) c0 = fs_upper(c0);
        for (uint32_t q = 0; q < n_pat; ++q) {
            const uint64_t pb = pat_off[q], m = pat_off[q + 1] - pb;
            if (m == 0 || m > room || pp[pb] != c0) continue;  // Lookup("") is nil
            uint64_t j = 1;
            for (; j < m; ++j) {
                uint8_t c = __ldg(seqs + g + j);
                if (up) c = fs_upper(c);
                if (c != pp[pb + j]) break;
            }
            if (j == m) {
                const unsigned long long slot = atomicAdd(n_hits, 1ull);
                if (slot < cap) {
                    hit_seq[slot] = (uint32_t)lo;
                    hit_pos[slot] = g - seq_off[lo];
                    hit_pat[slot] = q;
                }
            }
        }
    }
}

}  // namespace

// *n_hits (device) must be zero on entry; it ends as the number of occurrences (may exceed cap:
// only the first `cap` appended are stored).
int launch_find_sites(const uint8_t *d_seqs, const uint64_t *d_seq_off, uint64_t n_seq, uint64_t total_bytes,
                      const uint8_t *d_pats, const uint64_t *d_pat_off, uint32_t n_pat, uint64_t pat_bytes, uint32_t flags,
                      uint32_t *d_hit_seq, uint64_t *d_hit_pos, uint32_t *d_hit_pat, uint64_t cap,
                      unsigned long long *d_n_hits, cudaStream_t st) {
    if (n_seq == 0 || n_pat == 0 || total_bytes == 0) return PG_OK;
    const unsigned blocks = (unsigned)std::min<uint64_t>((total_bytes + FS_THREADS - 1) / FS_THREADS, (uint64_t)sm_count() * 32);
    if (pat_bytes <= FS_SMEM_PATTERN_BYTES)
        find_sites_kernel<true><<<blocks, FS_THREADS, pat_bytes, st>>>(d_seqs, d_seq_off, n_seq, d_pats, d_pat_off, n_pat, flags, d_hit_seq,
                                                                      d_hit_pos, d_hit_pat, cap, d_n_hits);
    else
        find_sites_kernel<false><<<blocks, FS_THREADS, 0, st>>>(d_seqs, d_seq_off, n_seq, d_pats, d_pat_off, n_pat, flags, d_hit_seq,
                                                               d_hit_pos, d_hit_pat, cap, d_n_hits);
    PG_LAUNCH_CHECK("find_sites_kernel");
    return PG_OK;
}

}  // namespace pg
