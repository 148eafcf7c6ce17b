// K5 -- contig x motif methylation-pattern table (the operator nanomotif's binnary commands obtain
// from the external Rust package: epymetheus.methylation_pattern, call site nanomotif/main.py:167-178).
// Spec: DESIGN.md section 4 (K5) -- per (contig, motif): occurrences on both strands joined to the
// pileup rows that pass the read-coverage filters; n_motif_obs, sum n_mod, sum n_valid_cov, and the
// median of the per-occurrence fractions n_mod / n_valid_cov.
#include "common.cuh"

namespace nmb {

__device__ __forceinline__ bool row_hits(const uint32_t *__restrict__ plane_fwd,
                                         const uint32_t *__restrict__ plane_rev, int64_t g, uint8_t strand) {
    const uint32_t *pl = strand ? plane_rev : plane_fwd;
    return (__ldg(pl + (g >> 5)) >> (g & 31)) & 1u;
}

// Pass 1: per contig number of hits, sum of n_mod and of n_valid_cov.  Rows are sorted by contig, so
// equal contigs are aggregated inside the warp before touching the counters.
__global__ void __launch_bounds__(256) pattern_count_kernel(
    const int64_t *__restrict__ gpos, const uint8_t *__restrict__ strand, const int32_t *__restrict__ contig,
    const int32_t *__restrict__ n_mod, const int32_t *__restrict__ n_cov, int64_t n_rows,
    const uint32_t *__restrict__ plane_fwd, const uint32_t *__restrict__ plane_rev,
    unsigned long long *__restrict__ stats /* [n_contigs][3] */) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    for (int64_t r0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) - lane; r0 < n_rows; r0 += stride) {
        const int64_t r = r0 + lane;
        bool hit = false;
        int c = -1;
        unsigned long long m = 0, v = 0;
        if (r < n_rows) {
            hit = row_hits(plane_fwd, plane_rev, gpos[r], strand[r]);
            if (hit) { c = contig[r]; m = (unsigned)n_mod[r]; v = (unsigned)n_cov[r]; }
        }
        const unsigned hits = __ballot_sync(0xFFFFFFFFu, hit);
        if (!hits) continue;
        if (hit) {
            const unsigned peers = __match_any_sync(hits, c);
            const int leader = __ffs(peers) - 1;
            unsigned long long sm = 0, sv = 0;
            for (unsigned rest = peers; rest; rest &= rest - 1) {  // segmented sum over the peer lanes
                const int src = __ffs(rest) - 1;
                sm += __shfl_sync(peers, m, src);
                sv += __shfl_sync(peers, v, src);
            }
            if (lane == leader) {
                atomicAdd(&stats[3 * (size_t)c + 0], (unsigned long long)__popc(peers));
                atomicAdd(&stats[3 * (size_t)c + 1], sm);
                atomicAdd(&stats[3 * (size_t)c + 2], sv);
            }
        }
    }
}

// offsets[c] = exclusive prefix sum of stats[c][0]; cursor[c] = 0.
__global__ void __launch_bounds__(1024) pattern_offsets_kernel(const unsigned long long *__restrict__ stats,
                                                               int n_contigs, long long *__restrict__ offsets,
                                                               int *__restrict__ cursor) {
    __shared__ long long s_part[1024];
    const int t = threadIdx.x;
    const int per = (n_contigs + 1023) / 1024;
    const int b = t * per, e = min(n_contigs, b + per);
    long long sum = 0;
    for (int i = b; i < e; ++i) sum += (long long)stats[3 * (size_t)i];
    s_part[t] = sum;
    __syncthreads();
    if (t == 0) {
        long long run = 0;
        for (int i = 0; i < 1024; ++i) { const long long x = s_part[i]; s_part[i] = run; run += x; }
    }
    __syncthreads();
    long long run = s_part[t];
    for (int i = b; i < e; ++i) {
        offsets[i] = run;
        cursor[i] = 0;
        run += (long long)stats[3 * (size_t)i];
    }
    if (t == 1023 || e == n_contigs) offsets[n_contigs] = run;  // total (written by whoever ends the range)
}

// Pass 2: write the per-occurrence fraction of every hit into its contig's segment.
__global__ void __launch_bounds__(256) pattern_write_kernel(
    const int64_t *__restrict__ gpos, const uint8_t *__restrict__ strand, const int32_t *__restrict__ contig,
    const int32_t *__restrict__ n_mod, const int32_t *__restrict__ n_cov, int64_t n_rows,
    const uint32_t *__restrict__ plane_fwd, const uint32_t *__restrict__ plane_rev,
    const long long *__restrict__ offsets, int *__restrict__ cursor, double *__restrict__ fractions) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += stride) {
        if (!row_hits(plane_fwd, plane_rev, gpos[r], strand[r])) continue;
        const int c = contig[r];
        const int slot = atomicAdd(&cursor[c], 1);
        fractions[offsets[c] + slot] = (double)n_mod[r] / (double)n_cov[r];
    }
}

// k-th smallest (0-based) of v[0..n) for non-negative doubles by MSB-first radix select on the bit
// patterns (IEEE order == integer order for x >= 0).  One block; s_hist has 256 entries.
__device__ double block_select(const double *__restrict__ v, long long n, long long k, unsigned *s_hist,
                               unsigned long long *s_state) {
    unsigned long long prefix = 0, mask = 0;
    for (int shift = 56; shift >= 0; shift -= 8) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) s_hist[i] = 0;
        __syncthreads();
        for (long long i = threadIdx.x; i < n; i += blockDim.x) {
            const unsigned long long b = (unsigned long long)__double_as_longlong(v[i]);
            if ((b & mask) == prefix) atomicAdd(&s_hist[(b >> shift) & 0xFF], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            long long kk = (long long)s_state[0];
            int d = 0;
            for (; d < 256; ++d) {
                if (kk < (long long)s_hist[d]) break;
                kk -= s_hist[d];
            }
            s_state[0] = (unsigned long long)kk;
            s_state[1] = (unsigned long long)d;
        }
        __syncthreads();
        prefix |= s_state[1] << shift;
        mask |= 0xFFull << shift;
        __syncthreads();
    }
    (void)k;
    return __longlong_as_double((long long)prefix);
}

// One block per contig: median of its fraction segment (mean of the two middle values for even n).
__global__ void __launch_bounds__(128) pattern_median_kernel(const double *__restrict__ fractions,
                                                             const long long *__restrict__ offsets,
                                                             int n_contigs, double *__restrict__ median) {
    __shared__ unsigned s_hist[256];
    __shared__ unsigned long long s_state[2];
    const int c = blockIdx.x;
    const long long b = offsets[c], n = offsets[c + 1] - b;
    if (n <= 0) {
        if (threadIdx.x == 0) median[c] = __longlong_as_double(0x7FF8000000000000ll);  // NaN: no observation
        return;
    }
    const double *v = fractions + b;
    if (n <= 32) {  // one warp, one value per lane: rank by counting
        if (threadIdx.x >= 32) return;
        const int lane = threadIdx.x;
        const double x = lane < n ? v[lane] : 0.0;
        int rank = 0;
        for (int j = 0; j < (int)n; ++j) {
            const double y = __shfl_sync(0xFFFFFFFFu, x, j);
            rank += (y < x) || (y == x && j < lane);
        }
        const int k_hi = (int)(n / 2), k_lo = (int)((n - 1) / 2);
        const unsigned hi = __ballot_sync(0xFFFFFFFFu, lane < n && rank == k_hi);
        const unsigned lo = __ballot_sync(0xFFFFFFFFu, lane < n && rank == k_lo);
        const double a = __shfl_sync(0xFFFFFFFFu, x, __ffs(lo) - 1);
        const double d = __shfl_sync(0xFFFFFFFFu, x, __ffs(hi) - 1);
        if (lane == 0) median[c] = (a + d) / 2;  // a == d for odd n: (a + a) / 2 == a exactly
        return;
    }
    if (threadIdx.x == 0) s_state[0] = (unsigned long long)((n - 1) / 2);
    __syncthreads();
    const double a = block_select(v, n, (n - 1) / 2, s_hist, s_state);
    double d = a;
    if ((n & 1) == 0) {
        __syncthreads();
        if (threadIdx.x == 0) s_state[0] = (unsigned long long)(n / 2);
        __syncthreads();
        d = block_select(v, n, n / 2, s_hist, s_state);
    }
    if (threadIdx.x == 0) median[c] = (a + d) / 2;
}

static inline unsigned pattern_blocks(int64_t n) {
    int64_t b = (n + 255) / 256;
    if (b > 148 * 16) b = 148 * 16;
    return (unsigned)(b < 1 ? 1 : b);
}

}  // namespace nmb

extern "C" {

int nmb_pattern_stats(const int64_t *gpos, const uint8_t *strand, const int32_t *contig_id,
                      const int32_t *n_mod, const int32_t *n_valid_cov, int64_t n_rows,
                      const uint32_t *plane_fwd, const uint32_t *plane_rev, int32_t n_contigs,
                      int64_t *stats, int64_t *offsets, int32_t *cursor, void *stream) {
    NMB_REQUIRE(n_rows >= 0 && n_contigs > 0, "nmb_pattern_stats: n_rows=%lld n_contigs=%d", (long long)n_rows,
                n_contigs);
    NMB_REQUIRE(plane_fwd && plane_rev && stats && offsets && cursor, "nmb_pattern_stats: null argument");
    cudaStream_t s = (cudaStream_t)stream;
    NMB_CUDA(cudaMemsetAsync(stats, 0, (size_t)n_contigs * 3 * sizeof(int64_t), s));
    if (n_rows > 0) {
        NMB_REQUIRE(gpos && strand && contig_id && n_mod && n_valid_cov, "nmb_pattern_stats: null column");
        nmb::pattern_count_kernel<<<nmb::pattern_blocks(n_rows), 256, 0, s>>>(
            gpos, strand, contig_id, n_mod, n_valid_cov, n_rows, plane_fwd, plane_rev,
            (unsigned long long *)stats);
        NMB_CUDA(cudaGetLastError());
    }
    nmb::pattern_offsets_kernel<<<1, 1024, 0, s>>>((const unsigned long long *)stats, n_contigs,
                                                  (long long *)offsets, cursor);
    NMB_CUDA(cudaGetLastError());
    return NMB_OK;
}

int nmb_pattern_median(const int64_t *gpos, const uint8_t *strand, const int32_t *contig_id,
                       const int32_t *n_mod, const int32_t *n_valid_cov, int64_t n_rows,
                       const uint32_t *plane_fwd, const uint32_t *plane_rev, int32_t n_contigs,
                       const int64_t *offsets, int32_t *cursor, double *fractions, double *median,
                       void *stream) {
    NMB_REQUIRE(n_rows >= 0 && n_contigs > 0, "nmb_pattern_median: bad sizes");
    NMB_REQUIRE(plane_fwd && plane_rev && offsets && cursor && median, "nmb_pattern_median: null argument");
    cudaStream_t s = (cudaStream_t)stream;
    if (n_rows > 0) {
        NMB_REQUIRE(gpos && strand && contig_id && n_mod && n_valid_cov && fractions,
                    "nmb_pattern_median: null column");
        nmb::pattern_write_kernel<<<nmb::pattern_blocks(n_rows), 256, 0, s>>>(
            gpos, strand, contig_id, n_mod, n_valid_cov, n_rows, plane_fwd, plane_rev,
            (const long long *)offsets, cursor, fractions);
        NMB_CUDA(cudaGetLastError());
    }
    nmb::pattern_median_kernel<<<n_contigs, 128, 0, s>>>(fractions, (const long long *)offsets, n_contigs,
                                                        median);
    NMB_CUDA(cudaGetLastError());
    return NMB_OK;
}

int nmb_segment_median(const double *fractions, const int64_t *offsets, int64_t n_segments, double *median,
                       void *stream) {
    NMB_REQUIRE(offsets && median && n_segments >= 0 && n_segments < (1ll << 31), "nmb_segment_median: bad argument");
    if (n_segments == 0) return NMB_OK;
    nmb::pattern_median_kernel<<<(unsigned)n_segments, 128, 0, (cudaStream_t)stream>>>(
        fractions, (const long long *)offsets, (int)n_segments, median);
    NMB_CUDA(cudaGetLastError());
    return NMB_OK;
}

}  // extern "C"
