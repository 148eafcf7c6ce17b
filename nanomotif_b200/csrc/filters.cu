// Pileup filters on columnar device arrays (replaces the polars expressions of
// nanomotif/dataload.py:191-247).  Each kernel writes a keep mask in input row order.
#include "common.cuh"

namespace nmb {

// dataload.py:199 -- Nvalid_cov > min_coverage (strict).
__global__ void __launch_bounds__(256) filter_coverage_kernel(const int64_t *__restrict__ cov, int64_t n,
                                                              int64_t min_cov, uint8_t *__restrict__ keep) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride)
        keep[r] = cov[r] > min_cov;
}

// dataload.py:211-216 -- per contig_mod group: number of rows and number with fraction > threshold.
__global__ void __launch_bounds__(256) group_counts_kernel(const int32_t *__restrict__ group,
                                                           const double *__restrict__ frac, int64_t n,
                                                           int n_groups, double thr,
                                                           unsigned long long *__restrict__ counts) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) {
        const int g = group[r];
        if (g < 0 || g >= n_groups) continue;
        // rows of one group are mostly consecutive: aggregate equal groups inside the warp first
        const unsigned peers = __match_any_sync(__activemask(), g);
        const int leader = __ffs(peers) - 1;
        const unsigned mods = __ballot_sync(peers, frac[r] > thr) & peers;
        if ((int)(threadIdx.x & 31) == leader) {
            atomicAdd(&counts[2 * g], (unsigned long long)__popc(peers));
            if (mods) atomicAdd(&counts[2 * g + 1], (unsigned long long)__popc(mods));
        }
    }
}

// dataload.py:217-224 -- keep rows of groups with n_mod / n_pos > min_frequency and n_mod > min_mods.
__global__ void __launch_bounds__(256) group_keep_kernel(const int32_t *__restrict__ group, int64_t n,
                                                         int n_groups,
                                                         const unsigned long long *__restrict__ counts,
                                                         double min_freq, long long min_mods,
                                                         uint8_t *__restrict__ keep) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) {
        const int g = group[r];
        uint8_t k = 0;
        if (g >= 0 && g < n_groups) {
            const double n_pos = (double)counts[2 * g], n_mod = (double)counts[2 * g + 1];
            k = ((n_mod / n_pos) > min_freq) && ((long long)counts[2 * g + 1] > min_mods);
        }
        keep[r] = k;
    }
}

// dataload.py:228-247 -- a row survives when its fraction equals the maximum fraction over the rows of
// the same (contig, strand), ANY mod type, with position in [p - d, p + d], or when it is below the
// methylation threshold.  Rows must be sorted by (contig, position); strands and mod types may interleave.
__global__ void __launch_bounds__(256) filter_adjacency_kernel(
    const int32_t *__restrict__ contig, const int64_t *__restrict__ pos, const uint8_t *__restrict__ strand,
    const double *__restrict__ frac, int64_t n, double thr, int dist, uint8_t *__restrict__ keep) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) {
        const double f = frac[r];
        if (f < thr) {  // low fractions never need the window
            keep[r] = 1;
            continue;
        }
        const int c = contig[r];
        const int64_t p = pos[r];
        const uint8_t s = strand[r];
        double mx = f;
        for (int64_t q = r - 1; q >= 0 && contig[q] == c && pos[q] >= p - dist; --q)
            if (strand[q] == s) mx = fmax(mx, frac[q]);
        for (int64_t q = r + 1; q < n && contig[q] == c && pos[q] <= p + dist; ++q)
            if (strand[q] == s) mx = fmax(mx, frac[q]);
        // NaN fractions: polars' max ignores nulls; a NaN row compares false on both tests and is dropped
        keep[r] = (f == mx) ? 1 : 0;
    }
}

// 1 when rows are sorted by (contig, position) ascending, else 0 (checked before filter_adjacency).
__global__ void __launch_bounds__(256) check_sorted_kernel(const int32_t *__restrict__ contig,
                                                           const int64_t *__restrict__ pos, int64_t n,
                                                           int *__restrict__ unsorted) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x + 1; r < n; r += stride)
        if (contig[r] < contig[r - 1] || (contig[r] == contig[r - 1] && pos[r] < pos[r - 1])) *unsorted = 1;
}

static inline unsigned blocks_for(int64_t n) {
    int64_t b = (n + 255) / 256;
    if (b > 148 * 32) b = 148 * 32;
    return (unsigned)(b < 1 ? 1 : b);
}

}  // namespace nmb

extern "C" {

int nmb_filter_coverage(const int64_t *n_valid_cov, int64_t n_rows, int64_t min_coverage, uint8_t *keep,
                        void *stream) {
    NMB_REQUIRE(n_rows >= 0, "nmb_filter_coverage: n_rows=%lld", (long long)n_rows);
    if (n_rows == 0) return NMB_OK;
    NMB_REQUIRE(n_valid_cov && keep, "nmb_filter_coverage: null argument");
    nmb::filter_coverage_kernel<<<nmb::blocks_for(n_rows), 256, 0, (cudaStream_t)stream>>>(n_valid_cov, n_rows,
                                                                                         min_coverage, keep);
    NMB_CUDA(cudaGetLastError());
    return NMB_OK;
}

int nmb_filter_min_mod_frequency(const int32_t *group_id, const double *fraction_mod, int64_t n_rows,
                                 int32_t n_groups, double methylation_threshold, double min_mod_frequency,
                                 int64_t min_mods_pr_contig, int64_t *group_counts, uint8_t *keep, void *stream) {
    NMB_REQUIRE(n_rows >= 0 && n_groups > 0, "nmb_filter_min_mod_frequency: n_rows=%lld n_groups=%d",
                (long long)n_rows, n_groups);
    NMB_REQUIRE(group_counts, "nmb_filter_min_mod_frequency: null scratch");
    cudaStream_t s = (cudaStream_t)stream;
    NMB_CUDA(cudaMemsetAsync(group_counts, 0, (size_t)n_groups * 2 * sizeof(int64_t), s));
    if (n_rows == 0) return NMB_OK;
    NMB_REQUIRE(group_id && fraction_mod && keep, "nmb_filter_min_mod_frequency: null argument");
    nmb::group_counts_kernel<<<nmb::blocks_for(n_rows), 256, 0, s>>>(group_id, fraction_mod, n_rows, n_groups,
                                                                    methylation_threshold,
                                                                    (unsigned long long *)group_counts);
    NMB_CUDA(cudaGetLastError());
    nmb::group_keep_kernel<<<nmb::blocks_for(n_rows), 256, 0, s>>>(group_id, n_rows, n_groups,
                                                                  (const unsigned long long *)group_counts,
                                                                  min_mod_frequency, (long long)min_mods_pr_contig,
                                                                  keep);
    NMB_CUDA(cudaGetLastError());
    return NMB_OK;
}

int nmb_filter_adjacency(const int32_t *contig_id, const int64_t *pos, const uint8_t *strand,
                         const double *fraction_mod, int64_t n_rows, double methylation_threshold,
                         int32_t adjacency_distance, uint8_t *keep, int32_t *unsorted_flag, void *stream) {
    NMB_REQUIRE(n_rows >= 0 && adjacency_distance >= 0, "nmb_filter_adjacency: bad sizes");
    NMB_REQUIRE(unsorted_flag, "nmb_filter_adjacency: null flag");
    cudaStream_t s = (cudaStream_t)stream;
    NMB_CUDA(cudaMemsetAsync(unsorted_flag, 0, sizeof(int32_t), s));
    if (n_rows == 0) return NMB_OK;
    NMB_REQUIRE(contig_id && pos && strand && fraction_mod && keep, "nmb_filter_adjacency: null argument");
    nmb::check_sorted_kernel<<<nmb::blocks_for(n_rows), 256, 0, s>>>(contig_id, pos, n_rows, unsorted_flag);
    NMB_CUDA(cudaGetLastError());
    nmb::filter_adjacency_kernel<<<nmb::blocks_for(n_rows), 256, 0, s>>>(
        contig_id, pos, strand, fraction_mod, n_rows, methylation_threshold, adjacency_distance, keep);
    NMB_CUDA(cudaGetLastError());
    return NMB_OK;
}

}  // extern "C"
