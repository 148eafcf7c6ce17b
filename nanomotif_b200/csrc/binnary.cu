// K9 -- binnary feature matrix straight from the K5 arrays on the device.
//
// Replaces the polars pipeline between epymetheus.methylation_pattern and the contamination / inclusion models:
//   main.py:192-193                             keep cells with n_motif_obs * mean_read_cov >= threshold
//   binnary/data_processing.py:174-187 add_bin  attach bins, drop unbinned contigs (:190-191)
//   :193-199                                    per (bin, motif_mod): sum(value * n_obs) / sum(n_obs)
//   :201-211                                    a contig with a kept cell gets every motif_mod of its bin: own value,
//                                               else the bin mean
//   :255-269 create_matrix                      pivot contig x motif_mod, missing -> 0
// Inputs are the dense K5 outputs stats [n_motifs][n_contigs][3] = {n_motif_obs, sum n_mod, sum n_valid_cov} and
// value [n_motifs][n_contigs] (median; NULL = weighted mean sum n_mod / sum n_valid_cov).  The row order (bin, contig
// name) and the feature order (sorted motif_mod strings) are string sorts and stay on the host, which gets two small
// flag arrays back between the two kernels.  Sums run in ascending contig order per (motif, bin) -- one thread each,
// no floating-point atomics -- so the result is bit-identical to the host restatement (nanomotif_b200/tables.py).
#include "common.cuh"

namespace nmb {

__device__ __forceinline__ bool cell_kept(const int64_t *st, double thr) {
    const int64_t obs = st[0];
    if (obs <= 0) return false;
    const double n = (double)obs;
    const double mean_cov = (double)st[2] / n;  // mean_read_cov
    return n * mean_cov >= thr;                 // main.py:192-193
}

__device__ __forceinline__ double cell_value(const int64_t *st, const double *value, int64_t i) {
    return value ? value[i] : (double)st[1] / (double)st[2];
}

// one thread per (motif, bin): walk the bin's contigs in ascending order
__global__ void __launch_bounds__(128) bin_means_kernel(const int64_t *__restrict__ stats, const double *__restrict__ value,
                                                        int n_motifs, int n_contigs, const int64_t *__restrict__ bin_off,
                                                        const int32_t *__restrict__ bin_contigs, int n_bins, double thr,
                                                        uint8_t *__restrict__ keep, double *__restrict__ bin_mean,
                                                        uint8_t *__restrict__ bin_has, uint8_t *__restrict__ contig_has) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)n_motifs * n_bins) return;
    const int m = (int)(t / n_bins), b = (int)(t - (int64_t)m * n_bins);
    double num = 0.0, den = 0.0;
    for (int64_t k = bin_off[b]; k < bin_off[b + 1]; ++k) {
        const int c = bin_contigs[k];
        const int64_t i = (int64_t)m * n_contigs + c;
        const int64_t *st = stats + i * 3;
        const bool kept = cell_kept(st, thr);
        keep[i] = kept;
        if (kept) {
            const double w = (double)st[0];
            num = __dadd_rn(num, __dmul_rn(cell_value(st, value, i), w));  // no FMA contraction: the host restatement
            den = __dadd_rn(den, w);                                        // (numpy) rounds the product, then the sum
            contig_has[c] = 1;  // benign race: every writer stores 1
        }
    }
    bin_has[t] = den > 0.0;
    bin_mean[t] = num / den;  // NaN for an empty (bin, motif): never read (bin_has = 0)
}

// matrix[r][f] for the selected rows / features
__global__ void __launch_bounds__(256) bin_matrix_kernel(const int64_t *__restrict__ stats, const double *__restrict__ value,
                                                         const uint8_t *__restrict__ keep, const double *__restrict__ bin_mean,
                                                         const uint8_t *__restrict__ bin_has,
                                                         const int32_t *__restrict__ contig_bin, int n_contigs, int n_bins,
                                                         const int32_t *__restrict__ rows, int64_t n_rows,
                                                         const int32_t *__restrict__ feats, int n_feat,
                                                         double *__restrict__ matrix) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_rows * n_feat) return;
    const int64_t r = t / n_feat;
    const int f = (int)(t - r * n_feat);
    const int c = rows[r], m = feats[f], b = contig_bin[c];
    const int64_t i = (int64_t)m * n_contigs + c;
    double v = 0.0;
    if (keep[i]) v = cell_value(stats + i * 3, value, i);
    else if (bin_has[(int64_t)m * n_bins + b]) v = bin_mean[(int64_t)m * n_bins + b];
    matrix[t] = v;
}

}  // namespace nmb

extern "C" {

int nmb_bin_means(const int64_t *stats, const double *value, int32_t n_motifs, int32_t n_contigs, const int64_t *bin_off,
                  const int32_t *bin_contigs, int32_t n_bins, double threshold, uint8_t *keep, double *bin_mean,
                  uint8_t *bin_has, uint8_t *contig_has, void *stream) {
    NMB_REQUIRE(n_motifs >= 0 && n_contigs >= 0 && n_bins >= 0, "nmb_bin_means: bad sizes");
    NMB_REQUIRE(keep && contig_has, "nmb_bin_means: null output");
    cudaStream_t s = (cudaStream_t)stream;
    NMB_CUDA(cudaMemsetAsync(keep, 0, (size_t)n_motifs * n_contigs, s));  // cells of unbinned contigs are not kept
    NMB_CUDA(cudaMemsetAsync(contig_has, 0, (size_t)n_contigs, s));
    const int64_t n = (int64_t)n_motifs * n_bins;
    if (n == 0) return NMB_OK;
    NMB_REQUIRE(stats && bin_off && bin_contigs && bin_mean && bin_has, "nmb_bin_means: null argument");
    nmb::bin_means_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(stats, value, n_motifs, n_contigs, bin_off, bin_contigs,
                                                                      n_bins, threshold, keep, bin_mean, bin_has, contig_has);
    NMB_CUDA(cudaGetLastError());
    return NMB_OK;
}

int nmb_bin_matrix(const int64_t *stats, const double *value, const uint8_t *keep, const double *bin_mean,
                   const uint8_t *bin_has, const int32_t *contig_bin, int32_t n_contigs, int32_t n_bins, const int32_t *rows,
                   int64_t n_rows, const int32_t *feats, int32_t n_feat, double *matrix, void *stream) {
    NMB_REQUIRE(n_rows >= 0 && n_feat >= 0, "nmb_bin_matrix: bad sizes");
    const int64_t n = n_rows * n_feat;
    if (n == 0) return NMB_OK;
    NMB_REQUIRE(stats && keep && bin_mean && bin_has && contig_bin && rows && feats && matrix, "nmb_bin_matrix: null argument");
    nmb::bin_matrix_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        stats, value, keep, bin_mean, bin_has, contig_bin, n_contigs, n_bins, rows, n_rows, feats, n_feat, matrix);
    NMB_CUDA(cudaGetLastError());
    return NMB_OK;
}

}  // extern "C"
