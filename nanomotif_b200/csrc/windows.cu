// K4 -- motif-growth step: window extraction, active-set filter + column histogram, PSSM + KL.
// Reference behaviour replaced:
//   DNAsequence.sample_at_indices / sample_at_index      nanomotif/seq.py:148-189
//   EqualLengthDNASet.reverse_compliment                 nanomotif/seq.py:387-389
//   EqualLengthDNASet.convert_to_DNAarray (one-hot)      nanomotif/seq.py:474-478
//   DNAarray.filter_sequence_matches                     nanomotif/seq.py:499-524
//   DNAarray.pssm                                        nanomotif/seq.py:526-537
//   scipy.stats.entropy(meth_pssm, bin_pssm)             nanomotif/find_motifs_bin.py:974
// A window is three 64-bit words (x bits, y bits, N bits; bit j = column j) instead of the
// reference's (W, 4) int64 one-hot rows (1312 bytes per window at W = 41).
#include "common.cuh"

namespace nmb {

__device__ __forceinline__ uint32_t seq_word(const uint32_t *seq_records, int plane, int64_t gw,
                                             int64_t n_words) {
    if (gw < 0 || gw >= n_words) return 0u;
    const int64_t tile = gw / kTileWords;
    return __ldg(seq_records + tile * kSeqRecWords + plane * kSeqPlaneWords + kHalo +
                 word_slot((int)(gw % kTileWords)));
}
__device__ __forceinline__ uint32_t n_word(const uint32_t *nonacgt, int64_t gw, int64_t n_words) {
    if (gw < 0 || gw >= n_words) return 0xFFFFFFFFu;
    return __ldg(nonacgt + kHalo + gw);
}
__device__ __forceinline__ uint64_t take_bits(uint32_t w0, uint32_t w1, uint32_t w2, int sh, int width) {
    const uint64_t lo = (uint64_t)w0 | ((uint64_t)w1 << 32);
    uint64_t v = lo >> sh;
    if (sh) v |= (uint64_t)w2 << (64 - sh);
    return v & ((1ull << width) - 1ull);
}

__global__ void __launch_bounds__(256) extract_windows_kernel(
    const uint32_t *__restrict__ seq_records, const uint32_t *__restrict__ nonacgt, int n_tiles,
    const int64_t *__restrict__ gpos, const uint8_t *__restrict__ strand, int64_t n, int padding,
    uint64_t *__restrict__ windows) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int width = 2 * padding + 1;
    const int64_t n_words = (int64_t)n_tiles * kTileWords;
    const int64_t s = gpos[i] - padding;  // first column on the forward strand
    const int64_t gw = s >> 5;            // arithmetic shift: floor for negative s
    const int sh = (int)(s & 31);
    uint64_t x = take_bits(seq_word(seq_records, 0, gw, n_words), seq_word(seq_records, 0, gw + 1, n_words),
                           seq_word(seq_records, 0, gw + 2, n_words), sh, width);
    uint64_t y = take_bits(seq_word(seq_records, 1, gw, n_words), seq_word(seq_records, 1, gw + 1, n_words),
                           seq_word(seq_records, 1, gw + 2, n_words), sh, width);
    uint64_t nn = take_bits(n_word(nonacgt, gw, n_words), n_word(nonacgt, gw + 1, n_words),
                            n_word(nonacgt, gw + 2, n_words), sh, width);
    if (strand[i]) {  // reverse complement: reverse the columns, A<->T and G<->C flip the low bit
        const int drop = 64 - width;
        x = __brevll(x) >> drop;
        nn = __brevll(nn) >> drop;
        y = (__brevll(~y & ((1ull << width) - 1ull)) >> drop) & ~nn;
    }
    windows[3 * i + 0] = x;
    windows[3 * i + 1] = y;
    windows[3 * i + 2] = nn;
}

// grid = (row blocks, n_motifs).  Column masks of the motif are built once per block in shared
// memory: allow[b] has bit j set when base b is accepted at column j.
// row_begin / row_end (may be NULL = all rows): rows examined by motif m.  keep_shared != 0: keep is one
// array of n flags shared by all motifs (their ranges are disjoint), else keep is [n_motifs][n].
__global__ void __launch_bounds__(256) window_hist_kernel(
    const uint64_t *__restrict__ windows, const uint8_t *__restrict__ alive, int64_t n, int width,
    const nmb_motif *__restrict__ masks, const int64_t *__restrict__ row_begin,
    const int64_t *__restrict__ row_end, int n_counts_all, int *__restrict__ hist,
    unsigned long long *__restrict__ n_active, uint8_t *__restrict__ keep, int keep_shared) {
    __shared__ uint64_t s_allow[4];
    __shared__ uint64_t s_wild;
    __shared__ int s_hist[NMB_MAX_WINDOW * 4];
    __shared__ int s_active;
    const int m = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid < 4) {
        uint64_t a = 0;
        for (int j = 0; j < width; ++j)
            if (masks[m].allowed[j] & (1 << tid)) a |= 1ull << j;
        s_allow[tid] = a;
    } else if (tid == 4) {
        uint64_t w = 0;
        for (int j = 0; j < width; ++j)
            if ((masks[m].allowed[j] & 0xF) == 0xF) w |= 1ull << j;
        s_wild = w;
        s_active = 0;
    }
    for (int i = tid; i < width * 4; i += 256) s_hist[i] = 0;
    __syncthreads();
    const uint64_t full = (width == 64) ? ~0ull : ((1ull << width) - 1ull);
    const uint64_t aA = s_allow[0], aT = s_allow[1], aG = s_allow[2], aC = s_allow[3], wild = s_wild;

    const int64_t r0 = row_begin ? row_begin[m] : 0, r1 = row_end ? row_end[m] : n;
    const int64_t base_rows = r0 + (int64_t)blockIdx.x * 256;
    if (base_rows >= r1) return;  // uniform per block (after the barrier above)
    const int64_t i = base_rows + tid;
    bool kept = false;
    uint64_t isA = 0, isT = 0, isG = 0, isC = 0;
    if (i < r1 && (!alive || alive[i])) {
        const uint64_t x = windows[3 * i], y = windows[3 * i + 1], nn = windows[3 * i + 2];
        const uint64_t acgt = ~nn & full;
        isA = ~x & ~y & acgt;
        isT = ~x & y & acgt;
        isG = x & ~y & acgt;
        isC = x & y & acgt;
        // one-hot(row) <= mask everywhere (seq.py:518); an N row is all ones and needs a wildcard
        const uint64_t bad = (isA & ~aA) | (isT & ~aT) | (isG & ~aG) | (isC & ~aC) | (nn & ~wild & full);
        kept = bad == 0;
        if (n_counts_all) {  // N adds one to all four bases (seq.py:41-48, 537)
            isA |= nn & full;
            isT |= nn & full;
            isG |= nn & full;
            isC |= nn & full;
        }
    }
    if (keep && i < r1) keep[(keep_shared ? 0 : (int64_t)m * n) + i] = kept ? 1 : 0;
    const uint32_t kb = __ballot_sync(0xFFFFFFFFu, kept);
    if (kb) {
        if (lane == 0) atomicAdd(&s_active, __popc(kb));
        for (int j = 0; j < width; ++j) {
            const uint32_t cA = __popc(__ballot_sync(0xFFFFFFFFu, kept && ((isA >> j) & 1)));
            const uint32_t cT = __popc(__ballot_sync(0xFFFFFFFFu, kept && ((isT >> j) & 1)));
            const uint32_t cG = __popc(__ballot_sync(0xFFFFFFFFu, kept && ((isG >> j) & 1)));
            const uint32_t cC = __popc(__ballot_sync(0xFFFFFFFFu, kept && ((isC >> j) & 1)));
            if (lane == 0) {
                if (cA) atomicAdd(&s_hist[j * 4 + 0], cA);
                if (cT) atomicAdd(&s_hist[j * 4 + 1], cT);
                if (cG) atomicAdd(&s_hist[j * 4 + 2], cG);
                if (cC) atomicAdd(&s_hist[j * 4 + 3], cC);
            }
        }
    }
    __syncthreads();
    for (int k = tid; k < width * 4; k += 256)
        if (s_hist[k]) atomicAdd(&hist[(int64_t)m * width * 4 + k], s_hist[k]);
    if (tid == 0 && s_active) atomicAdd(&n_active[m], (unsigned long long)s_active);
}

// One block per motif, one thread per column.
__global__ void __launch_bounds__(64) pssm_kl_kernel(const int *__restrict__ hist,
                                                     const long long *__restrict__ n_active, int width,
                                                     const double *__restrict__ bg, double *__restrict__ pssm,
                                                     double *__restrict__ kl) {
    const int m = blockIdx.x, j = threadIdx.x;
    if (j >= width) return;
    const double n = (double)n_active[m];
    double p[4], q[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        p[b] = (double)hist[((int64_t)m * width + j) * 4 + b] / n;  // seq.py:537
        pssm[((int64_t)m * 4 + b) * width + j] = p[b];
        q[b] = bg[b * width + j];
    }
    // scipy.stats.entropy: normalise both columns, sum rel_entr
    const double ps = ((p[0] + p[1]) + p[2]) + p[3];
    const double qs = ((q[0] + q[1]) + q[2]) + q[3];
    double acc = 0.0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const double pb = p[b] / ps, qb = q[b] / qs;
        double t;
        if (pb > 0.0 && qb > 0.0) t = pb * log(pb / qb);
        else if (pb == 0.0 && qb >= 0.0) t = 0.0;
        else if (pb != pb || qb != qb) t = pb + qb;  // NaN propagates
        else t = __longlong_as_double(0x7FF0000000000000ll);  // +inf
        acc += t;
    }
    kl[(int64_t)m * width + j] = acc;
}

}  // namespace nmb

extern "C" {

int nmb_extract_windows(const nmb_assembly *a, const int64_t *gpos, const uint8_t *strand, int64_t n,
                        int32_t padding, uint64_t *windows, void *stream) {
    NMB_REQUIRE(a, "nmb_extract_windows: null assembly");
    NMB_REQUIRE(padding >= 0 && 2 * padding + 1 <= NMB_MAX_WINDOW,
                "nmb_extract_windows: window width %d exceeds %d", 2 * padding + 1, NMB_MAX_WINDOW);
    NMB_REQUIRE(n >= 0, "nmb_extract_windows: n=%lld", (long long)n);
    if (n == 0) return NMB_OK;
    NMB_REQUIRE(gpos && strand && windows, "nmb_extract_windows: null argument");
    nmb::extract_windows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        a->seq_records, a->nonacgt, a->n_tiles, gpos, strand, n, padding, windows);
    NMB_CUDA(cudaGetLastError());
    return NMB_OK;
}

int nmb_window_hist(const uint64_t *windows, const uint8_t *alive, int64_t n, int32_t width,
                    const nmb_motif *masks, int32_t n_motifs, int32_t n_counts_all, int32_t *hist,
                    int64_t *n_active, uint8_t *keep, void *stream) {
    NMB_REQUIRE(width >= 1 && width <= NMB_MAX_WINDOW, "nmb_window_hist: width=%d", width);
    NMB_REQUIRE(n >= 0 && n_motifs >= 0, "nmb_window_hist: n=%lld n_motifs=%d", (long long)n, n_motifs);
    if (n_motifs == 0) return NMB_OK;
    NMB_REQUIRE(masks && hist && n_active, "nmb_window_hist: null argument");
    NMB_REQUIRE(n_motifs <= 65535, "nmb_window_hist: at most 65535 motifs per call");
    cudaStream_t s = (cudaStream_t)stream;
    NMB_CUDA(cudaMemsetAsync(hist, 0, (size_t)n_motifs * width * 4 * sizeof(int32_t), s));
    NMB_CUDA(cudaMemsetAsync(n_active, 0, (size_t)n_motifs * sizeof(int64_t), s));
    if (n == 0) return NMB_OK;
    NMB_REQUIRE(windows, "nmb_window_hist: null windows");
    dim3 grid((unsigned)((n + 255) / 256), (unsigned)n_motifs);
    nmb::window_hist_kernel<<<grid, 256, 0, s>>>(windows, alive, n, width, masks, nullptr, nullptr, n_counts_all,
                                                hist, (unsigned long long *)n_active, keep, 0);
    NMB_CUDA(cudaGetLastError());
    return NMB_OK;
}

int nmb_window_hist_ranges(const uint64_t *windows, const uint8_t *alive, int64_t n, int32_t width,
                           const nmb_motif *masks, int32_t n_motifs, const int64_t *row_begin,
                           const int64_t *row_end, int64_t max_rows, int32_t n_counts_all, int32_t *hist,
                           int64_t *n_active, uint8_t *keep_rows, void *stream) {
    NMB_REQUIRE(width >= 1 && width <= NMB_MAX_WINDOW, "nmb_window_hist_ranges: width=%d", width);
    NMB_REQUIRE(n >= 0 && n_motifs >= 0 && max_rows >= 0, "nmb_window_hist_ranges: bad sizes");
    if (n_motifs == 0) return NMB_OK;
    NMB_REQUIRE(masks && hist && n_active && row_begin && row_end, "nmb_window_hist_ranges: null argument");
    NMB_REQUIRE(n_motifs <= 65535, "nmb_window_hist_ranges: at most 65535 motifs per call");
    cudaStream_t s = (cudaStream_t)stream;
    NMB_CUDA(cudaMemsetAsync(hist, 0, (size_t)n_motifs * width * 4 * sizeof(int32_t), s));
    NMB_CUDA(cudaMemsetAsync(n_active, 0, (size_t)n_motifs * sizeof(int64_t), s));
    if (n == 0 || max_rows == 0) return NMB_OK;
    NMB_REQUIRE(windows, "nmb_window_hist_ranges: null windows");
    dim3 grid((unsigned)((max_rows + 255) / 256), (unsigned)n_motifs);
    nmb::window_hist_kernel<<<grid, 256, 0, s>>>(windows, alive, n, width, masks, row_begin, row_end,
                                                n_counts_all, hist, (unsigned long long *)n_active, keep_rows, 1);
    NMB_CUDA(cudaGetLastError());
    return NMB_OK;
}

int nmb_pssm_kl(const int32_t *hist, const int64_t *n_active, int32_t n_motifs, int32_t width,
                const double *bg_pssm, double *pssm, double *kl, void *stream) {
    NMB_REQUIRE(width >= 1 && width <= NMB_MAX_WINDOW, "nmb_pssm_kl: width=%d", width);
    NMB_REQUIRE(n_motifs >= 0, "nmb_pssm_kl: n_motifs=%d", n_motifs);
    if (n_motifs == 0) return NMB_OK;
    NMB_REQUIRE(hist && n_active && bg_pssm && pssm && kl, "nmb_pssm_kl: null argument");
    nmb::pssm_kl_kernel<<<n_motifs, 64, 0, (cudaStream_t)stream>>>(
        hist, (const long long *)n_active, width, bg_pssm, pssm, kl);
    NMB_CUDA(cudaGetLastError());
    return NMB_OK;
}

}  // extern "C"
