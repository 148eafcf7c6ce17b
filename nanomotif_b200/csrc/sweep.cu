// K8 -- exhaustive candidate sweep (BASELINE.json configs[4]): the counts of EVERY IUPAC motif of length 4..8
// at every modified position, without scanning per motif.
//
// Brute force is motif x bp work: 15^4 + ... + 15^8 = 2.7e9 motifs over a 2 Gbp assembly is ~5e18 motif*bp.  The
// counts that motif_model_bin returns (find_motifs_bin.py:1265-1331: pileup rows classified methylated /
// unmethylated whose position is the modified base of an occurrence, both strands) are additive over the
// CONCRETE sequence contexts of the rows, so one pass over the assembly suffices:
//   sweep_hist_kernel    for every window of 8 letters inside a contig and every classified row under it:
//                        hist[8][o][class][window] += 1, o = offset of the row in the window ('+' rows use
//                        the window, '-' rows its reverse complement with the offset mirrored -- the reference
//                        scans the reverse-complement motif on the forward text, find_motifs_bin.py:1317).
//                        Letters have FIVE states: A, T, G, C and "other" (non-ACGT), because the regex
//                        wildcard matches any character while sets do not.
//                        Only k = 8 is counted row by row (8 REDs per row instead of 4+5+6+7+8 = 30): a motif of
//                        k < 8 letters is the PREFIX of the (k+1)-letter motifs that extend it by one letter, so
//                            hist[k][o][c][w] = sum_j hist[k+1][o][c][5 w + j]  +  the k-windows that have no
//                                                                                  (k+1)-extension inside the contig
//                        ('+' rows: the one window that ends at the contig end; '-' rows, whose motif is the
//                        reverse complement and therefore extends to the LEFT in contig coordinates: the one window
//                        that starts at the contig start).  The kernel adds those edge windows directly (a handful
//                        per contig) and sweep_marginalise_kernel folds the levels 8 -> 7 -> ... -> 4 afterwards
//                        (nmb_sweep_finalize).
//   sweep_expand_kernel  subset-sum (zeta) transform, one axis at a time: 5 letter states -> 15 IUPAC letters
//                        (N includes "other").  After all axes, table[L_0 .. L_{k-1}] is the count of motif L.
//   sweep_filter_kernel  posterior-mean / support filter over a finished table -> candidate list.
// The modified position's own letter is fixed to one concrete base by the caller (a row under any other base
// cannot exist), which keeps the largest table at 15^7 entries.
#include "scan.cuh"

namespace nmb {

constexpr int kSweepMinK = 4, kSweepMaxK = 8;
__host__ __device__ constexpr int pow5(int k) { return k == 0 ? 1 : 5 * pow5(k - 1); }
// first counter of length k in the histogram array: sum_{j<k} j * 2 * 5^j
__host__ __device__ constexpr int64_t sweep_hist_offset(int k) {
    return k <= kSweepMinK ? 0 : sweep_hist_offset(k - 1) + (int64_t)(k - 1) * 2 * pow5(k - 1);
}
constexpr int64_t kSweepHistSize = sweep_hist_offset(kSweepMaxK + 1);  // 7 567 500 counters

struct SweepParams {
    const uint32_t *seq_records, *nonacgt, *cls;  // cls: the class records of ONE mod type
    const int64_t *contig_start, *contig_len;
    uint32_t *hist;
    int tile_begin, n_tiles_total, contig_begin, contig_end;
};

// word w (0..16) of chunk t in a lane-interleaved plane; word 16 = first word of chunk t + 1
// (`next` supplies it for the last chunk of the tile)
__device__ __forceinline__ uint32_t chunk_word(const uint32_t *body, int t, int w, uint32_t next) {
    if (w < NW) return body[(w >> 2) * kSlotStride + t * 4 + (w & 3)];
    return t + 1 < kTileChunks ? body[(t + 1) * 4] : next;
}

template <int K>
__device__ __forceinline__ void sweep_add(uint32_t *hist, int code8, int rc8, unsigned plus, unsigned minus, int cls,
                                          int rem) {
    if (rem < K) return;  // the window of K letters must lie inside the contig
    uint32_t *h = hist + sweep_hist_offset(K);
    const int code = code8 / pow5(kSweepMaxK - K);  // first K letters
    const int rc = rc8 % pow5(K);                   // reverse complement of the first K letters
    for (unsigned b = plus & ((1u << K) - 1u); b; b &= b - 1) {
        const int o = __ffs(b) - 1;
        atomicAdd(h + (size_t)(o * 2 + cls) * pow5(K) + code, 1u);
    }
    for (unsigned b = minus & ((1u << K) - 1u); b; b &= b - 1) {
        const int o = K - 1 - (__ffs(b) - 1);  // offset of the row in the reverse-complement window
        atomicAdd(h + (size_t)(o * 2 + cls) * pow5(K) + rc, 1u);
    }
}

// ---- bipartite shapes X{3,4} N{4..8} Y{3,4} over ACGT (the second half of BASELINE config 5) ----
// Shape (a, g, b): a concrete letters, g wildcards, b concrete letters.  Block of shape (a, g, b):
// [o = 0..a+b-1][class][4^(a+b)], o counted over the concrete letters only; window index =
// xL | yL << a | xR << 2a | yR << (2a + b) with x / y the high / low code bits of the letters, letter i at bit i.
constexpr int kBipMinGap = 4, kBipMaxGap = 8;
__host__ __device__ constexpr int64_t bip_block(int a, int b) { return (int64_t)(a + b) * 2 * (1 << (2 * (a + b))); }
__host__ __device__ constexpr int64_t bip_offset(int a, int g, int b) {
    // order: g major, then (3,3) (3,4) (4,3) (4,4)
    return (int64_t)(g - kBipMinGap) * (bip_block(3, 3) + bip_block(3, 4) + bip_block(4, 3) + bip_block(4, 4)) +
           (a == 3 ? (b == 3 ? 0 : bip_block(3, 3)) : bip_block(3, 3) + bip_block(3, 4) + (b == 3 ? 0 : bip_block(4, 3)));
}
constexpr int64_t kBipHistSize = bip_offset(3, kBipMaxGap + 1, 3);

__device__ __forceinline__ unsigned rev_bits(unsigned v, int n) { return __brev(v) >> (32 - n); }

// Window of shape (A, G, B) starting at bit `s` of the 64-bit planes: '+' rows count under the window, '-' rows
// under its reverse complement, whose shape is (B, G, A).
template <int A, int G, int B>
__device__ __forceinline__ void bip_add(uint32_t *hist, uint64_t X, uint64_t Y, uint64_t N, uint64_t plus, uint64_t minus,
                                        int s, int cls, int rem, bool do_plus = true, bool do_minus = true) {
    constexpr int span = A + G + B;
    if (rem < span) return;
    constexpr unsigned ma = (1u << A) - 1u, mb = (1u << B) - 1u;
    const unsigned nl = (unsigned)(N >> s) & ma, nr = (unsigned)(N >> (s + A + G)) & mb;
    if (nl | nr) return;  // a concrete letter cannot match a non-ACGT letter
    const unsigned pl = do_plus ? (unsigned)(plus >> s) & ma : 0u, pr = do_plus ? (unsigned)(plus >> (s + A + G)) & mb : 0u;
    const unsigned ql = do_minus ? (unsigned)(minus >> s) & ma : 0u, qr = do_minus ? (unsigned)(minus >> (s + A + G)) & mb : 0u;
    if (!(pl | pr | ql | qr)) return;
    const unsigned xl = (unsigned)(X >> s) & ma, yl = (unsigned)(Y >> s) & ma;
    const unsigned xr = (unsigned)(X >> (s + A + G)) & mb, yr = (unsigned)(Y >> (s + A + G)) & mb;
    if (pl | pr) {
        uint32_t *h = hist + bip_offset(A, G, B);
        const unsigned code = xl | (yl << A) | (xr << (2 * A)) | (yr << (2 * A + B));
        for (unsigned b = pl | (pr << A); b; b &= b - 1) {
            const int o = __ffs(b) - 1;
            atomicAdd(h + (size_t)(o * 2 + cls) * (1u << (2 * (A + B))) + code, 1u);
        }
    }
    if (ql | qr) {  // the motif is the reverse complement of the window: shape (B, G, A), letters reversed, A<->T G<->C
        uint32_t *h = hist + bip_offset(B, G, A);
        const unsigned code = rev_bits(xr, B) | (rev_bits(~yr & mb, B) << B) | (rev_bits(xl, A) << (2 * B)) |
                              (rev_bits(~yl & ma, A) << (2 * B + A));
        // forward offset j (left part) -> motif offset B + (A - 1 - j); right part j -> B - 1 - j
        for (unsigned b = ql; b; b &= b - 1) {
            const int o = B + (A - 1 - (__ffs(b) - 1));
            atomicAdd(h + (size_t)(o * 2 + cls) * (1u << (2 * (A + B))) + code, 1u);
        }
        for (unsigned b = qr; b; b &= b - 1) {
            const int o = B - 1 - (__ffs(b) - 1);
            atomicAdd(h + (size_t)(o * 2 + cls) * (1u << (2 * (A + B))) + code, 1u);
        }
    }
}

// Only the shape (4, G, 4) is counted row by row.  The three shorter shapes are prefixes / suffixes of it at the MOTIF
// level -- (4,G,3) drops the motif's last letter, (3,G,4) its first, (3,G,3) then the last letter of (3,G,4) -- so
// nmb_sweep_bipartite_finalize gets them by summing over the dropped letter (4 concrete letters).  A window is counted
// directly only when the letter that would extend it is not a concrete letter of the same contig (non-ACGT, or past
// the contig's edge: inter-contig padding is flagged in the non-ACGT plane), because then no longer window covers it.
// '+' rows: motif = window, so "last letter" extends to the right and "first" to the left; '-' rows: motif = reverse
// complement, so the directions swap.  left_n = the letter left of the window start is not concrete.
template <int G>
__device__ __forceinline__ void bip_add_gap(uint32_t *hist, uint64_t X, uint64_t Y, uint64_t N, uint64_t plus,
                                            uint64_t minus, int s, int cls, int rem, bool left_n) {
    bip_add<4, G, 4>(hist, X, Y, N, plus, minus, s, cls, rem);
    const bool right4_n = (N >> (s + 7 + G)) & 1;  // the letter after a (4,G,3) window
    const bool right3_n = (N >> (s + 6 + G)) & 1;  // the letter after a (3,G,3) / before-last of (3,G,4)
    // fwd (4,G,3): '+' motif (4,G,3) and '-' motif (3,G,4) both extend to the right
    if (right4_n) bip_add<4, G, 3>(hist, X, Y, N, plus, minus, s, cls, rem);
    // fwd (3,G,4): '+' motif (3,G,4) and '-' motif (4,G,3) both extend to the left
    if (left_n) bip_add<3, G, 4>(hist, X, Y, N, plus, minus, s, cls, rem);
    // fwd (3,G,3): '+' motif comes from (3,G,4) at the same start (right extension), '-' motif from the reverse
    // complement of fwd (4,G,3) one letter to the left
    if (right3_n | left_n) bip_add<3, G, 3>(hist, X, Y, N, plus, minus, s, cls, rem, right3_n, left_n);
}

// One CTA per tile, one lane per 512-bp chunk, positions in order: the 8-letter window code slides by one letter
// per position (base-5 digits, most significant first; the reverse-complement code slides the other way).
template <bool BIPARTITE>
__global__ void __launch_bounds__(kTileChunks) sweep_hist_kernel(const SweepParams p) {
    // The records are read straight from global memory / L2 (the lane-interleaved layout keeps the lanes' words
    // adjacent): every word is used once, and without a 50 KB tile per CTA in shared memory three times as many
    // warps are resident to cover the latency of the divergent REDs (ncu on the staged version: issue-active 16 %,
    // long-scoreboard 16 per issue).
    const int tid = threadIdx.x;
    const int tile = p.tile_begin + blockIdx.x;
    const uint32_t *sx = p.seq_records + (size_t)tile * kSeqRecWords;
    const uint32_t *sy = sx + kSeqPlaneWords;
    const int info = reinterpret_cast<const int32_t *>(sy + kSeqPlaneWords)[tid];
    if (info < 0) return;
    const int contig = info & kChunkIdMask;
    if (contig < p.contig_begin || contig >= p.contig_end) return;
    const uint32_t *scls = p.cls + (size_t)tile * kClsRecWords;
    const int64_t chunk_pos = ((int64_t)tile * kTileChunks + tid) * NMB_CHUNK_BP;
    const int64_t contig_start_pos = __ldg(p.contig_start + contig);
    const int64_t contig_end_pos = contig_start_pos + __ldg(p.contig_len + contig);
    const int n_here = (int)min((int64_t)NMB_CHUNK_BP, contig_end_pos - chunk_pos);  // window starts in this chunk
    // word 16 of every plane: the record halo / the next tile's class record for the last chunk
    const bool last_chunk = tid + 1 == kTileChunks;
    const bool has_next = tile + 1 < p.n_tiles_total;
    uint32_t next_cls[4] = {0, 0, 0, 0};
    if (last_chunk && has_next)
        for (int k = 0; k < 4; ++k) next_cls[k] = __ldg(p.cls + (size_t)(tile + 1) * kClsRecWords + k * kTileWords);
    const uint32_t next_x = sx[kHalo + kTileWords], next_y = sy[kHalo + kTileWords];
    const uint32_t *gn = p.nonacgt + kHalo + (size_t)tile * kTileWords + tid * NW;

    uint64_t X = 0, Y = 0, N = 0, C0 = 0, C1 = 0, C2 = 0, C3 = 0;  // bit i = position (32 * word + i) of the current pair
    auto load_pair = [&](int w) {  // words w and w + 1 of the seven planes
        const uint64_t x0 = chunk_word(sx + kHalo, tid, w, next_x), x1 = w + 1 <= NW ? chunk_word(sx + kHalo, tid, w + 1, next_x) : 0;
        const uint64_t y0 = chunk_word(sy + kHalo, tid, w, next_y), y1 = w + 1 <= NW ? chunk_word(sy + kHalo, tid, w + 1, next_y) : 0;
        X = x0 | (x1 << 32);
        Y = y0 | (y1 << 32);
        N = (uint64_t)__ldg(gn + w) | ((uint64_t)(w + 1 <= NW ? __ldg(gn + w + 1) : 0xFFFFFFFFu) << 32);
        const uint64_t a0 = chunk_word(scls, tid, w, next_cls[0]), a1 = w + 1 <= NW ? chunk_word(scls, tid, w + 1, next_cls[0]) : 0;
        const uint64_t b0 = chunk_word(scls + kTileWords, tid, w, next_cls[1]), b1 = w + 1 <= NW ? chunk_word(scls + kTileWords, tid, w + 1, next_cls[1]) : 0;
        const uint64_t c0 = chunk_word(scls + 2 * kTileWords, tid, w, next_cls[2]), c1 = w + 1 <= NW ? chunk_word(scls + 2 * kTileWords, tid, w + 1, next_cls[2]) : 0;
        const uint64_t d0 = chunk_word(scls + 3 * kTileWords, tid, w, next_cls[3]), d1 = w + 1 <= NW ? chunk_word(scls + 3 * kTileWords, tid, w + 1, next_cls[3]) : 0;
        C0 = a0 | (a1 << 32);
        C1 = b0 | (b1 << 32);
        C2 = c0 | (c1 << 32);
        C3 = d0 | (d1 << 32);
    };
    auto letter = [&](int i) -> int {  // state of position i (0..63) of the current pair: A T G C other
        return ((N >> i) & 1) ? 4 : (int)((((X >> i) & 1) << 1) | ((Y >> i) & 1));
    };
    auto comp = [](int d) -> int { return d < 4 ? (d ^ 1) : 4; };  // A<->T, G<->C

    load_pair(0);
    bool left_n = (__ldg(gn - 1) >> 31) & 1;  // the letter before the chunk (padding and pad words are flagged)
    int code8 = 0, rc8 = 0, first = 0;  // first = the window's leading letter (leaves on the next slide)
    for (int i = 0; i < kSweepMaxK; ++i) {
        const int d = letter(i);
        code8 = code8 * 5 + d;
        rc8 += comp(d) * pow5(i);
        if (i == 0) first = d;
    }
    for (int s = 0; s < n_here; ++s) {
        const int b = s & 31;
        if (b == 0 && s) load_pair(s >> 5);
        if (BIPARTITE) {  // spans of 10..16 letters: bits b .. b + 15 of the pairs
            const int rem16 = (int)min((int64_t)16, contig_end_pos - (chunk_pos + s));
            if ((C0 | C2) >> b & 0xFFFF) {
                bip_add_gap<4>(p.hist, X, Y, N, C0, C2, b, 0, rem16, left_n);
                bip_add_gap<5>(p.hist, X, Y, N, C0, C2, b, 0, rem16, left_n);
                bip_add_gap<6>(p.hist, X, Y, N, C0, C2, b, 0, rem16, left_n);
                bip_add_gap<7>(p.hist, X, Y, N, C0, C2, b, 0, rem16, left_n);
                bip_add_gap<8>(p.hist, X, Y, N, C0, C2, b, 0, rem16, left_n);
            }
            if ((C1 | C3) >> b & 0xFFFF) {
                bip_add_gap<4>(p.hist, X, Y, N, C1, C3, b, 1, rem16, left_n);
                bip_add_gap<5>(p.hist, X, Y, N, C1, C3, b, 1, rem16, left_n);
                bip_add_gap<6>(p.hist, X, Y, N, C1, C3, b, 1, rem16, left_n);
                bip_add_gap<7>(p.hist, X, Y, N, C1, C3, b, 1, rem16, left_n);
                bip_add_gap<8>(p.hist, X, Y, N, C1, C3, b, 1, rem16, left_n);
            }
            left_n = (N >> b) & 1;
            continue;
        }
        const int rem = (int)min((int64_t)kSweepMaxK, contig_end_pos - (chunk_pos + s));
        const unsigned mp = (unsigned)(C0 >> b) & 0xFF, np = (unsigned)(C1 >> b) & 0xFF;  // rows under the window, '+'
        const unsigned mm = (unsigned)(C2 >> b) & 0xFF, nm = (unsigned)(C3 >> b) & 0xFF;  // '-'
        if (rem >= kSweepMaxK) {  // the common case: the 8-window lies inside the contig
            if (mp | mm) sweep_add<8>(p.hist, code8, rc8, mp, mm, 0, rem);
            if (np | nm) sweep_add<8>(p.hist, code8, rc8, np, nm, 1, rem);
        }
        // k < 8 comes from marginalising k + 1 (nmb_sweep_finalize) except for the windows without an extension inside
        // the contig: '+' rows under the k-window that ENDS at the contig end (rem == k), '-' rows under the k-window
        // that STARTS at the contig start
        const bool at_start = s == 0 && chunk_pos == contig_start_pos;
        if (rem < kSweepMaxK || at_start) {
#define NMB_EDGE(K)                                                                                              \
            {                                                                                                    \
                const unsigned pm = rem == K ? mp : 0u, pn = rem == K ? np : 0u;                                 \
                const unsigned qm = at_start ? mm : 0u, qn = at_start ? nm : 0u;                                 \
                if (pm | qm) sweep_add<K>(p.hist, code8, rc8, pm, qm, 0, rem);                                   \
                if (pn | qn) sweep_add<K>(p.hist, code8, rc8, pn, qn, 1, rem);                                   \
            }
            NMB_EDGE(4) NMB_EDGE(5) NMB_EDGE(6) NMB_EDGE(7)
#undef NMB_EDGE
        }
        // slide: drop the leading letter, append the letter at s + 8
        const int d_new = letter(b + kSweepMaxK);
        code8 = (code8 - first * pow5(kSweepMaxK - 1)) * 5 + d_new;
        rc8 = (rc8 - comp(first)) / 5 + comp(d_new) * pow5(kSweepMaxK - 1);
        first = letter(b + 1);
    }
}

// hist[k][o][c][w] += sum_{j < 5} hist[k + 1][o][c][5 w + j] for o < k: a k-letter motif is the prefix of its five
// one-letter extensions (the trailing letter is the least significant base-5 digit).
__global__ void __launch_bounds__(256) sweep_marginalise_kernel(uint32_t *__restrict__ hist, int k) {
    int n5 = 1;
    for (int i = 0; i < k; ++i) n5 *= 5;
    const int64_t n = (int64_t)k * 2 * n5;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t oc = i / n5, w = i - oc * n5;  // oc = o * 2 + class, the same index in both levels (o < k)
    int64_t off_k = 0, off_k1 = 0;               // sweep_hist_offset(k), sweep_hist_offset(k + 1)
    for (int j = kSweepMinK, p5 = pow5(kSweepMinK); j <= k; ++j, p5 *= 5) {
        if (j < k) off_k += (int64_t)j * 2 * p5;
        off_k1 += (int64_t)j * 2 * p5;
    }
    const uint32_t *src = hist + off_k1 + oc * (int64_t)n5 * 5 + w * 5;
    hist[off_k + i] += src[0] + src[1] + src[2] + src[3] + src[4];
}

// Bipartite marginals (see bip_add_gap).  mode 0: shape (a, b) += sum over the LAST right letter of shape (a, b + 1),
// same offsets; mode 1: shape (a, b) += sum over the FIRST left letter of shape (a + 1, b), offset o <- o + 1.
__global__ void __launch_bounds__(256) bip_marginalise_kernel(uint32_t *__restrict__ hist, int g, int a, int b, int mode) {
    const int n_code = 1 << (2 * (a + b));
    const int64_t n = (int64_t)(a + b) * 2 * n_code;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int oc = (int)(i / n_code), code = (int)(i - (int64_t)oc * n_code);
    const int o = oc >> 1, cls = oc & 1;
    const unsigned xl = code & ((1u << a) - 1), yl = (code >> a) & ((1u << a) - 1);
    const unsigned xr = (code >> (2 * a)) & ((1u << b) - 1), yr = (code >> (2 * a + b)) & ((1u << b) - 1);
    const int sa = mode ? a + 1 : a, sb = mode ? b : b + 1;
    const uint32_t *src = hist + bip_offset(sa, g, sb) + (size_t)(((mode ? o + 1 : o) * 2 + cls)) * (1u << (2 * (sa + sb)));
    uint32_t sum = 0;
    for (unsigned t = 0; t < 4; ++t) {
        const unsigned tx = t >> 1, ty = t & 1;
        unsigned sxl = xl, syl = yl, sxr = xr, syr = yr;
        if (mode) { sxl = (xl << 1) | tx; syl = (yl << 1) | ty; }
        else { sxr = xr | (tx << b); syr = yr | (ty << b); }
        sum += src[sxl | (syl << sa) | (sxr << (2 * sa)) | (syr << (2 * sa + sb))];
    }
    hist[bip_offset(a, g, b) + i] += sum;
}

// dst[outer][15][inner] = subset sums of src[outer][5][inner] over the letters of each IUPAC code.
// Letter states: A T G C other; IUPAC order = nanomotif/constants.py:2 (A T G C R Y S W K M B D H V N).
__global__ void __launch_bounds__(256) sweep_expand_kernel(const uint32_t *__restrict__ src, uint32_t *__restrict__ dst,
                                                           int64_t outer, int64_t inner) {
    const int64_t n = outer * inner;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int64_t o = i / inner, r = i - o * inner;
        const uint32_t *s = src + o * 5 * inner + r;
        const uint32_t a = s[0], t = s[inner], g = s[2 * inner], c = s[3 * inner], x = s[4 * inner];
        uint32_t *d = dst + o * 15 * inner + r;
        d[0] = a; d[inner] = t; d[2 * inner] = g; d[3 * inner] = c;
        d[4 * inner] = a + g;           // R
        d[5 * inner] = c + t;           // Y
        d[6 * inner] = g + c;           // S
        d[7 * inner] = a + t;           // W
        d[8 * inner] = g + t;           // K
        d[9 * inner] = a + c;           // M
        d[10 * inner] = c + g + t;      // B
        d[11 * inner] = a + g + t;      // D
        d[12 * inner] = a + c + t;      // H
        d[13 * inner] = a + c + g;      // V
        d[14 * inner] = a + c + g + t + x;  // N: the regex wildcard also matches non-ACGT letters
    }
}

// dst[...] = src[... with the digit of axis `axis_stride` fixed to `digit` ...]: drops one base-5 axis.
__global__ void __launch_bounds__(256) sweep_slice_kernel(const uint32_t *__restrict__ src, uint32_t *__restrict__ dst,
                                                          int64_t n_out, int64_t axis_stride, int digit) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_out) return;
    const int64_t hi = i / axis_stride, lo = i - hi * axis_stride;
    dst[i] = src[(hi * 5 + digit) * axis_stride + lo];
}

// Candidates of one finished table pair: posterior mean (5 + n_mod) / (10 + n_mod + n_nomod) >= min_mean and
// n_mod >= min_mod (Beta(5, 5) prior, nanomotif/model.py:8-9).  Appends the table index; *n_out counts all hits.
__global__ void __launch_bounds__(256) sweep_filter_kernel(const uint32_t *__restrict__ n_mod,
                                                           const uint32_t *__restrict__ n_nomod, int64_t n,
                                                           double min_mean, uint32_t min_mod, int64_t *__restrict__ out,
                                                           int64_t capacity, unsigned long long *__restrict__ n_out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t m = n_mod[i];
        if (m < min_mod) continue;
        const double mean = (5.0 + m) / (10.0 + m + n_nomod[i]);
        if (mean < min_mean) continue;
        const unsigned long long slot = atomicAdd(n_out, 1ull);
        if ((int64_t)slot < capacity) out[slot] = i;
    }
}

}  // namespace nmb

extern "C" {

int64_t nmb_sweep_hist_size(void) { return nmb::kSweepHistSize; }

int64_t nmb_sweep_bipartite_size(void) { return nmb::kBipHistSize; }

static int sweep_launch(bool bipartite, const nmb_assembly *a, const uint32_t *class_records_of_modtype,
                        int32_t tile_begin, int32_t tile_count, int32_t contig_begin, int32_t contig_end, uint32_t *hist,
                        void *stream);

int nmb_sweep_hist(const nmb_assembly *a, const uint32_t *class_records_of_modtype, int32_t tile_begin,
                   int32_t tile_count, int32_t contig_begin, int32_t contig_end, uint32_t *hist, void *stream) {
    return sweep_launch(false, a, class_records_of_modtype, tile_begin, tile_count, contig_begin, contig_end, hist, stream);
}

int nmb_sweep_bipartite(const nmb_assembly *a, const uint32_t *class_records_of_modtype, int32_t tile_begin,
                        int32_t tile_count, int32_t contig_begin, int32_t contig_end, uint32_t *hist, void *stream) {
    return sweep_launch(true, a, class_records_of_modtype, tile_begin, tile_count, contig_begin, contig_end, hist, stream);
}

static int sweep_launch(bool bipartite, const nmb_assembly *a, const uint32_t *class_records_of_modtype,
                        int32_t tile_begin, int32_t tile_count, int32_t contig_begin, int32_t contig_end, uint32_t *hist,
                        void *stream) {
    NMB_REQUIRE(a && class_records_of_modtype && hist, "nmb_sweep_hist: null argument");
    NMB_REQUIRE(tile_begin >= 0 && tile_count >= 0 && tile_begin + tile_count <= a->n_tiles,
                "nmb_sweep_hist: tiles [%d, %d) outside the assembly", tile_begin, tile_begin + tile_count);
    if (tile_count == 0) return NMB_OK;
    nmb::SweepParams p{a->seq_records, a->nonacgt, class_records_of_modtype, a->contig_start, a->contig_len, hist,
                       tile_begin, a->n_tiles, contig_begin, contig_end};
    if (bipartite)
        nmb::sweep_hist_kernel<true><<<tile_count, nmb::kTileChunks, 0, (cudaStream_t)stream>>>(p);
    else
        nmb::sweep_hist_kernel<false><<<tile_count, nmb::kTileChunks, 0, (cudaStream_t)stream>>>(p);
    NMB_CUDA(cudaGetLastError());
    return NMB_OK;
}

int nmb_sweep_finalize(uint32_t *hist, void *stream) {
    NMB_REQUIRE(hist, "nmb_sweep_finalize: null argument");
    for (int k = nmb::kSweepMaxK - 1; k >= nmb::kSweepMinK; --k) {  // 8 -> 7, then 7 -> 6, ... each level complete first
        const int64_t n = (int64_t)k * 2 * nmb::pow5(k);
        nmb::sweep_marginalise_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(hist, k);
        NMB_CUDA(cudaGetLastError());
    }
    return NMB_OK;
}

int nmb_sweep_bipartite_finalize(uint32_t *hist, void *stream) {
    NMB_REQUIRE(hist, "nmb_sweep_bipartite_finalize: null argument");
    auto run = [&](int g, int a, int b, int mode) {
        const int64_t n = (int64_t)(a + b) * 2 * (1ll << (2 * (a + b)));
        nmb::bip_marginalise_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(hist, g, a, b, mode);
    };
    for (int g = nmb::kBipMinGap; g <= nmb::kBipMaxGap; ++g) {
        run(g, 4, 3, 0);  // (4,g,3) from (4,g,4): last letter
        run(g, 3, 4, 1);  // (3,g,4) from (4,g,4): first letter
        run(g, 3, 3, 0);  // (3,g,3) from the now complete (3,g,4): last letter
        NMB_CUDA(cudaGetLastError());
    }
    return NMB_OK;
}

int nmb_sweep_slice(const uint32_t *src, uint32_t *dst, int64_t n_out, int64_t axis_stride, int32_t digit,
                    void *stream) {
    NMB_REQUIRE(src && dst && n_out > 0 && axis_stride > 0 && digit >= 0 && digit < 5, "nmb_sweep_slice: bad argument");
    nmb::sweep_slice_kernel<<<(unsigned)((n_out + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, dst, n_out,
                                                                                             axis_stride, digit);
    NMB_CUDA(cudaGetLastError());
    return NMB_OK;
}

int nmb_sweep_expand(const uint32_t *src, uint32_t *dst, int64_t outer, int64_t inner, void *stream) {
    NMB_REQUIRE(src && dst && outer > 0 && inner > 0, "nmb_sweep_expand: bad argument");
    int64_t blocks = (outer * inner + 255) / 256;
    if (blocks > 148 * 64) blocks = 148 * 64;
    nmb::sweep_expand_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, dst, outer, inner);
    NMB_CUDA(cudaGetLastError());
    return NMB_OK;
}

int nmb_sweep_filter(const uint32_t *n_mod, const uint32_t *n_nomod, int64_t n, double min_mean, int64_t min_mod,
                     int64_t *out_index, int64_t capacity, int64_t *n_out, void *stream) {
    NMB_REQUIRE(n_mod && n_nomod && n_out && n >= 0 && capacity >= 0 && min_mod >= 0, "nmb_sweep_filter: bad argument");
    NMB_REQUIRE(capacity == 0 || out_index, "nmb_sweep_filter: null output");
    NMB_CUDA(cudaMemsetAsync(n_out, 0, sizeof(int64_t), (cudaStream_t)stream));
    if (n == 0) return NMB_OK;
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    nmb::sweep_filter_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        n_mod, n_nomod, n, min_mean, (uint32_t)(min_mod > 0xFFFFFFFFll ? 0xFFFFFFFFll : min_mod), out_index, capacity,
        (unsigned long long *)n_out);
    NMB_CUDA(cudaGetLastError());
    return NMB_OK;
}

}  // extern "C"
