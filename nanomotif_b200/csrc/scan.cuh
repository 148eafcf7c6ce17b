// Bit-parallel motif matcher shared by K2 (scan_count) and K3 (match_plane).
//
// Each lane owns one 256-bp chunk = NW (8) consecutive 32-bit words of the two sequence bit-planes
// x (high code bit) and y (low code bit), plus H halo words on both sides.  For a motif with
// constrained positions j_0 < ... < j_{K-1} and modified-base position mp, the match plane aligned
// at the modified base is
//     M[p] = AND_i  ind_{j_i}( base[p + j_i - mp] )
// evaluated as two Horner chains so that every constrained position costs one funnel shift of the
// running plane plus ONE logic op per word:
//     left  chain (j_i <= mp, ascending):  L <- ind_j & (L << (j - j_prev))
//     right chain (j_i >  mp, descending): R <- ind_j & (R >> (j_prev - j))
//     M = (L << (mp - j_last_left)) & (R >> (j_first_right - mp))
// Two register representations of the sequence:
//   XY     the raw planes; a single-base indicator is folded into the lop3 immediate
//          (R <- lop3<base>(x, y, shifted R)).  Used by warps whose chunks are pure ACGT.
//   PLANES four one-hot planes A,T,G,C with non-ACGT letters (and inter-contig padding) cleared, so
//          such letters fail every constrained position and only match '.', which is the regex
//          semantics of the reference.  Used by flagged warps.
// The branch on the base is uniform across the CTA (it depends on the motif only).
#pragma once
#include "common.cuh"

namespace nmb {

constexpr int NW = kChunkWords;  // words per lane

template <int H, bool PLANES>
struct LaneSeq;

template <int H>
struct LaneSeq<H, false> {
    static constexpr int XW = NW + 2 * H;  // words held per plane: [-H, NW + H)
    uint32_t x[XW], y[XW];
    template <int C>  // indicator of single base C (0=A 1=T 2=G 3=C) at plane word i, ANDed with s
    __device__ __forceinline__ uint32_t and_base(int i, uint32_t s) const {
        constexpr int T = (C == 0 ? 0x03 : C == 1 ? 0x0C : C == 2 ? 0x30 : 0xC0) & 0xAA;
        return lop3<T>(x[i], y[i], s);
    }
    // indicator of an arbitrary set given as four uniform masks (0 / ~0), ANDed with s
    __device__ __forceinline__ uint32_t and_set(int i, uint32_t mA, uint32_t mT, uint32_t mG,
                                                uint32_t mC, uint32_t s) const {
        const uint32_t g0 = (y[i] & mT) | (~y[i] & mA);  // x = 0: A or T
        const uint32_t g1 = (y[i] & mC) | (~y[i] & mG);  // x = 1: G or C
        return ((x[i] & g1) | (~x[i] & g0)) & s;
    }
};

template <int H>
struct LaneSeq<H, true> {
    static constexpr int XW = NW + 2 * H;
    uint32_t p[4][XW];  // one-hot planes, non-ACGT cleared
    template <int C>
    __device__ __forceinline__ uint32_t and_base(int i, uint32_t s) const {
        return p[C][i] & s;
    }
    __device__ __forceinline__ uint32_t and_set(int i, uint32_t mA, uint32_t mT, uint32_t mG,
                                                uint32_t mC, uint32_t s) const {
        return ((p[0][i] & mA) | (p[1][i] & mT) | (p[2][i] & mG) | (p[3][i] & mC)) & s;
    }
};

// Program header and entries (see common.cuh::Program): entries are stored in processing order,
// left chain first; entry = set code | (shift from the previously processed entry) << 8.
struct ProgramView {
    const uint16_t *ent;
    int n_left, n_right, sl, sr;
};

__device__ __forceinline__ ProgramView load_program(const Program *p) {
    const uint32_t hdr = __ldg(reinterpret_cast<const uint32_t *>(p));
    ProgramView v;
    v.n_left = hdr & 0xFF;
    v.n_right = (hdr >> 8) & 0xFF;
    v.sl = (hdr >> 16) & 0xFF;
    v.sr = (hdr >> 24) & 0xFF;
    v.ent = p->ent;
    return v;
}

template <bool LEFT, int H>
__device__ __forceinline__ void chain_shift_words(uint32_t (&c)[NW + H]) {  // shift by 32
    constexpr int CW = NW + H;
    if (!LEFT) {
#pragma unroll
        for (int i = 0; i < CW; ++i) c[i] = (i + 1 < CW) ? c[i + 1] : 0u;
    } else {
#pragma unroll
        for (int i = CW - 1; i >= 0; --i) c[i] = (i > 0) ? c[i - 1] : 0u;
    }
}

// word i of the right chain is lane word i (plane index i + H); word i of the left chain is lane
// word i - H (plane index i).
template <bool LEFT, int H>
__device__ __forceinline__ uint32_t shifted(const uint32_t (&c)[NW + H], int i, int s) {
    constexpr int CW = NW + H;
    if (!LEFT) return __funnelshift_r(c[i], (i + 1 < CW) ? c[i + 1] : 0u, s);
    return __funnelshift_l((i > 0) ? c[i - 1] : 0u, c[i], s);
}

template <bool LEFT, int H>
__device__ __forceinline__ void chain_shift_bits(uint32_t (&c)[NW + H], int s) {  // 0 < s < 32
    constexpr int CW = NW + H;
    if (!LEFT) {
#pragma unroll
        for (int i = 0; i < CW; ++i) c[i] = shifted<LEFT, H>(c, i, s);
    } else {
#pragma unroll
        for (int i = CW - 1; i >= 0; --i) c[i] = shifted<LEFT, H>(c, i, s);
    }
}

template <int C, bool LEFT, int H, bool PLANES>
__device__ __forceinline__ void step_base(uint32_t (&c)[NW + H], const LaneSeq<H, PLANES> &q, int s) {
    constexpr int CW = NW + H;
    if (!LEFT) {
#pragma unroll
        for (int i = 0; i < CW; ++i) c[i] = q.template and_base<C>(i + H, shifted<LEFT, H>(c, i, s));
    } else {
#pragma unroll
        for (int i = CW - 1; i >= 0; --i) c[i] = q.template and_base<C>(i, shifted<LEFT, H>(c, i, s));
    }
}

template <bool LEFT, int H, bool PLANES>
__device__ __forceinline__ void step_set(uint32_t (&c)[NW + H], const LaneSeq<H, PLANES> &q, int code,
                                         int s) {
    constexpr int CW = NW + H;
    const uint32_t mA = (code & 1) ? 0xFFFFFFFFu : 0u, mT = (code & 2) ? 0xFFFFFFFFu : 0u;
    const uint32_t mG = (code & 4) ? 0xFFFFFFFFu : 0u, mC = (code & 8) ? 0xFFFFFFFFu : 0u;
    if (!LEFT) {
#pragma unroll
        for (int i = 0; i < CW; ++i) c[i] = q.and_set(i + H, mA, mT, mG, mC, shifted<LEFT, H>(c, i, s));
    } else {
#pragma unroll
        for (int i = CW - 1; i >= 0; --i) c[i] = q.and_set(i, mA, mT, mG, mC, shifted<LEFT, H>(c, i, s));
    }
}

// First entry of a chain: the running plane is the indicator itself (no shift, no AND).
template <int C, bool LEFT, int H, bool PLANES>
__device__ __forceinline__ void init_base(uint32_t (&c)[NW + H], const LaneSeq<H, PLANES> &q) {
#pragma unroll
    for (int i = 0; i < NW + H; ++i) c[i] = q.template and_base<C>(LEFT ? i : i + H, 0xFFFFFFFFu);
}
template <bool LEFT, int H, bool PLANES>
__device__ __forceinline__ void init_set(uint32_t (&c)[NW + H], const LaneSeq<H, PLANES> &q, int code) {
    const uint32_t mA = (code & 1) ? 0xFFFFFFFFu : 0u, mT = (code & 2) ? 0xFFFFFFFFu : 0u;
    const uint32_t mG = (code & 4) ? 0xFFFFFFFFu : 0u, mC = (code & 8) ? 0xFFFFFFFFu : 0u;
#pragma unroll
    for (int i = 0; i < NW + H; ++i) c[i] = q.and_set(LEFT ? i : i + H, mA, mT, mG, mC, 0xFFFFFFFFu);
}

// Run the n >= 1 program entries at ent[0..n) over chain c.  Entry codes: 1/2/4/8 single base,
// kEntShift32 = "shift the chain by one word" (emitted by the compiler for gaps >= 32), anything
// else a degenerate set (0 = empty set).  All shifts are < 32 and the first entry's shift is 0.
constexpr int kEntShift32 = 0x10;

template <bool LEFT, int H, bool PLANES>
__device__ __forceinline__ void run_chain(uint32_t (&c)[NW + H], const LaneSeq<H, PLANES> &q,
                                          const uint16_t *ent, int n) {
    uint32_t e = __ldg(ent);
    {
        const int code = e & 0xFF;
        if (n > 1) e = __ldg(ent + 1);
        if (code == 1) init_base<0, LEFT, H, PLANES>(c, q);
        else if (code == 2) init_base<1, LEFT, H, PLANES>(c, q);
        else if (code == 4) init_base<2, LEFT, H, PLANES>(c, q);
        else if (code == 8) init_base<3, LEFT, H, PLANES>(c, q);
        else init_set<LEFT, H, PLANES>(c, q, code);
    }
#pragma unroll 1
    for (int i = 1; i < n; ++i) {
        const uint32_t cur = e;
        if (i + 1 < n) e = __ldg(ent + i + 1);  // prefetch the next entry
        const int s = cur >> 8;
        const int code = cur & 0xFF;
        if (code == 1) step_base<0, LEFT, H, PLANES>(c, q, s);
        else if (code == 2) step_base<1, LEFT, H, PLANES>(c, q, s);
        else if (code == 4) step_base<2, LEFT, H, PLANES>(c, q, s);
        else if (code == 8) step_base<3, LEFT, H, PLANES>(c, q, s);
        else if (code == kEntShift32) chain_shift_words<LEFT, H>(c);
        else step_set<LEFT, H, PLANES>(c, q, code, s);
    }
}

// Evaluate one program on the lane's words.  On return the match plane aligned at mod_pos is
//     M[k] = L[k + H] & funnel_r(R[k], R[k + 1], sr)        for k in [0, NW)
// (left chain empty -> L = ~0; right chain empty -> R = ~0, sr = 0).
template <int H, bool PLANES>
__device__ __forceinline__ int eval_program(const ProgramView &pv, const LaneSeq<H, PLANES> &q,
                                            uint32_t (&L)[NW + H], uint32_t (&R)[NW + H]) {
    constexpr int CW = NW + H;
    if (pv.n_left > 0) {
        run_chain<true, H, PLANES>(L, q, pv.ent, pv.n_left);
        if (pv.sl) {  // only when the modified base itself is a wildcard (sl < 32 by construction)
            chain_shift_bits<true, H>(L, pv.sl);
        }
    } else {
#pragma unroll
        for (int i = 0; i < CW; ++i) L[i] = 0xFFFFFFFFu;
    }
    if (pv.n_right > 0) {
        run_chain<false, H, PLANES>(R, q, pv.ent + pv.n_left, pv.n_right);
        return pv.sr;
    }
#pragma unroll
    for (int i = 0; i < CW; ++i) R[i] = 0xFFFFFFFFu;
    return 0;
}

template <int H, bool PLANES>
__device__ __forceinline__ void match_words(const ProgramView &pv, const LaneSeq<H, PLANES> &q,
                                            uint32_t (&m)[NW]) {
    uint32_t L[NW + H], R[NW + H];
    const int sr = eval_program<H, PLANES>(pv, q, L, R);
#pragma unroll
    for (int k = 0; k < NW; ++k) m[k] = L[k + H] & __funnelshift_r(R[k], R[k + 1], sr);
}

// Load the lane's words [-H, NW+H) of one plane; `base` points at lane word 0 and is 16-byte
// aligned (shared or global memory).
template <int H>
__device__ __forceinline__ void load_plane(const uint32_t *base, uint32_t (&w)[NW + 2 * H]) {
    const uint4 a = *reinterpret_cast<const uint4 *>(base);
    const uint4 b = *reinterpret_cast<const uint4 *>(base + 4);
    w[H + 0] = a.x; w[H + 1] = a.y; w[H + 2] = a.z; w[H + 3] = a.w;
    w[H + 4] = b.x; w[H + 5] = b.y; w[H + 6] = b.z; w[H + 7] = b.w;
#pragma unroll
    for (int i = 0; i < H; ++i) {
        w[i] = base[i - H];
        w[NW + H + i] = base[NW + i];
    }
}

// Build both representations of the lane's sequence words.  sx / sy point at lane word 0 of the x
// and y planes (shared or global); gn at word -H of the non-ACGT plane (global).
template <int H>
__device__ __forceinline__ void load_xy(const uint32_t *sx, const uint32_t *sy, LaneSeq<H, false> &q) {
    load_plane<H>(sx, q.x);
    load_plane<H>(sy, q.y);
}
template <int H>
__device__ __forceinline__ void load_planes(const uint32_t *sx, const uint32_t *sy, const uint32_t *gn,
                                            LaneSeq<H, true> &q) {
    uint32_t x[NW + 2 * H], y[NW + 2 * H];
    load_plane<H>(sx, x);
    load_plane<H>(sy, y);
#pragma unroll
    for (int i = 0; i < NW + 2 * H; ++i) {
        const uint32_t ok = ~__ldg(gn + i);
        q.p[0][i] = ~x[i] & ~y[i] & ok;
        q.p[1][i] = ~x[i] & y[i] & ok;
        q.p[2][i] = x[i] & ~y[i] & ok;
        q.p[3][i] = x[i] & y[i] & ok;
    }
}

}  // namespace nmb
