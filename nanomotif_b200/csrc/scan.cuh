// Bit-parallel motif matcher shared by K2 (scan_count) and K3 (match_plane).
//
// Each lane owns one 256-bp chunk = NW (8) consecutive 32-bit words of the two sequence bit-planes
// x (high code bit) and y (low code bit), plus H halo words on both sides.  For a motif with
// constrained positions j_0 < ... < j_{K-1} and modified-base position mp, the match plane aligned
// at the modified base is
//     M[p] = AND_i  ind_{j_i}( base[p + j_i - mp] )
// evaluated as two Horner chains so that every constrained position costs one funnel shift of the
// running plane plus ONE lop3 per word:
//     left  chain (j_i <= mp, ascending):  L <- ind_j(x, y) & (L << (j - j_prev))
//     right chain (j_i >  mp, descending): R <- ind_j(x, y) & (R >> (j_prev - j))
//     M = (L << (mp - j_last_left)) & (R >> (j_first_right - mp))
// ind_j is one of 14 two-variable boolean functions, so the allowed-set is folded into the lop3
// immediate; the switch on the set code is uniform across the CTA.  Non-ACGT letters (and the
// padding between contigs) must fail every constrained position: warps that touch a flagged chunk
// take the HASN variant, which spends one more lop3 per word.
#pragma once
#include "common.cuh"

namespace nmb {

constexpr int NW = kChunkWords;  // words per lane

template <int H>
struct LaneSeq {
    static constexpr int XW = NW + 2 * H;  // words held per plane: [-H, NW + H)
    uint32_t x[XW], y[XW], n[XW];
};

// Entry of a Program: allowed-set code | motif position << 8.
struct ProgramView {
    const uint16_t *ent;
    int n, n_left, mod_pos;
};

__device__ __forceinline__ ProgramView load_program(const Program *p) {
    const uint32_t hdr = __ldg(reinterpret_cast<const uint32_t *>(p));
    ProgramView v;
    v.n = hdr & 0xFF;
    v.n_left = (hdr >> 8) & 0xFF;
    v.mod_pos = (hdr >> 16) & 0xFF;
    v.ent = reinterpret_cast<const uint16_t *>(reinterpret_cast<const uint8_t *>(p) + 4);
    return v;
}

template <int M, bool HASN, bool LEFT, int H>
__device__ __forceinline__ void chain_step(uint32_t (&c)[NW + H], const LaneSeq<H> &q, int s) {
    constexpr int CW = NW + H;
    constexpr int T = set_truth(M);
    if (!LEFT) {  // word i of the chain is lane word i; planes are indexed from -H
#pragma unroll
        for (int i = 0; i < CW; ++i) {
            const uint32_t hi = (i + 1 < CW) ? c[i + 1] : 0u;
            const uint32_t sh = __funnelshift_r(c[i], hi, s);
            if (HASN)
                c[i] = lop3<(T & 0x55)>(q.x[i + H], q.y[i + H], q.n[i + H]) & sh;
            else
                c[i] = lop3<(T & 0xAA)>(q.x[i + H], q.y[i + H], sh);
        }
    } else {  // word i of the chain is lane word i - H
#pragma unroll
        for (int i = CW - 1; i >= 0; --i) {
            const uint32_t lo = (i > 0) ? c[i - 1] : 0u;
            const uint32_t sh = __funnelshift_l(lo, c[i], s);
            if (HASN)
                c[i] = lop3<(T & 0x55)>(q.x[i], q.y[i], q.n[i]) & sh;
            else
                c[i] = lop3<(T & 0xAA)>(q.x[i], q.y[i], sh);
        }
    }
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

template <bool LEFT, int H>
__device__ __forceinline__ void chain_shift_bits(uint32_t (&c)[NW + H], int s) {  // 0 < s < 32
    constexpr int CW = NW + H;
    if (!LEFT) {
#pragma unroll
        for (int i = 0; i < CW; ++i) c[i] = __funnelshift_r(c[i], (i + 1 < CW) ? c[i + 1] : 0u, s);
    } else {
#pragma unroll
        for (int i = CW - 1; i >= 0; --i) c[i] = __funnelshift_l((i > 0) ? c[i - 1] : 0u, c[i], s);
    }
}

#define NMB_SET_CASES(X) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14)

template <bool HASN, bool LEFT, int H>
__device__ __forceinline__ void chain_apply(uint32_t (&c)[NW + H], const LaneSeq<H> &q, int code,
                                            int s) {
    while (s >= 32) {
        chain_shift_words<LEFT, H>(c);
        s -= 32;
    }
    switch (code) {
#define NMB_CASE(m)                           \
    case m:                                   \
        chain_step<m, HASN, LEFT, H>(c, q, s); \
        break;
        NMB_SET_CASES(NMB_CASE)
#undef NMB_CASE
        case 15:  // wildcard entries are never compiled into a program; keep it a pure shift
            if (s) chain_shift_bits<LEFT, H>(c, s);
            break;
        default:  // empty set: nothing matches
#pragma unroll
            for (int i = 0; i < NW + H; ++i) c[i] = 0u;
            break;
    }
}

// Match words of the lane's NW words, aligned at the program's mod_pos.
template <bool HASN, int H>
__device__ __forceinline__ void match_words(const ProgramView &pv, const LaneSeq<H> &q,
                                            uint32_t (&m)[NW]) {
    constexpr int CW = NW + H;
    uint32_t L[CW], R[CW];
#pragma unroll
    for (int i = 0; i < CW; ++i) {
        L[i] = 0xFFFFFFFFu;
        R[i] = 0xFFFFFFFFu;
    }
    // left chain: entries [0, n_left) ascending
    int prev = 0;
    if (pv.n_left > 0) {
        uint32_t e = __ldg(pv.ent);
        prev = e >> 8;
        int s = 0;
#pragma unroll 1
        for (int i = 0; i < pv.n_left; ++i) {
            const uint32_t cur = e;
            if (i + 1 < pv.n_left) e = __ldg(pv.ent + i + 1);  // prefetch next entry
            chain_apply<HASN, true, H>(L, q, cur & 0xFF, s);
            s = (int)(e >> 8) - (int)(cur >> 8);
            prev = cur >> 8;
        }
        int sl = pv.mod_pos - prev;  // modified base itself is a wildcard: pure shift
        while (sl >= 32) {
            chain_shift_words<true, H>(L);
            sl -= 32;
        }
        if (sl) chain_shift_bits<true, H>(L, sl);
    }
    // right chain: entries (n_left, n] descending
    int sr = 0;
    if (pv.n > pv.n_left) {
        uint32_t e = __ldg(pv.ent + pv.n - 1);
        int s = 0;
#pragma unroll 1
        for (int i = pv.n - 1; i >= pv.n_left; --i) {
            const uint32_t cur = e;
            if (i - 1 >= pv.n_left) e = __ldg(pv.ent + i - 1);
            chain_apply<HASN, false, H>(R, q, cur & 0xFF, s);
            s = (int)(cur >> 8) - (int)(e >> 8);
            prev = cur >> 8;
        }
        sr = prev - pv.mod_pos;
        while (sr >= 32) {
            chain_shift_words<false, H>(R);
            sr -= 32;
        }
    }
#pragma unroll
    for (int k = 0; k < NW; ++k)
        m[k] = L[k + H] & __funnelshift_r(R[k], R[k + 1], sr);
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

}  // namespace nmb
