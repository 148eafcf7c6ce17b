// Bit-parallel motif matcher shared by K2 (scan_count) and K3 (match_plane).
//
// Each lane owns one 512-bp chunk = NW (16) consecutive 32-bit words of the two sequence bit-planes
// x (high code bit) and y (low code bit), plus H halo words on both sides.  For a motif with
// constrained positions j_0 < ... < j_{K-1} the start-aligned match plane
//     S[p] = AND_i  ind_{j_i}( base[p + j_i] )
// is evaluated as ONE Horner chain from the last constrained position to the first, so that every
// constrained position costs one funnel shift of the running plane plus one lop3 per word:
//     C <- ind_{j_K-1};   C <- ind_{j_i} & (C >> (j_{i+1} - j_i))   for i = K-2 .. 0
// and the plane aligned at the modified base is M = C << mod_pos (one more funnel shift).  The chain
// runs over words [-H, NW + H): the right halo feeds the >> shifts, the left halo the final <<.
// Two register representations of the sequence:
//   XY   the raw planes; a single-base indicator is folded into the lop3 immediate
//        (C <- lop3<base>(x, y, shifted C)).  Used by warps whose chunks are pure ACGT.
//   XYN  additionally the non-ACGT plane n, ANDed out of every indicator so that such letters (and the
//        inter-contig padding) fail every constrained position and only match '.', the regex
//        semantics of the reference.  One extra logic op per word and
//        step; used only by warps that touch a flagged chunk.
// The branch on the base is uniform across the CTA (it depends on the motif only).
#pragma once
#include "common.cuh"

namespace nmb {

constexpr int NW = kChunkWords;  // words per lane

template <int H, bool HASN>
struct LaneSeq;

// lop3 immediates over inputs (a, b, c) = (0xF0, 0xCC, 0xAA)
constexpr int kLutMux = 0xCA;       // a ? b : c
constexpr int kLutAndNot = 0x30;    // a & ~b
constexpr int kLutAnd = 0xC0;       // a & b

// Every indicator below is written so that it depends on the running (shifted) plane s: otherwise the
// compiler hoists the step-invariant part (e.g. ~y, or ind(x, y) & ~n) out of the motif loop and keeps
// one extra register per word and plane alive.
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
        const uint32_t g0 = lop3<kLutMux>(y[i], mT, mA);  // x = 0: T or A  (masks change every step,
        const uint32_t g1 = lop3<kLutMux>(y[i], mC, mG);  // x = 1: C or G   so nothing is hoistable)
        return lop3<kLutMux>(x[i], g1, g0) & s;
    }
};

template <int H>
struct LaneSeq<H, true> {
    static constexpr int XW = NW + 2 * H;
    uint32_t x[XW], y[XW], n[XW];  // n = non-ACGT plane (inter-contig padding included)
    template <int C>
    __device__ __forceinline__ uint32_t and_base(int i, uint32_t s) const {
        constexpr int T = (C == 0 ? 0x03 : C == 1 ? 0x0C : C == 2 ? 0x30 : 0xC0) & 0xAA;
        return lop3<T>(x[i], y[i], lop3<kLutAndNot>(s, n[i], 0u));
    }
    __device__ __forceinline__ uint32_t and_set(int i, uint32_t mA, uint32_t mT, uint32_t mG,
                                                uint32_t mC, uint32_t s) const {
        const uint32_t g0 = lop3<kLutMux>(y[i], mT, mA);
        const uint32_t g1 = lop3<kLutMux>(y[i], mC, mG);
        return lop3<kLutMux>(x[i], g1, g0) & lop3<kLutAndNot>(s, n[i], 0u);
    }
};

// Program (common.cuh): header n | mod_pos << 8 | len << 16, then n entries in processing order
// (last constrained position first).  entry = set code | (distance to the previous entry) << 8.
struct ProgramView {
    const uint16_t *ent;
    int n, mod_pos;
};

__device__ __forceinline__ ProgramView load_program(const Program *p) {
    const uint32_t hdr = __ldg(reinterpret_cast<const uint32_t *>(p));
    ProgramView v;
    v.n = hdr & 0xFF;
    v.mod_pos = (hdr >> 8) & 0xFF;
    v.ent = p->ent;
    return v;
}

template <int H>
__device__ __forceinline__ uint32_t shifted(const uint32_t (&c)[NW + 2 * H], int i, int s) {
    return __funnelshift_r(c[i], (i + 1 < NW + 2 * H) ? c[i + 1] : 0u, s);
}

template <int C, int H, bool HASN>
__device__ __forceinline__ void step_base(uint32_t (&c)[NW + 2 * H], const LaneSeq<H, HASN> &q, int s) {
#pragma unroll
    for (int i = 0; i < NW + 2 * H; ++i) c[i] = q.template and_base<C>(i, shifted<H>(c, i, s));
}

template <int H, bool HASN>
__device__ __forceinline__ void step_set(uint32_t (&c)[NW + 2 * H], const LaneSeq<H, HASN> &q, int code,
                                         int s) {
    const uint32_t mA = (code & 1) ? 0xFFFFFFFFu : 0u, mT = (code & 2) ? 0xFFFFFFFFu : 0u;
    const uint32_t mG = (code & 4) ? 0xFFFFFFFFu : 0u, mC = (code & 8) ? 0xFFFFFFFFu : 0u;
#pragma unroll
    for (int i = 0; i < NW + 2 * H; ++i) c[i] = q.and_set(i, mA, mT, mG, mC, shifted<H>(c, i, s));
}

// First entry of the chain: the running plane is the indicator itself (no shift, no AND).
template <int C, int H, bool HASN>
__device__ __forceinline__ void init_base(uint32_t (&c)[NW + 2 * H], const LaneSeq<H, HASN> &q) {
#pragma unroll
    for (int i = 0; i < NW + 2 * H; ++i) c[i] = q.template and_base<C>(i, 0xFFFFFFFFu);
}
template <int H, bool HASN>
__device__ __forceinline__ void init_set(uint32_t (&c)[NW + 2 * H], const LaneSeq<H, HASN> &q, int code) {
    const uint32_t mA = (code & 1) ? 0xFFFFFFFFu : 0u, mT = (code & 2) ? 0xFFFFFFFFu : 0u;
    const uint32_t mG = (code & 4) ? 0xFFFFFFFFu : 0u, mC = (code & 8) ? 0xFFFFFFFFu : 0u;
#pragma unroll
    for (int i = 0; i < NW + 2 * H; ++i) c[i] = q.and_set(i, mA, mT, mG, mC, 0xFFFFFFFFu);
}

// Entry codes: 1/2/4/8 single base, kEntShift32 = "shift the chain by one word" (emitted by the motif
// compiler for gaps >= 32), anything else a degenerate set (0 = empty set).  All shifts are < 32 and
// the first entry's shift is 0.
constexpr int kEntShift32 = 0x10;

// Start-aligned match words of the lane's words [-H, NW + H); valid for words [-H, NW).
template <int H, bool HASN>
__device__ __forceinline__ void run_chain(const ProgramView &pv, const LaneSeq<H, HASN> &q,
                                          uint32_t (&c)[NW + 2 * H]) {
    uint32_t e = __ldg(pv.ent);
    {
        const int code = e & 0xFF;
        if (pv.n > 1) e = __ldg(pv.ent + 1);
        if (code == 1) init_base<0, H, HASN>(c, q);
        else if (code == 2) init_base<1, H, HASN>(c, q);
        else if (code == 4) init_base<2, H, HASN>(c, q);
        else if (code == 8) init_base<3, H, HASN>(c, q);
        else init_set<H, HASN>(c, q, code);
    }
#pragma unroll 1
    for (int i = 1; i < pv.n; ++i) {
        const uint32_t cur = e;
        if (i + 1 < pv.n) e = __ldg(pv.ent + i + 1);  // prefetch the next entry
        const int s = cur >> 8;
        const int code = cur & 0xFF;
        if (code == 1) step_base<0, H, HASN>(c, q, s);
        else if (code == 2) step_base<1, H, HASN>(c, q, s);
        else if (code == 4) step_base<2, H, HASN>(c, q, s);
        else if (code == 8) step_base<3, H, HASN>(c, q, s);
        else if (code == kEntShift32) {
#pragma unroll
            for (int k = 0; k < NW + 2 * H; ++k) c[k] = (k + 1 < NW + 2 * H) ? c[k + 1] : 0u;
        } else step_set<H, HASN>(c, q, code, s);
    }
}

// Word k (0 <= k < NW) of the match plane aligned at mod_pos: M[p] = S[p - mod_pos].
template <int H>
__device__ __forceinline__ uint32_t aligned_word(const uint32_t (&c)[NW + 2 * H], int k, int sh, bool far) {
    if (H == 1 || !far) return __funnelshift_l(c[k + H - 1], c[k + H], sh);
    return __funnelshift_l(c[k + H - 2], c[k + H - 1], sh);
}

template <int H, bool HASN>
__device__ __forceinline__ void match_words(const ProgramView &pv, const LaneSeq<H, HASN> &q,
                                            uint32_t (&m)[NW]) {
    uint32_t c[NW + 2 * H];
    run_chain<H, HASN>(pv, q, c);
    const bool far = pv.mod_pos >= 32;  // only possible when H == 2
    const int sh = pv.mod_pos & 31;
#pragma unroll
    for (int k = 0; k < NW; ++k) m[k] = aligned_word<H>(c, k, sh, far);
}

// Load the lane's words [-H, NW+H) of one plane; `base` points at lane word 0 and is 16-byte
// aligned (shared or global memory).
template <int H>
__device__ __forceinline__ void load_plane(const uint32_t *base, uint32_t (&w)[NW + 2 * H]) {
#pragma unroll
    for (int v = 0; v < NW / 4; ++v) {
        const uint4 a = *reinterpret_cast<const uint4 *>(base + 4 * v);
        w[H + 4 * v + 0] = a.x; w[H + 4 * v + 1] = a.y; w[H + 4 * v + 2] = a.z; w[H + 4 * v + 3] = a.w;
    }
#pragma unroll
    for (int i = 0; i < H; ++i) {
        w[i] = base[i - H];
        w[NW + H + i] = base[NW + i];
    }
}

// sx / sy point at lane word 0 of the x and y planes (shared or global); gn at word -H of the
// non-ACGT plane (global).
template <int H>
__device__ __forceinline__ void load_xy(const uint32_t *sx, const uint32_t *sy, LaneSeq<H, false> &q) {
    load_plane<H>(sx, q.x);
    load_plane<H>(sy, q.y);
}
template <int H>
__device__ __forceinline__ void load_xyn(const uint32_t *sx, const uint32_t *sy, const uint32_t *gn,
                                         LaneSeq<H, true> &q) {
    load_plane<H>(sx, q.x);
    load_plane<H>(sy, q.y);
#pragma unroll
    for (int i = 0; i < NW + 2 * H; ++i) q.n[i] = __ldg(gn + i);
}

}  // namespace nmb
