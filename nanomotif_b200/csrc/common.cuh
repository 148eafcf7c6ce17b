// Shared host/device helpers for libnmb200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "nmb200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libnmb200 is written for sm_100a (B200) only"
#endif

namespace nmb {

// ---------------------------------------------------------------------------------------------
// error plumbing (never throw across the C ABI)
// ---------------------------------------------------------------------------------------------
char *last_error_buffer();  // thread-local, 512 bytes (api.cu)

#define NMB_FAIL(code, ...)                                        \
    do {                                                           \
        snprintf(nmb::last_error_buffer(), 512, __VA_ARGS__);      \
        return (code);                                             \
    } while (0)

#define NMB_CUDA(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess)                                                                  \
            NMB_FAIL(NMB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),      \
                     __FILE__, __LINE__);                                                       \
    } while (0)

#define NMB_REQUIRE(cond, ...)                          \
    do {                                                \
        if (!(cond)) NMB_FAIL(NMB_ERR_INVALID, __VA_ARGS__); \
    } while (0)

// ---------------------------------------------------------------------------------------------
// layout constants (mirrors of nmb200.h)
// ---------------------------------------------------------------------------------------------
constexpr int kChunkWords = NMB_CHUNK_WORDS;       // words per lane chunk (512 bp)
constexpr int kTileWords = NMB_TILE_WORDS;         // words per tile (65536 bp)
constexpr int kTileChunks = NMB_TILE_CHUNKS;       // 128 chunks = 128 threads
constexpr int kHalo = NMB_HALO_WORDS;              // duplicated halo words per side
constexpr int kSeqPlaneWords = NMB_SEQ_PLANE_WORDS;
constexpr int kSeqRecWords = NMB_SEQ_REC_WORDS;
constexpr int kClsRecWords = NMB_CLS_REC_WORDS;
constexpr int kSeqRecBytes = kSeqRecWords * 4;     // 16960
constexpr int kClsRecBytes = kClsRecWords * 4;     // 32768
constexpr int kMaxLen = NMB_MAX_MOTIF_LEN;
// chunk_info flags (nmb200.h): the lane's words [-2, 18) ...
constexpr int kChunkFlagN = 1 << 30;      // ... may hold a non-ACGT LETTER of a contig: every constrained position must test it
constexpr int kChunkFlagEdge = 1 << 29;   // ... hold inter-contig padding only: first and last motif position must avoid it
constexpr int kChunkFlagLetter = 1 << 28; // the chunk itself holds a non-ACGT letter (written by the packer, input of the two above)
constexpr int kChunkIdMask = (1 << 28) - 1;

static_assert(kTileWords == kTileChunks * kChunkWords, "tile = 128 lane chunks");
static_assert(kChunkWords == 16 && kTileChunks == 128, "NMB_WORD_SLOT assumes 128 chunks of 16 words");
constexpr int kSlotStride = kTileChunks * 4;       // words between the uint4 vectors of one lane (512)
__host__ __device__ constexpr int word_slot(int w) { return NMB_WORD_SLOT(w); }
static_assert(kChunkWords % 4 == 0 && kChunkWords <= 32, "lane chunk is a whole number of uint4");
static_assert(kSeqRecBytes % 16 == 0 && kClsRecBytes % 16 == 0, "bulk copies need 16-byte sizes");

// Compiled scan program of one motif strand (device representation, 128 bytes).  Only constrained
// positions appear, in processing order (LAST position first).  entry = allowed-set code (1..14;
// 0 = never matches) | (distance to the previously processed position) << 8; gaps >= 32 are split
// off as "shift one word" pseudo entries (code 0x10), so every shift is < 32.
struct Program {
    uint8_t n;        // entries (pseudo entries included)
    uint8_t mod_pos;  // modified-base position within the stripped motif
    uint8_t len;
    uint8_t reserved;
    uint16_t ent[kMaxLen];
};
static_assert(sizeof(Program) == 128, "Program must be 128 bytes");
constexpr int kProgramBytesPerMotif = 2 * sizeof(Program);  // forward + reverse complement

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__

template <int LUT>
__device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(d) : "r"(a), "r"(b), "r"(c), "n"(LUT));
    return d;
}

// Truth table of "base (x=high bit, y=low bit) is in set m" over lop3 inputs a=x, b=y.
// A=(0,0) T=(0,1) G=(1,0) C=(1,1); m bit0=A bit1=T bit2=G bit3=C.
__host__ __device__ constexpr int set_truth(int m) {
    return ((m & 1) ? 0x03 : 0) | ((m & 2) ? 0x0C : 0) | ((m & 4) ? 0x30 : 0) | ((m & 8) ? 0xC0 : 0);
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// 1-D bulk async copy global -> shared (TMA engine, SASS UBLKCP), completion on an mbarrier.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes,
                                         uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ---- small building blocks shared by the compaction-style kernels (positions.cu, bedmethyl.cu) ----

// Exclusive scan of n int64 in place by one block; total -> *n_out.
template <int kThreads>
__global__ void __launch_bounds__(kThreads) scan_counts_kernel(int64_t *__restrict__ v, int64_t n,
                                                           int64_t *__restrict__ n_out) {
    __shared__ int64_t s_part[kThreads];
    const int t = threadIdx.x;
    const int64_t per = (n + kThreads - 1) / kThreads;
    const int64_t b = t * per, e = min(n, b + per);
    int64_t sum = 0;
    for (int64_t i = b; i < e; ++i) sum += v[i];
    s_part[t] = sum;
    __syncthreads();
    if (t == 0) {
        int64_t run = 0;
        for (int i = 0; i < kThreads; ++i) {
            const int64_t x = s_part[i];
            s_part[i] = run;
            run += x;
        }
        *n_out = run;
    }
    __syncthreads();
    int64_t run = s_part[t];
    for (int64_t i = b; i < e; ++i) {
        const int64_t x = v[i];
        v[i] = run;
        run += x;
    }
}

#endif  // __CUDACC__

}  // namespace nmb
