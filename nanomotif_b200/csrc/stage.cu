// Host -> device staging for PAGEABLE host tables.
//
// The frames nanomotif hands to its workers (find_motifs_bin.py:399-427) live in ordinary (pageable) host memory: Arrow
// buffers of a polars frame.  cudaMemcpy from pageable memory is staged by the driver through one small pinned buffer
// on one thread (measured 4.6 GB/s for a 3.2 GB table on the B200 box); page-locking the source in place
// (cudaHostRegister) costs more than the copy.  The stager keeps T host threads, each with its own CUDA stream and two
// pinned slots: a thread memcpy()s chunk k of the source into a slot while the DMA engine drains the other slot, so the
// copy runs at min(T x memcpy rate, PCIe rate).  The copy is ordered after the work already enqueued on the caller's
// stream and the caller's stream is ordered after it; the call returns once the SOURCE has been read (the caller may
// free it), not when the DMA has finished.
#include <string.h>

#include <thread>
#include <vector>

#include "common.cuh"

struct nmb_stager {
    int device;
    int n_threads;
    int64_t slot_bytes;
    std::vector<cudaStream_t> streams;
    std::vector<uint8_t *> slots;       // 2 per thread
    std::vector<cudaEvent_t> slot_done; // 2 per thread
    std::vector<cudaEvent_t> done;      // 1 per thread
    cudaEvent_t begin;
};

extern "C" {

int nmb_stager_create(int64_t slot_bytes, int32_t n_threads, nmb_stager **out) {
    NMB_REQUIRE(out && slot_bytes >= 4096 && n_threads >= 1 && n_threads <= 64, "nmb_stager_create: bad arguments");
    nmb_stager *s = new nmb_stager();
    NMB_CUDA(cudaGetDevice(&s->device));
    s->n_threads = n_threads;
    s->slot_bytes = slot_bytes;
    NMB_CUDA(cudaEventCreateWithFlags(&s->begin, cudaEventDisableTiming));
    for (int t = 0; t < n_threads; ++t) {
        cudaStream_t st;
        NMB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        s->streams.push_back(st);
        cudaEvent_t e;
        NMB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        s->done.push_back(e);
        for (int k = 0; k < 2; ++k) {
            uint8_t *p = nullptr;
            NMB_CUDA(cudaHostAlloc((void **)&p, (size_t)slot_bytes, cudaHostAllocDefault));
            s->slots.push_back(p);
            NMB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            s->slot_done.push_back(e);
        }
    }
    *out = s;
    return NMB_OK;
}

int nmb_stager_destroy(nmb_stager *s) {
    if (!s) return NMB_OK;
    for (auto st : s->streams) cudaStreamSynchronize(st);
    for (auto p : s->slots) cudaFreeHost(p);
    for (auto e : s->slot_done) cudaEventDestroy(e);
    for (auto e : s->done) cudaEventDestroy(e);
    for (auto st : s->streams) cudaStreamDestroy(st);
    cudaEventDestroy(s->begin);
    delete s;
    return NMB_OK;
}

// Copy bytes [off, off + n) of the logical concatenation of the pieces into dst.
static void gather_bytes(uint8_t *dst, const void *const *srcs, const int64_t *piece_off, int64_t n_src, int64_t off,
                         int64_t n) {
    int64_t lo = 0, hi = n_src;  // largest piece with piece_off[piece] <= off
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (piece_off[mid] <= off) lo = mid; else hi = mid;
    }
    for (int64_t p = lo; n > 0 && p < n_src; ++p) {
        const int64_t within = off - piece_off[p];
        int64_t take = piece_off[p + 1] - off;
        if (take > n) take = n;
        if (take > 0) memcpy(dst, (const uint8_t *)srcs[p] + within, (size_t)take);
        dst += take; off += take; n -= take;
    }
}

static int stager_run(nmb_stager *s, void *dst_dev, const void *src_host, const void *const *srcs,
                      const int64_t *piece_off, int64_t n_src, int64_t bytes, void *stream, bool narrow = false,
                      int32_t *overflow_h = nullptr);

// dst[i] = (int32) src[i]; returns true when a value did not fit
static bool narrow_i64(int32_t *dst, const int64_t *src, int64_t n) {
    int64_t bad = 0;
    for (int64_t i = 0; i < n; ++i) {
        const int64_t v = src[i];
        dst[i] = (int32_t)v;
        bad |= v ^ (int64_t)(int32_t)v;
    }
    return bad != 0;
}

int nmb_stager_copy(nmb_stager *s, void *dst_dev, const void *src_host, int64_t bytes, void *stream) {
    NMB_REQUIRE(s && bytes >= 0, "nmb_stager_copy: bad arguments");
    if (bytes == 0) return NMB_OK;
    NMB_REQUIRE(dst_dev && src_host, "nmb_stager_copy: null buffer");
    return stager_run(s, dst_dev, src_host, nullptr, nullptr, 0, bytes, stream);
}

int nmb_stager_copy_narrow(nmb_stager *s, int32_t *dst_dev, const int64_t *src_host, int64_t n, int32_t *overflow_h,
                           void *stream) {
    NMB_REQUIRE(s && n >= 0 && overflow_h, "nmb_stager_copy_narrow: bad arguments");
    *overflow_h = 0;
    if (n == 0) return NMB_OK;
    NMB_REQUIRE(dst_dev && src_host, "nmb_stager_copy_narrow: null buffer");
    return stager_run(s, dst_dev, src_host, nullptr, nullptr, 0, n * 4, stream, true, overflow_h);
}

int nmb_stager_gather(nmb_stager *s, void *dst_dev, const void *const *srcs_h, const int64_t *piece_off_h,
                      int64_t n_src, void *stream) {
    NMB_REQUIRE(s && n_src >= 0, "nmb_stager_gather: bad arguments");
    if (n_src == 0) return NMB_OK;
    NMB_REQUIRE(dst_dev && srcs_h && piece_off_h, "nmb_stager_gather: null argument");
    const int64_t bytes = piece_off_h[n_src];
    NMB_REQUIRE(piece_off_h[0] == 0 && bytes >= 0, "nmb_stager_gather: piece offsets must start at 0 and ascend");
    if (bytes == 0) return NMB_OK;
    return stager_run(s, dst_dev, nullptr, srcs_h, piece_off_h, n_src, bytes, stream);
}

static int stager_run(nmb_stager *s, void *dst_dev, const void *src_host, const void *const *srcs,
                      const int64_t *piece_off, int64_t n_src, int64_t bytes, void *stream, bool narrow,
                      int32_t *overflow_h) {
    cudaStream_t caller = (cudaStream_t)stream;
    int current = s->device;
    NMB_CUDA(cudaGetDevice(&current));
    if (current != s->device) NMB_FAIL(NMB_ERR_INVALID, "nmb_stager: the stager belongs to device %d, the current device is %d",
                                       s->device, current);
    NMB_CUDA(cudaEventRecord(s->begin, caller));  // dst may still be in use by work enqueued before this call
    const int64_t n_chunks = (bytes + s->slot_bytes - 1) / s->slot_bytes;
    const int T = (int)(n_chunks < s->n_threads ? n_chunks : s->n_threads);
    std::vector<cudaError_t> err(T, cudaSuccess);
    std::vector<int> overflow(T, 0);
    auto body = [&](int t) {
        cudaError_t e = cudaSetDevice(s->device);
        cudaStream_t st = s->streams[t];
        if (e == cudaSuccess) e = cudaStreamWaitEvent(st, s->begin, 0);
        int k = 0;
        for (int64_t c = t; c < n_chunks && e == cudaSuccess; c += T, ++k) {
            const int slot = 2 * t + (k & 1);
            const int64_t off = c * s->slot_bytes;
            const int64_t n = bytes - off < s->slot_bytes ? bytes - off : s->slot_bytes;
            if (k >= 2) e = cudaEventSynchronize(s->slot_done[slot]);  // the DMA that last read this slot
            if (e != cudaSuccess) break;
            if (narrow)  // off / n count destination bytes: int32 elements made from int64 ones
                overflow[t] |= narrow_i64((int32_t *)s->slots[slot], (const int64_t *)src_host + off / 4, n / 4);
            else if (src_host) memcpy(s->slots[slot], (const uint8_t *)src_host + off, (size_t)n);
            else gather_bytes(s->slots[slot], srcs, piece_off, n_src, off, n);
            e = cudaMemcpyAsync((uint8_t *)dst_dev + off, s->slots[slot], (size_t)n, cudaMemcpyHostToDevice, st);
            if (e == cudaSuccess) e = cudaEventRecord(s->slot_done[slot], st);
        }
        if (e == cudaSuccess) e = cudaEventRecord(s->done[t], st);
        err[t] = e;
    };
    // slots are reused by the NEXT call: their last DMAs must have drained before a thread overwrites them
    for (int t = 0; t < T; ++t) {
        NMB_CUDA(cudaEventSynchronize(s->slot_done[2 * t]));
        NMB_CUDA(cudaEventSynchronize(s->slot_done[2 * t + 1]));
    }
    std::vector<std::thread> threads;
    for (int t = 1; t < T; ++t) threads.emplace_back(body, t);
    body(0);
    for (auto &th : threads) th.join();
    for (int t = 0; t < T; ++t) {
        if (err[t] != cudaSuccess)
            NMB_FAIL(NMB_ERR_CUDA, "nmb_stager: %s", cudaGetErrorString(err[t]));
        NMB_CUDA(cudaStreamWaitEvent(caller, s->done[t], 0));
        if (overflow_h && overflow[t]) *overflow_h = 1;
    }
    return NMB_OK;
}

}  // extern "C"
