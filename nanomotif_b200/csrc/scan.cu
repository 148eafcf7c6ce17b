// K2 -- fused motif scan + gather-join + segmented reduce, and the motif compiler.
//
// Replaces, for a whole batch of (motif, contig-set) pairs in one launch:
//   utils.subseq_indices                      nanomotif/utils.py:44-67        (regex scan)
//   methylated_motif_occourances              nanomotif/find_motifs_bin.py:1234-1263 (np.isin join)
//   motif_model_contig (count step)           nanomotif/find_motifs_bin.py:1285-1331
//   motif_model_bin (sum over contigs)        nanomotif/find_motifs_bin.py:1265-1283
//
// Persistent CTAs draw work items (tile, block of <=32 motifs) from a device counter.  A tile is one
// self-contained 17 KB sequence record + one 32 KB class record, brought into shared memory by two
// cp.async.bulk (TMA) copies, each completing on its own mbarrier; the next item's sequence record is
// copied while the current item's motifs are evaluated.  Four CTAs of 128 threads are resident per
// SM, so the remaining copy latency is hidden by the other CTAs' bit-parallel evaluation.  Counts are popcounts of
// match & class-plane, reduced with warp REDUX, accumulated per CTA in shared memory and flushed
// with one 64-bit atomic per (motif, counter) per item.
#include "scan.cuh"

namespace nmb {

constexpr int kScanThreads = kTileChunks;                                       // 128
constexpr int kMaxMpi = NMB_MAX_MOTIFS_PER_ITEM;

// A FAMILY is a run of consecutive motifs of one work item that share all constrained positions but one -- the children
// of one search expansion (find_motifs_bin.py:1116-1145 adds one base at one position to the same parent).  The parent's
// two chains are evaluated once; a child is then one indicator plane, shifted to the modified base's alignment, ANDed
// with the parent's aligned planes: ~6 logic ops per word and strand instead of 4 per word, strand and constrained
// position.  Offsets are relative to the modified base, which all members share.
struct FamInfo {
    uint8_t n;      // at the first member (leader): members in the run (>= 2); 0 elsewhere / for single motifs
    uint8_t set;    // allowed-set of the member's extra position
    int8_t delta;   // its offset from the modified base, |delta| <= 31
    uint8_t pad;
};
static_assert(sizeof(FamInfo) == 4, "FamInfo is 4 bytes");
static_assert(NMB_MAX_MOTIFS_PER_ITEM == 32, "one lane per motif of a block holds its family word");

struct ScanParams {
    const uint32_t *seq_records;
    const uint32_t *nonacgt;
    const uint32_t *cls;
    const Program *programs;
    const nmb_job *jobs;
    const int32_t *contig_group;
    const int64_t *contig_start, *contig_len;
    unsigned long long *out;
    int *counter;             // dynamic item scheduling: {next item, finished CTAs}, zero at launch and at exit; or null
    const FamInfo *fam;       // per motif: family run length at a leader, the member's extra (set, offset); or null
    const Program *parents;   // per motif: the family's parent program (valid at leaders)
    int n_jobs, n_items, mpi, n_tiles;
};

struct ItemMeta {
    int job, tile, mblk, pad;
};

__device__ __forceinline__ ItemMeta decode_item(const ScanParams &p, int item) {
    int lo = 0, hi = p.n_jobs;  // largest j with item_offset[j] <= item
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(&p.jobs[mid].item_offset) <= item) lo = mid; else hi = mid;
    }
    const int local = item - __ldg(&p.jobs[lo].item_offset);
    const int nblk = (__ldg(&p.jobs[lo].motif_count) + p.mpi - 1) / p.mpi;
    ItemMeta m;
    m.job = lo;
    m.tile = __ldg(&p.jobs[lo].tile_begin) + local / nblk;  // tile-major: concurrent CTAs share a tile in L2
    m.mblk = local % nblk;
    m.pad = 0;
    return m;
}

__device__ __forceinline__ int group_of(const nmb_job &job, int contig, const int32_t *contig_group) {
    if (contig < job.contig_begin || contig >= job.contig_end) return -1;
    if (job.group_mode == 0) return 0;
    if (job.group_mode == 1) return contig - job.contig_begin;
    return __ldg(contig_group + contig);
}

// Indicator plane of allowed-set `code` over the lane's words [-H, NW + H) (all fourteen sets; 0 for anything else).
template <int H, bool HASN>
__device__ __forceinline__ void set_indicator(const LaneSeq<H, HASN> &q, int code, uint32_t (&ind)[NW + 2 * H]) {
    switch (code) {
#define NMB_CASE(m)                                                                            \
    case m:                                                                                    \
        _Pragma("unroll") for (int i = 0; i < NW + 2 * H; ++i) ind[i] = q.template and_code<m>(i, 0xFFFFFFFFu); \
        break;
        NMB_CASE(1) NMB_CASE(2) NMB_CASE(3) NMB_CASE(4) NMB_CASE(5) NMB_CASE(6) NMB_CASE(7)
        NMB_CASE(8) NMB_CASE(9) NMB_CASE(10) NMB_CASE(11) NMB_CASE(12) NMB_CASE(13) NMB_CASE(14)
#undef NMB_CASE
        default:
#pragma unroll
            for (int i = 0; i < NW + 2 * H; ++i) ind[i] = 0u;
            break;
    }
}

// One strand of one family member: the parent's aligned match words m[0..NW) ANDed with the indicator of the extra
// position's set seen `shift` positions to the right (negative: to the left), popcounted against two class planes.
template <int H, bool HASN>
__device__ __forceinline__ void count_member_strand(const LaneSeq<H, HASN> &q, const uint32_t (&m)[NW], int code, int shift,
                                                    const uint32_t *cl_a, const uint32_t *cl_b, uint32_t &cnt_a,
                                                    uint32_t &cnt_b) {
    constexpr int CW = NW + 2 * H;
    uint32_t ind[CW];
    set_indicator<H, HASN>(q, code, ind);
    if (shift < 0) {  // warp-uniform: look one word to the left, then shift right by 32 - |shift|
#pragma unroll
        for (int i = CW - 1; i > 0; --i) ind[i] = ind[i - 1];
        ind[0] = 0u;
    }
    const int s = shift & 31;
#pragma unroll
    for (int h = 0; h < NW; h += 4) {
        const uint4 pa = *reinterpret_cast<const uint4 *>(cl_a + (h >> 2) * kSlotStride);
        const uint4 pb = *reinterpret_cast<const uint4 *>(cl_b + (h >> 2) * kSlotStride);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t w = m[h + k] & __funnelshift_r(ind[h + k + H], ind[h + k + H + 1], s);
            cnt_a += __popc(w & (k == 0 ? pa.x : k == 1 ? pa.y : k == 2 ? pa.z : pa.w));
            cnt_b += __popc(w & (k == 0 ? pb.x : k == 1 ? pb.y : k == 2 ? pb.z : pb.w));
        }
    }
}

// Evaluate the item's motifs on this lane's chunk and accumulate the four counters.
template <int H, bool PLANES, bool FAM>
__device__ __forceinline__ void score_motifs(const ScanParams &p, const nmb_job &job, const ItemMeta &meta,
                                             const LaneSeq<H, PLANES> &q, const LaneEdge &edge, const uint32_t *cl,
                                             bool valid,
                                             bool uniform, unsigned peers, bool leader, int g, int primary,
                                             uint32_t (*acc)[4]) {
    const int m_begin = job.motif_begin + meta.mblk * p.mpi;
    const int m_count = min(p.mpi, job.motif_count - meta.mblk * p.mpi);
    // family table of the whole block in ONE coalesced load (lane l holds motif l's word; kMaxMpi == 32): looked up
    // with shuffles below, so no motif waits for a dependent global load before its program can be fetched
    uint32_t fam_w = 0;
    if (FAM && !PLANES && !edge.edge && (int)(threadIdx.x & 31) < m_count)
        fam_w = __ldg(reinterpret_cast<const uint32_t *>(p.fam + m_begin) + (threadIdx.x & 31));

    // per-lane counts of one motif -> warp sums -> the CTA's accumulators / the output
    auto flush = [&](int mi, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
        // n_mod | n_nomod << 16 per strand (per-lane counts are <= 512)
        uint32_t pk_f = valid ? (c0 | (c1 << 16)) : 0u;
        uint32_t pk_r = valid ? (c2 | (c3 << 16)) : 0u;
        const long long row = job.out_base + (long long)(meta.mblk * p.mpi + mi) * job.n_groups;
        // two 16-bit fields per word survive a 32-lane sum (<= 8192)
        if (uniform) {  // all counted lanes of the warp feed the same output row (the common case)
            pk_f = __reduce_add_sync(0xFFFFFFFFu, pk_f);
            pk_r = __reduce_add_sync(0xFFFFFFFFu, pk_r);
        } else {  // contig / group boundary inside the warp: segmented sums over equal-group lanes
            pk_f = __reduce_add_sync(peers, pk_f);
            pk_r = __reduce_add_sync(peers, pk_r);
        }
        if (leader && (pk_f | pk_r)) {
            const uint32_t v[4] = {pk_f & 0xFFFFu, pk_f >> 16, pk_r & 0xFFFFu, pk_r >> 16};
            if (g == primary) {
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (v[c]) atomicAdd(&acc[mi][c], v[c]);
            } else {
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (v[c]) atomicAdd(p.out + (row + g) * 4 + c, (unsigned long long)v[c]);
            }
        }
    };

    int mi = 0;
#pragma unroll 1
    while (mi < m_count) {
        // A family (see FamInfo) is taken as one parent + members only by warps without non-ACGT letters or contig
        // edges in reach; the other warps take the members one by one (every member keeps its own program).
        int run = 1;
        const Program *prog = p.programs + (size_t)(m_begin + mi) * 2;
        if (FAM && !PLANES && !edge.edge) {  // warp-uniform
            const int n = __shfl_sync(0xFFFFFFFFu, fam_w, mi) & 0xFF;
            if (n >= 2) {
                run = n;
                prog = p.parents + (m_begin + mi);
            }
        }
        // ONE program serves both strands (scan.cuh: run_chain_pair) -- the motif's own, or the family's parent
        const ProgramView pv = load_program(prog);
        uint32_t c[NW + 2 * H], d[NW + 2 * H];
        if (run_chain_pair<H, PLANES>(pv, q, c, d, edge)) {  // else: no occurrence in this warp's chunks
            const bool far = pv.mod_pos >= 32;  // only possible when H == 2
            const int sh = pv.mod_pos & 31;
            if (!FAM || run == 1) {
                uint32_t cnt[4] = {0, 0, 0, 0};  // n_mod '+', n_nomod '+', n_mod '-', n_nomod '-'
#pragma unroll
                for (int h = 0; h < NW; h += 4) {
                    uint4 pl[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        pl[k] = *reinterpret_cast<const uint4 *>(cl + k * kTileWords + (h >> 2) * kSlotStride);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t mf = aligned_word<H>(c, h + k, sh, far);
                        const uint32_t mr = aligned_word_rc<H>(d, h + k, sh, far);
                        const uint32_t w0 = k == 0 ? pl[0].x : k == 1 ? pl[0].y : k == 2 ? pl[0].z : pl[0].w;
                        const uint32_t w1 = k == 0 ? pl[1].x : k == 1 ? pl[1].y : k == 2 ? pl[1].z : pl[1].w;
                        const uint32_t w2 = k == 0 ? pl[2].x : k == 1 ? pl[2].y : k == 2 ? pl[2].z : pl[2].w;
                        const uint32_t w3 = k == 0 ? pl[3].x : k == 1 ? pl[3].y : k == 2 ? pl[3].z : pl[3].w;
                        cnt[0] += __popc(mf & w0);  // occurrences whose modified base is methylated, '+'
                        cnt[1] += __popc(mf & w1);  // ... unmethylated, '+'
                        cnt[2] += __popc(mr & w2);  // reverse-complement occurrences, '-' strand rows
                        cnt[3] += __popc(mr & w3);
                    }
                }
                flush(mi, cnt[0], cnt[1], cnt[2], cnt[3]);
            } else {
                // the parent's two match planes aligned at the modified base; the chains are dead after this
                uint32_t mf[NW], mr[NW], any = 0;
#pragma unroll
                for (int k = 0; k < NW; ++k) {
                    mf[k] = aligned_word<H>(c, k, sh, far);
                    mr[k] = aligned_word_rc<H>(d, k, sh, far);
                    any |= mf[k] | mr[k];
                }
                if (__any_sync(0xFFFFFFFFu, any != 0)) {  // else every member counts 0 in this warp
#pragma unroll 1
                    for (int k = 0; k < run; ++k) {
                        const uint32_t fi = __shfl_sync(0xFFFFFFFFu, fam_w, mi + k);
                        const int code = (fi >> 8) & 0xF;
                        const int delta = (int)(int8_t)((fi >> 16) & 0xFF);
                        uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
                        // forward: the member occurs at p iff the parent does and base[p + delta] is in the set;
                        // reverse complement (aligned at ITS modified base q): base[q - delta] in the complementary set
                        count_member_strand<H, PLANES>(q, mf, code, delta, cl, cl + kTileWords, c0, c1);
                        count_member_strand<H, PLANES>(q, mr, comp_set(code), -delta, cl + 2 * kTileWords,
                                                       cl + 3 * kTileWords, c2, c3);
                        flush(mi + k, c0, c1, c2, c3);
                    }
                }
            }
        }
        mi += run;
    }
}

// One tile (sequence record + class record = 49.7 KB) per CTA in shared memory; four CTAs of four warps are resident
// per SM, so while one CTA waits for its bulk copies the other three keep the ALUs busy.  (Measured alternatives on
// B200: a 2-stage ring with two resident CTAs, and a split ring with three, were both slower -- profiles/r01_notes.md.)
constexpr int kScanSmemBytes = kSeqRecBytes + kClsRecBytes;

__device__ __forceinline__ void bar_sync_1(int n_threads) { asm volatile("bar.sync 1, %0;" ::"r"(n_threads) : "memory"); }

// The two records of a tile arrive on SEPARATE mbarriers.  The lanes copy their sequence words into registers at the
// start of an item, after which the sequence half of the buffer is dead: the next item's index is decoded and its
// sequence record is already on its way while this item's motifs are evaluated; only the class record (2/3 of the
// bytes) has to wait for the end of the item.  ncu's source view had 19 % of all warp samples on the mbarrier spin in
// front of a tile (profiles/r02_notes.md).
template <int H, int STAGES, bool FAM>
__global__ void __launch_bounds__(kScanThreads, 4) scan_count_kernel(const ScanParams p) {
    static_assert(STAGES == 1, "one tile buffer per CTA");
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t seq_bar, cls_bar;
    __shared__ ItemMeta s_meta[2];
    __shared__ uint32_t s_acc[2][kMaxMpi][4];

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    if (tid == 0) {
        mbar_init(&seq_bar, 1);
        mbar_init(&cls_bar, 1);
        fence_barrier_init();
    }
    for (int i = tid; i < 2 * kMaxMpi * 4; i += kScanThreads) (&s_acc[0][0][0])[i] = 0;
    __syncthreads();

    // Items are handed out through a global counter when the caller provides one: a CTA that drew cheap items (3
    // instead of 4 motifs, dead chains) simply takes more of them, so the launch ends when the WORK runs out, not
    // when the unluckiest CTA of a static round-robin finishes.  Items still start in tile-major order.  The index
    // of the next item is drawn as soon as the previous draw has been used, so the atomic's round trip is hidden.
    const bool dynamic = p.counter != nullptr;
    int next_item = 0, n_drawn = 0;  // thread 0 only
    auto draw = [&]() {
        next_item = dynamic ? atomicAdd(p.counter, 1) : (int)blockIdx.x + n_drawn * (int)gridDim.x;
        ++n_drawn;
    };
    // thread 0: decode the drawn item into s_meta[slot] and start the copy of its sequence record; an exhausted
    // work list is signalled through an empty phase of the sequence barrier
    auto issue_seq = [&](int slot) {
        const int item = next_item;
        if (item >= p.n_items) {
            s_meta[slot].job = -1;
            mbar_arrive(&seq_bar);
            return;
        }
        const ItemMeta m = decode_item(p, item);
        s_meta[slot] = m;
        fence_proxy_async();  // order the earlier generic reads of this buffer before the async writes
        mbar_expect_tx(&seq_bar, kSeqRecBytes);
        bulk_g2s(smem, p.seq_records + (size_t)m.tile * kSeqRecWords, kSeqRecBytes, &seq_bar);
        draw();
    };
    auto issue_cls = [&](int slot) {
        const ItemMeta m = s_meta[slot];
        if (m.job < 0) return;
        const int modtype = __ldg(&p.jobs[m.job].modtype);
        fence_proxy_async();
        mbar_expect_tx(&cls_bar, kClsRecBytes);
        bulk_g2s(smem + kSeqRecBytes, p.cls + ((size_t)modtype * p.n_tiles + m.tile) * kClsRecWords, kClsRecBytes, &cls_bar);
    };
    if (tid == 0) {
        draw();
        issue_seq(0);
        issue_cls(0);
    }

    const uint32_t *sx = reinterpret_cast<const uint32_t *>(smem);
    const uint32_t *sy = sx + kSeqPlaneWords;
    const int32_t *sinfo = reinterpret_cast<const int32_t *>(sy + kSeqPlaneWords);
    const uint32_t *scls = sx + kSeqRecWords;
    for (int k = 0;; ++k) {
        const int par = k & 1;
        mbar_wait(&seq_bar, (uint32_t)par);
        const ItemMeta meta = s_meta[par];
        if (meta.job < 0) break;  // CTA-uniform: every thread reads the same word
        const nmb_job job = p.jobs[meta.job];

        const int info = sinfo[tid];
        const int contig = info < 0 ? -1 : (info & kChunkIdMask);
        const int g = contig < 0 ? -1 : group_of(job, contig, p.contig_group);
        const bool valid = g >= 0;
        // group whose counts go through the CTA's shared accumulators: the one of the tile's first chunk
        const int info0 = sinfo[0];
        const int primary = info0 < 0 ? -1 : group_of(job, info0 & kChunkIdMask, p.contig_group);

        const unsigned vmask = __ballot_sync(0xFFFFFFFFu, valid);
        // lanes that are not counted (padding, contigs outside the job) contribute zeros, so the warp
        // is uniform when all COUNTED lanes share one group
        const int first = vmask ? __ffs(vmask) - 1 : 0;
        const int g0 = __shfl_sync(0xFFFFFFFFu, g, first);
        const bool uniform = __all_sync(0xFFFFFFFFu, !valid || g == g0);
        unsigned peers = 0xFFFFFFFFu;
        bool leader = lane == first;
        if (!uniform) {
            peers = __match_any_sync(0xFFFFFFFFu, g);
            leader = valid && lane == __ffs(peers) - 1;
        }
        // matcher variant (scan.cuh): non-ACGT letters of a contig in reach -> full test at every step
        // (rare); inter-contig padding in reach -> plain steps, finished chains clipped to the contig
        const bool warp_n = __any_sync(0xFFFFFFFFu, valid && (info & kChunkFlagN));
        const bool warp_edge = __any_sync(0xFFFFFFFFu, valid && (info & kChunkFlagEdge));
        const uint32_t *cl = scls + tid * 4;  // lane-interleaved planes: vector v at cl + v * kSlotStride

        // after this point nobody reads the sequence half of the buffer any more
        auto seq_done_prefetch_next = [&]() {
            bar_sync_1(kScanThreads);
            if (tid == 0) issue_seq(par ^ 1);
            mbar_wait(&cls_bar, (uint32_t)par);
        };
        if (!vmask) {
            seq_done_prefetch_next();
        } else if (warp_n) {
            LaneSeq<H, true> q;
            load_xyn<H>(sx, sy, tid, p.nonacgt + kHalo + (size_t)meta.tile * kTileWords + tid * NW - H, q);
            seq_done_prefetch_next();
            const LaneEdge edge = {0, false, false};
            score_motifs<H, true, FAM>(p, job, meta, q, edge, cl, valid, uniform, peers, leader, g, primary, s_acc[par]);
        } else {
            LaneSeq<H, false> q;
            load_xy<H>(sx, sy, tid, q);
            seq_done_prefetch_next();
            const LaneEdge edge = lane_edge(warp_edge, info, (int64_t)meta.tile * kTileChunks + tid,
                                            p.contig_start, p.contig_len);
            score_motifs<H, false, FAM>(p, job, meta, q, edge, cl, valid, uniform, peers, leader, g, primary, s_acc[par]);
        }
        __syncthreads();  // everyone is done with the class record and with s_acc[par]
        if (tid == 0) issue_cls(par ^ 1);
        if (tid < kMaxMpi * 4) {
            const int mi = tid >> 2, c = tid & 3;
            const uint32_t v = s_acc[par][mi][c];
            if (v) {
                s_acc[par][mi][c] = 0;
                const long long row =
                    job.out_base + (long long)(meta.mblk * p.mpi + mi) * job.n_groups + primary;
                atomicAdd(p.out + row * 4 + c, (unsigned long long)v);
            }
        }
    }
    // every CTA has drawn its last (empty) item before it gets here: the last one to leave re-arms the counter
    if (dynamic && tid == 0 && atomicAdd(p.counter + 1, 1) == (int)gridDim.x - 1) {
        p.counter[0] = 0;
        p.counter[1] = 0;
    }
}

// ---------------------------------------------------------------------------------------------
// motif compiler: nmb_motif -> forward and reverse-complement Programs
// ---------------------------------------------------------------------------------------------
// Program of a motif strand given as allowed-sets per position (0xF = wildcard) in reading order.
// Gaps of 32 or more positions become "shift one word" pseudo entries so that every real entry carries a shift < 32.
// The first processed entry is the last motif position (shift 0); a stripped motif starts with a constrained position,
// so the chain ends aligned at position 0.
__device__ void build_program(const uint8_t *allowed, int len, int mp, Program &pr) {
    for (int i = 0; i < kMaxLen; ++i) pr.ent[i] = 0;
    pr.reserved = 0;
    int n = 0, prev = -1;
    for (int j = len - 1; j >= 0; --j) {
        const int a = allowed[j] & 0xF;
        if (a == 0xF && j != 0) continue;
        int d = prev < 0 ? 0 : prev - j;
        for (; d >= 32 && n < kMaxLen - 1; d -= 32) pr.ent[n++] = kEntShift32;
        if (n < kMaxLen) pr.ent[n++] = (uint16_t)(a | (d << 8));  // a == 0xF at j == 0: pure shift
        prev = j;
    }
    pr.n = (uint8_t)n; pr.mod_pos = (uint8_t)mp; pr.len = (uint8_t)len;
}

__global__ void compile_motifs_kernel(const nmb_motif *__restrict__ motifs, int n_motifs,
                                      Program *__restrict__ programs) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * n_motifs) return;
    const nmb_motif mt = motifs[t >> 1];
    const bool rc = t & 1;
    Program pr;
    int len = mt.len, mp = mt.mod_pos;
    if (len < 1 || len > kMaxLen || mp >= len) {  // invalid: compile to "never matches"
        for (int i = 0; i < kMaxLen; ++i) pr.ent[i] = 0;
        pr.reserved = 0;
        pr.n = 1; pr.mod_pos = 0; pr.len = 1;
        programs[t] = pr;
        return;
    }
    uint8_t allowed[kMaxLen];
    for (int j = 0; j < len; ++j) {
        int a = mt.allowed[rc ? (len - 1 - j) : j] & 0xF;
        if (rc) a = ((a & 5) << 1) | ((a & 10) >> 1);  // A<->T, G<->C (constants.py:14-20)
        allowed[j] = (uint8_t)a;
    }
    if (rc) mp = len - 1 - mp;  // motif.py:264
    build_program(allowed, len, mp, pr);
    programs[t] = pr;
}

// ---------------------------------------------------------------------------------------------
// family finder: runs of consecutive motifs inside one work item's motif block that share a parent
// ---------------------------------------------------------------------------------------------
// One WARP per motif block, lane l = motif l of the block (<= 32).  Offsets are relative to the modified base and
// limited to [-31, 31] (a member's extra position must be reachable with one halo word and the parent must fit one
// word span); motifs reaching further are never family members.  Pass 1: every lane tests whether it and its right
// neighbour are siblings (same number of constrained positions, all but one shared).  Then leaders are walked left
// to right: without a sibling to the right a motif is single; otherwise every lane to the right tests itself
// against the parent P = (leader AND leader + 1) and the run extends while consecutive lanes pass.
constexpr int kFamReach = 31;

__device__ __forceinline__ int fam_set_at(const nmb_motif *m, int len, int mod, int o) {  // 0xF outside / wildcard
    const int j = o + mod;
    return (j >= 0 && j < len) ? (__ldg(&m->allowed[j]) & 0xF) : 0xF;
}

__global__ void __launch_bounds__(256) find_families_kernel(const nmb_motif *__restrict__ motifs,
                                                            const nmb_job *__restrict__ jobs, int mpi,
                                                            FamInfo *__restrict__ fam, Program *__restrict__ parents) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    const int motif_begin = __ldg(&jobs[blockIdx.x].motif_begin), motif_count = __ldg(&jobs[blockIdx.x].motif_count);
    const int n_blocks = (motif_count + mpi - 1) / mpi;
    for (int blk = blockIdx.y * wpc + warp; blk < n_blocks; blk += gridDim.y * wpc) {
        const int m0 = motif_begin + blk * mpi;
        const int cnt = min(mpi, motif_count - blk * mpi);
        const bool have = lane < cnt;
        const nmb_motif *me = motifs + m0 + (have ? lane : 0);
        const int len = have ? __ldg(&me->len) : 0, mod = have ? __ldg(&me->mod_pos) : 0;
        // eligible: valid record whose positions all lie within [-31, 31] of the modified base
        const bool elig = have && len >= 1 && len <= kMaxLen && mod < len && mod <= kFamReach && len - 1 - mod <= kFamReach;
        FamInfo mine = {0, 0, 0, 0};
        // ---- pass 1: sibling of the right neighbour? ----
        bool sib = false;
        {
            const nmb_motif *nb = me + 1;
            const bool nb_have = lane + 1 < cnt;
            const int nlen = nb_have ? __ldg(&nb->len) : 0, nmod = nb_have ? __ldg(&nb->mod_pos) : 0;
            const bool nelig = nb_have && nlen >= 1 && nlen <= kMaxLen && nmod < nlen && nmod <= kFamReach &&
                               nlen - 1 - nmod <= kFamReach;
            if (elig && nelig) {
                int ka = 0, kb = 0, common = 0, lo = 99, hi = -99;
                bool bad = false;
                for (int o = -kFamReach; o <= kFamReach; ++o) {
                    const int a = fam_set_at(me, len, mod, o), b = fam_set_at(nb, nlen, nmod, o);
                    bad |= a == 0 || b == 0;  // an empty set never matches: leave such motifs to the general path
                    ka += a != 0xF;
                    kb += b != 0xF;
                    if (a != 0xF && a == b) { ++common; lo = min(lo, o); hi = max(hi, o); }
                }
                sib = !bad && ka == kb && ka >= 2 && common == ka - 1 && lo <= 0 && hi >= 0 && hi - lo + 1 <= 32 &&
                      fam_set_at(me, len, mod, 0) != 0xF && fam_set_at(me, len, mod, 0) == fam_set_at(nb, nlen, nmod, 0);
            }
        }
        const unsigned adj = __ballot_sync(0xFFFFFFFFu, sib);
        // ---- walk the leaders ----
        int leader = 0;
        while (leader < cnt) {  // warp-uniform
            int run = 1;
            if ((adj >> leader) & 1u) {
                const nmb_motif *A = motifs + m0 + leader, *B = A + 1;
                const int alen = __ldg(&A->len), amod = __ldg(&A->mod_pos), blen = __ldg(&B->len), bmod = __ldg(&B->mod_pos);
                // does this lane's motif hold every position of P = A & B, plus exactly one more?
                bool ok = elig && lane >= leader;
                int extra_set = 0, extra_o = 0, n_extra = 0, k_me = 0, k_par = 0;
                if (ok) {
                    for (int o = -kFamReach; o <= kFamReach; ++o) {
                        const int a = fam_set_at(A, alen, amod, o), b = fam_set_at(B, blen, bmod, o);
                        const int c = fam_set_at(me, len, mod, o);
                        const bool in_par = a != 0xF && a == b;
                        k_par += in_par;
                        k_me += c != 0xF;
                        if (in_par) ok &= c == a;
                        else if (c != 0xF) { ++n_extra; extra_set = c; extra_o = o; }
                        ok &= c != 0;
                    }
                    ok &= n_extra == 1 && k_me == k_par + 1;
                }
                const unsigned pass = __ballot_sync(0xFFFFFFFFu, ok) >> leader;  // bit 0 = the leader itself
                run = pass == 0xFFFFFFFFu ? 32 : __ffs(~pass) - 1;                // consecutive members from the leader on
                if (run < 2) run = 1;
                else {
                    if (lane >= leader && lane < leader + run) {
                        mine.n = lane == leader ? (uint8_t)run : 0;
                        mine.set = (uint8_t)extra_set;
                        mine.delta = (int8_t)extra_o;
                    }
                    if (lane == leader) {  // the parent's program: positions of P over its span, modified base inside
                        uint8_t allowed[32];
                        int lo = 99, hi = -99;
                        for (int o = -kFamReach; o <= kFamReach; ++o) {
                            const int a = fam_set_at(A, alen, amod, o), b = fam_set_at(B, blen, bmod, o);
                            if (a != 0xF && a == b) { lo = min(lo, o); hi = max(hi, o); }
                        }
                        for (int o = lo; o <= hi; ++o) {
                            const int a = fam_set_at(A, alen, amod, o), b = fam_set_at(B, blen, bmod, o);
                            allowed[o - lo] = (uint8_t)((a != 0xF && a == b) ? a : 0xF);
                        }
                        Program pr;
                        build_program(allowed, hi - lo + 1, -lo, pr);
                        parents[m0 + leader] = pr;
                    }
                }
            }
            leader += run;
        }
        if (have) fam[m0 + lane] = mine;
    }
}

}  // namespace nmb

extern "C" {

int nmb_compile_motifs(const nmb_motif *motifs, int32_t n_motifs, void *programs, void *stream) {
    NMB_REQUIRE(n_motifs >= 0, "nmb_compile_motifs: n_motifs=%d", n_motifs);
    if (n_motifs == 0) return NMB_OK;
    NMB_REQUIRE(motifs && programs, "nmb_compile_motifs: null argument");
    nmb::compile_motifs_kernel<<<(2 * n_motifs + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        motifs, n_motifs, (nmb::Program *)programs);
    NMB_CUDA(cudaGetLastError());
    return NMB_OK;
}

int64_t nmb_family_scratch_bytes(int32_t n_motifs) {
    const int64_t n = n_motifs > 0 ? n_motifs : 0;
    return ((n * (int64_t)sizeof(nmb::FamInfo) + 127) / 128) * 128 + n * (int64_t)sizeof(nmb::Program);
}

static int scan_count_impl(const nmb_assembly *a, const uint32_t *class_records, const void *programs,
                           const nmb_job *jobs, int32_t n_jobs, int32_t n_items, int32_t motifs_per_item,
                           int32_t max_motif_len, const int32_t *contig_group, int64_t *out, int32_t grid_ctas,
                           const nmb_motif *motifs, int32_t n_motifs, void *family_scratch, int32_t *work_counter,
                           void *stream);

int nmb_scan_count(const nmb_assembly *a, const uint32_t *class_records, const void *programs,
                   const nmb_job *jobs, int32_t n_jobs, int32_t n_items, int32_t motifs_per_item,
                   int32_t max_motif_len, const int32_t *contig_group, int64_t *out,
                   int32_t grid_ctas, void *stream) {
    return scan_count_impl(a, class_records, programs, jobs, n_jobs, n_items, motifs_per_item, max_motif_len,
                           contig_group, out, grid_ctas, nullptr, 0, nullptr, nullptr, stream);
}

int nmb_scan_count_balanced(const nmb_assembly *a, const uint32_t *class_records, const void *programs,
                            const nmb_job *jobs, int32_t n_jobs, int32_t n_items, int32_t motifs_per_item,
                            int32_t max_motif_len, const int32_t *contig_group, int64_t *out, int32_t grid_ctas,
                            int32_t *work_counter, void *stream) {
    NMB_REQUIRE(work_counter, "nmb_scan_count_balanced: null work counter");
    return scan_count_impl(a, class_records, programs, jobs, n_jobs, n_items, motifs_per_item, max_motif_len,
                           contig_group, out, grid_ctas, nullptr, 0, nullptr, work_counter, stream);
}

int nmb_scan_count_families(const nmb_assembly *a, const uint32_t *class_records, const void *programs,
                            const nmb_job *jobs, int32_t n_jobs, int32_t n_items, int32_t motifs_per_item,
                            int32_t max_motif_len, const int32_t *contig_group, int64_t *out, int32_t grid_ctas,
                            const nmb_motif *motifs, int32_t n_motifs, void *family_scratch, void *stream) {
    NMB_REQUIRE(motifs && family_scratch && n_motifs > 0, "nmb_scan_count_families: null argument");
    return scan_count_impl(a, class_records, programs, jobs, n_jobs, n_items, motifs_per_item, max_motif_len,
                           contig_group, out, grid_ctas, motifs, n_motifs, family_scratch, nullptr, stream);
}

static int scan_count_impl(const nmb_assembly *a, const uint32_t *class_records, const void *programs,
                           const nmb_job *jobs, int32_t n_jobs, int32_t n_items, int32_t motifs_per_item,
                           int32_t max_motif_len, const int32_t *contig_group, int64_t *out, int32_t grid_ctas,
                           const nmb_motif *motifs, int32_t n_motifs, void *family_scratch, int32_t *work_counter,
                           void *stream) {
    NMB_REQUIRE(a && class_records && programs && jobs && out, "nmb_scan_count: null argument");
    NMB_REQUIRE(n_jobs > 0 && n_items >= 0, "nmb_scan_count: n_jobs=%d n_items=%d", n_jobs, n_items);
    NMB_REQUIRE(motifs_per_item >= 1 && motifs_per_item <= NMB_MAX_MOTIFS_PER_ITEM,
                "nmb_scan_count: motifs_per_item=%d not in 1..%d", motifs_per_item,
                NMB_MAX_MOTIFS_PER_ITEM);
    NMB_REQUIRE(max_motif_len >= 1 && max_motif_len <= NMB_MAX_MOTIF_LEN,
                "nmb_scan_count: max_motif_len=%d not in 1..%d", max_motif_len, NMB_MAX_MOTIF_LEN);
    if (n_items == 0) return NMB_OK;
    nmb::ScanParams p;
    p.seq_records = a->seq_records;
    p.nonacgt = a->nonacgt;
    p.cls = class_records;
    p.programs = (const nmb::Program *)programs;
    p.jobs = jobs;
    p.contig_group = contig_group;
    p.contig_start = a->contig_start;
    p.contig_len = a->contig_len;
    p.out = (unsigned long long *)out;
    p.n_jobs = n_jobs;
    p.n_items = n_items;
    p.mpi = motifs_per_item;
    p.n_tiles = a->n_tiles;
    p.counter = work_counter;
    p.fam = nullptr;
    p.parents = nullptr;
    if (family_scratch && motifs_per_item >= 2) {  // group the motif blocks into families (device, one block per job)
        p.fam = (const nmb::FamInfo *)family_scratch;
        p.parents = (const nmb::Program *)((uint8_t *)family_scratch +
                                           ((n_motifs * (int64_t)sizeof(nmb::FamInfo) + 127) / 128) * 128);
        nmb::find_families_kernel<<<dim3(n_jobs, 4), 256, 0, (cudaStream_t)stream>>>(
            motifs, jobs, motifs_per_item, (nmb::FamInfo *)p.fam, (nmb::Program *)p.parents);
        NMB_CUDA(cudaGetLastError());
    }

    int grid = grid_ctas;
    if (grid <= 0) {
        int sms = nmb_device_sm_count();
        if (sms < 0) return sms;
        grid = 4 * sms;  // resident CTAs per SM (49.7 KB of shared memory, <= 128 registers)
    }
    if (grid > n_items) grid = n_items;
    cudaStream_t s = (cudaStream_t)stream;
    const int smem_bytes = nmb::kScanSmemBytes;
#define NMB_LAUNCH_SCAN(H, S, F)                                                                                      \
    do {                                                                                                              \
        NMB_CUDA(cudaFuncSetAttribute(nmb::scan_count_kernel<H, S, F>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                      smem_bytes));                                                                   \
        nmb::scan_count_kernel<H, S, F><<<grid, nmb::kScanThreads, smem_bytes, s>>>(p);                                \
    } while (0)
    const bool fam = p.fam != nullptr;
    if (max_motif_len <= 32) {  // one halo word covers a total shift (and a mod_pos) of at most 31
        if (fam) NMB_LAUNCH_SCAN(1, 1, true); else NMB_LAUNCH_SCAN(1, 1, false);
    } else {
        if (fam) NMB_LAUNCH_SCAN(2, 1, true); else NMB_LAUNCH_SCAN(2, 1, false);
    }
#undef NMB_LAUNCH_SCAN
    NMB_CUDA(cudaGetLastError());
    return NMB_OK;
}

}  // extern "C"
