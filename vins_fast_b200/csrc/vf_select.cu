// Track bookkeeping between the temporal LK and the corner top-up, one CTA per stream:
//   ReduceVector (stable compaction by status)          vins/estimator/feature_tracker.hpp:74-81, .cpp:69-71
//   ++track_count_                                       .cpp:76-79
//   SetFeatureExtractorMask: std::sort by track count + greedy min_dist keep   .cpp:219-246
//
// The reference's std::sort is unstable and its comparator sees the track count only, so the order among
// equally old features is whatever libstdc++'s introsort produces.  That order decides which of two close
// features survives and the order of ids_, so it is reproduced exactly (introsort_emul.h): thread 0 replays
// the introsort loop (median-of-3 + Hoare partition on ranges > 16, heap-sort fallback), and because the final
// insertion sort is stable the finish is a parallel stable rank by track count.
//
// The greedy keep ("visit in sorted order, keep a feature iff no kept feature lies within min_dist of its
// rounded centre", i.e. the filled cv::circle test dx^2+dy^2 <= r^2) is evaluated with a conflict bit-matrix
// and parallel rounds that reach the sequential answer: a feature is rejected once an earlier conflicting one
// is accepted, accepted once no earlier conflicting one is still undecided.
#include "introsort_emul.h"
#include "vf_common.cuh"
#include "vf_kernels.h"

namespace vf {

constexpr int ST_THREADS = 256;
constexpr int kConflictSmemCap = 512;   // conflict matrix lives in shared memory up to this many features

// dynamic shared memory layout for `cap` features
struct SelLayout {
    size_t off_pts, off_ids, off_cnt, off_items, off_order, off_state, off_words, off_conf, total;
    int words;
    __host__ __device__ explicit SelLayout(int cap) {
        words = (cap + 31) / 32;
        size_t o = 0;
        off_pts = o;   o += sizeof(float2) * cap;
        off_items = o; o += sizeof(SortItem) * cap;
        off_ids = o;   o += sizeof(int) * cap;
        off_cnt = o;   o += sizeof(int) * cap;
        off_order = o; o += sizeof(int) * cap;
        off_state = o; o += sizeof(int) * cap;                  // rounded centres packed (x | y << 16) after sorting
        off_words = o; o += sizeof(uint32_t) * words * 4;       // accepted / undecided masks, double use
        off_conf = o;  o += (cap <= kConflictSmemCap) ? sizeof(uint32_t) * (size_t)cap * words : 0;
        total = o;
    }
};

size_t select_tracked_smem_bytes(int cap) { return SelLayout(cap > 0 ? cap : 1).total + 64; }

__global__ void __launch_bounds__(ST_THREADS) select_tracked_kernel(const DevBatch* __restrict__ bp, int parity) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int warp_cnt[ST_THREADS / 32];
    __shared__ int stack[3 * 40];
    const DevBatch& b = *bp;
    const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, cap = b.cap;
    const SelLayout L(cap);
    float2* pts = reinterpret_cast<float2*>(smem_raw + L.off_pts);
    SortItem* items = reinterpret_cast<SortItem*>(smem_raw + L.off_items);
    int* ids = reinterpret_cast<int*>(smem_raw + L.off_ids);
    int* cnts = reinterpret_cast<int*>(smem_raw + L.off_cnt);
    int* order = reinterpret_cast<int*>(smem_raw + L.off_order);
    int* centre = reinterpret_cast<int*>(smem_raw + L.off_state);
    uint32_t* accw = reinterpret_cast<uint32_t*>(smem_raw + L.off_words);
    uint32_t* undw = accw + L.words;
    uint32_t* conf = (cap <= kConflictSmemCap) ? reinterpret_cast<uint32_t*>(smem_raw + L.off_conf)
                                               : b.conflict + (size_t)s * cap * L.words;
    const int words = L.words;
    int* counters = b.counters[parity] + s * C_COUNT;
    const int n_prev = b.counters[parity ^ 1][s * C_COUNT + C_N_TOTAL];
    const size_t so = (size_t)s * cap;

    // ---- 1. ReduceVector by the temporal status (order preserving) and ++track_count ----
    int n_tr = 0;
    for (int base = 0; base < n_prev; base += ST_THREADS) {
        const int i = base + tid;
        const bool keep = i < n_prev && b.temporal_st[so + i] != 0;
        const unsigned int bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) warp_cnt[wid] = __popc(bal);
        __syncthreads();
        int off = 0, tot = 0;
        for (int w = 0; w < ST_THREADS / 32; ++w) { const int c = warp_cnt[w]; if (w < wid) off += c; tot += c; }
        if (keep) {
            const int j = n_tr + off + __popc(bal & ((1u << lane) - 1u));
            pts[j] = b.fwd_pts[so + i];
            ids[j] = b.ids[parity ^ 1][so + i];
            const int c = b.cnt[parity ^ 1][so + i] + 1;
            cnts[j] = c;
            items[j].key = c; items[j].payload = j;
        }
        n_tr += tot;
        __syncthreads();
    }

    // ---- 2. libstdc++ std::sort order: introsort loop (sequential replay) + stable rank (parallel) ----
    if (tid == 0 && n_tr > 16) vf_introsort_loop(items, n_tr, stack);
    __syncthreads();
    for (int j = tid; j < n_tr; j += ST_THREADS) {
        const int kj = items[j].key;
        int r = 0;
        for (int k = 0; k < n_tr; ++k) {
            const int kk = items[k].key;
            r += (kk > kj) || (kk == kj && k < j);
        }
        order[r] = items[j].payload;
    }
    __syncthreads();
    for (int r = tid; r < n_tr; r += ST_THREADS) {
        const int j = order[r];
        b.sort_perm[so + r] = j;
        const float2 p = pts[j];
        centre[r] = (__float2int_rn(p.x) & 0xffff) | (__float2int_rn(p.y) << 16);   // cvRound; points are InBorder
    }
    for (int w = tid; w < words; w += ST_THREADS) {
        accw[w] = 0u;
        const int rem = n_tr - 32 * w;
        undw[w] = rem >= 32 ? 0xffffffffu : (rem > 0 ? ((1u << rem) - 1u) : 0u);
        accw[2 * words + w] = 0u; undw[2 * words + w] = 0u;   // staging halves
    }
    __syncthreads();

    // ---- 3. conflict bits: conf[r][w] bit k set iff feature 32w+k precedes r in sorted order and lies within min_dist ----
    const int R2 = b.min_dist * b.min_dist;
    for (int i = tid; i < n_tr * words; i += ST_THREADS) {
        const int r = i / words, w = i - r * words;
        uint32_t bits = 0u;
        const int cx = (int)(short)(centre[r] & 0xffff), cy = centre[r] >> 16;
        const int kend = min(32, r - 32 * w);
        for (int k = 0; k < kend; ++k) {
            const int c2 = centre[32 * w + k];
            const int dx = cx - (int)(short)(c2 & 0xffff), dy = cy - (c2 >> 16);
            if (dx * dx + dy * dy <= R2) bits |= 1u << k;
        }
        conf[(size_t)r * words + w] = bits;
    }
    __syncthreads();

    // ---- 4. rounds: reject if an earlier conflicting feature is accepted; accept if none is still undecided ----
    while (true) {
        bool any_undecided = false;
        for (int base = 0; base < n_tr; base += ST_THREADS) {
            const int r = base + tid;
            int decision = 0;   // 0 none, 1 accept, 2 reject
            if (r < n_tr && ((undw[r >> 5] >> (r & 31)) & 1u)) {
                bool hit_acc = false, hit_und = false;
                for (int w = 0; w <= (r >> 5); ++w) {
                    const uint32_t c = conf[(size_t)r * words + w];
                    hit_acc = hit_acc || (c & accw[w]);
                    hit_und = hit_und || (c & undw[w]);
                }
                decision = hit_acc ? 2 : (hit_und ? 0 : 1);
                any_undecided = any_undecided || decision == 0;
            }
            const unsigned int acc_b = __ballot_sync(0xffffffffu, decision == 1);
            const unsigned int dec_b = __ballot_sync(0xffffffffu, decision != 0);
            // staged update: written to the second half so that the other slabs of this round still see the old masks
            if (lane == 0 && base + wid * 32 < n_tr) {
                accw[2 * words + (base >> 5) + wid] = acc_b;
                undw[2 * words + (base >> 5) + wid] = dec_b;
            }
        }
        __syncthreads();
        for (int w = tid; w < words; w += ST_THREADS) {
            accw[w] |= accw[2 * words + w];
            undw[w] &= ~undw[2 * words + w];
        }
        if (!__syncthreads_or(any_undecided)) break;
    }

    // ---- 5. write the kept features in sorted order ----
    int n_kept = 0;
    for (int base = 0; base < n_tr; base += ST_THREADS) {
        const int r = base + tid;
        const bool keep = r < n_tr && ((accw[r >> 5] >> (r & 31)) & 1u);
        const unsigned int bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) warp_cnt[wid] = __popc(bal);
        __syncthreads();
        int off = 0, tot = 0;
        for (int w = 0; w < ST_THREADS / 32; ++w) { const int c = warp_cnt[w]; if (w < wid) off += c; tot += c; }
        if (keep) {
            const int k = n_kept + off + __popc(bal & ((1u << lane) - 1u));
            const int j = order[r];
            b.kps[parity][so + k] = pts[j];
            b.ids[parity][so + k] = ids[j];
            b.cnt[parity][so + k] = cnts[j];
        }
        n_kept += tot;
        __syncthreads();
    }
    if (tid == 0) {
        counters[C_N_PREV] = n_prev;
        counters[C_N_TRACKED] = n_tr;
        counters[C_N_KEPT] = n_kept;
    }
}

// Opt in to > 48 KB dynamic shared memory once per device (not a stream operation; done at batch creation,
// outside any stream capture).
cudaError_t select_tracked_prepare(int cap) {
    const size_t smem = select_tracked_smem_bytes(cap);
    if (smem <= 48 * 1024) return cudaSuccess;
    return cudaFuncSetAttribute(select_tracked_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

void launch_select_tracked(const DevBatch* d_batch, const DevBatch& h, int parity, cudaStream_t stream) {
    select_tracked_kernel<<<h.n_streams, ST_THREADS, select_tracked_smem_bytes(h.cap), stream>>>(d_batch, parity);
}

}  // namespace vf
