// Track bookkeeping between the temporal LK and the corner top-up, one CTA per stream:
//   ReduceVector (stable compaction by status)          vins/estimator/feature_tracker.hpp:74-81, .cpp:69-71
//   ++track_count_                                       .cpp:76-79
//   SetFeatureExtractorMask: std::sort by track count + greedy min_dist keep   .cpp:219-246
//
// The reference's std::sort is unstable and its comparator sees the track count only, so the order among
// equally old features is whatever libstdc++'s introsort produces.  That order decides which of two close
// features survives and the order of ids_, so it is reproduced exactly (introsort_emul.h): one warp replays
// the introsort loop (median-of-3 + Hoare partition on ranges > 16 with the partition itself evaluated in
// parallel, heap-sort fallback), and because the final insertion sort is stable the finish is a parallel
// stable rank by track count.
//
// The greedy keep ("visit in sorted order, keep a feature iff no kept feature lies within min_dist of its
// rounded centre", i.e. the filled cv::circle test dx^2+dy^2 <= r^2) is evaluated with a conflict bit-matrix
// and parallel rounds that reach the sequential answer: a feature is rejected once an earlier conflicting one
// is accepted, accepted once no earlier conflicting one is still undecided.
#include "introsort_emul.h"
#include "vf_common.cuh"
#include "vf_kernels.h"

namespace vf {

// one warp per SM sub-partition pair is enough for the latency-bound small case; the quadratic phases (stable rank, conflict
// matrix, rounds) of a large capacity want every thread an SM can hold
constexpr int ST_THREADS_SMALL = 256, ST_THREADS_LARGE = 1024, ST_LARGE_CAP = 512;
constexpr int kConflictSmemCap = 256;   // up to this many features: dense conflict bit-matrix in shared memory; above: spatial grid
constexpr int kGridCells = 4096;        // cells of the spatial grid (cell edge = power of two >= min_dist, chosen so that the image fits)

// dynamic shared memory layout for `cap` features
struct SelLayout {
    size_t off_pts, off_ids, off_cnt, off_items, off_order, off_src, off_state, off_lists, off_words, off_conf, off_grid, total;
    int words;
    __host__ __device__ explicit SelLayout(int cap) {
        words = (cap + 31) / 32;
        size_t o = 0;
        off_pts = o;   o += sizeof(float2) * cap;
        off_items = o; o += sizeof(SortItem) * cap;
        off_ids = o;   o += sizeof(int) * cap;
        off_cnt = o;   o += sizeof(int) * cap;
        off_order = o; o += sizeof(int) * cap;
        off_src = o;   o += sizeof(int) * cap;                  // index of the compacted feature in the previous frame's arrays
        off_state = o; o += sizeof(int) * cap;                  // rounded centres packed (x | y << 16) after sorting
        off_lists = o; o += sizeof(int) * (cap * 2 + 32);       // left / right stopper lists of the parallel partition; later the
                                                                // sentinel-padded key array of the final rank (cap + 32)
        off_words = o; o += sizeof(uint32_t) * words * 4;       // accepted / undecided masks, double use
        off_conf = o;  o += (cap <= kConflictSmemCap) ? sizeof(uint32_t) * (size_t)cap * words : 0;
        off_grid = o;  o += (cap <= kConflictSmemCap) ? 0 : sizeof(int) * (2 * kGridCells + 2);   // cell starts (+1), fill cursors
        total = o;
    }
};

size_t select_tracked_smem_bytes(int cap) { return SelLayout(cap > 0 ? cap : 1).total + 64; }

// std::__introsort_loop replayed exactly, one warp per range.  The control flow (depth limit, median-of-3, heap-sort
// fallback, "recurse into [cut,last), continue with [first,cut)") is the sequential algorithm of introsort_emul.h; since
// sub-ranges are disjoint, the order in which they are processed does not change the result, so every level of the
// recursion tree is processed in parallel (one warp per pending range, ranges handed from level to level through a
// shared-memory queue) and the chain is the tree depth (5 for 150 features) instead of the number of partitions (13).
// std::__unguarded_partition itself is evaluated in parallel inside the warp:
//   L = positions in (first, last) whose key does not precede the pivot's  (where the left scan stops),  ascending
//   R = positions in (first, last) which the pivot does not precede        (where the right scan stops), descending
//   the k-th swap of the sequential loop exchanges L[k] and R[k] for every k with L[k] < R[k]  (m of them), and the
//   returned cut is min(L[m], R[m-1]).
// (both scans only ever look at elements no swap has touched yet; tests/cpp/test_introsort.cpp checks the sequential
//  form against the real std::sort, the GPU tests check this form against both.)
__device__ __forceinline__ int warp_partition_pivot(SortItem* v, int first, int last, int* Lpos, int* Rpos) {
    const int lane = threadIdx.x & 31;
    const unsigned int lt = (1u << lane) - 1u;
    // std::__move_median_to_first(first, first+1, mid, last-1): every lane reads the three keys (broadcast loads,
    // one round trip) and reaches the same verdict; lane 0 performs the swap
    const int ia = first + 1, ib = first + (last - first) / 2, ic = last - 1;
    const int ka = v[ia].key, kb = v[ib].key, kc = v[ic].key;
    int im;
    if (ka > kb) im = (kb > kc) ? ib : ((ka > kc) ? ic : ia);
    else im = (ka > kc) ? ia : ((kb > kc) ? ic : ib);
    const int pk = (im == ia) ? ka : (im == ib ? kb : kc);     // key of the pivot that ends up in v[first]
    if (lane == 0) { const SortItem t = v[first]; v[first] = v[im]; v[im] = t; }
    __syncwarp();
    // both stopper lists from ONE pass over (first, last), 8 chunks of 32 keys in flight at a time; R is kept ascending
    // in Rpos (the sequential right scan meets its stoppers in descending order: R[k] = Rpos[nR-1-k])
    int nL = 0, nR = 0;
    for (int base = first + 1; base < last; base += 256) {
        int key[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int p = base + 32 * q + lane;
            key[q] = p < last ? v[p].key : 0;
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int p = base + 32 * q + lane;
            if (base + 32 * q < last) {                                   // uniform
                const bool in = p < last;
                const bool sl = in && !(key[q] > pk), sr = in && !(pk > key[q]);
                const unsigned int bl = __ballot_sync(0xffffffffu, sl);
                const unsigned int br = __ballot_sync(0xffffffffu, sr);
                if (sl) Lpos[nL + __popc(bl & lt)] = p;
                if (sr) Rpos[nR + __popc(br & lt)] = p;
                nL += __popc(bl); nR += __popc(br);
            }
        }
    }
    __syncwarp();
    const int K = min(nL, nR);
    int m = 0;
    for (int base = 0; base < K; base += 32) {
        const int k = base + lane;
        const bool ok = k < K && Lpos[k] < Rpos[nR - 1 - k];
        const unsigned int bal = __ballot_sync(0xffffffffu, ok);
        if (bal == 0xffffffffu) { m += 32; continue; }
        m += __ffs(~bal) - 1;
        break;
    }
    for (int k = lane; k < m; k += 32) {
        const int a = Lpos[k], c = Rpos[nR - 1 - k];
        const SortItem t = v[a]; v[a] = v[c]; v[c] = t;
    }
    const int a_m = m < nL ? Lpos[m] : 0x7fffffff, b_prev = m > 0 ? Rpos[nR - m] : 0x7fffffff;
    __syncwarp();
    return min(a_m, b_prev);
}

constexpr int kSortQueue = 256;   // pending ranges of one level: at most cap / 17 (cap <= 4096)

// All warps of the CTA.  queue: [2][kSortQueue] int4 (first, last, depth); qn: [2] counts; Lpos / Rpos: [cap] scratch
// (a range only uses the slots of its own positions).
__device__ void block_introsort_loop(SortItem* v, int n, int4* queue, int* qn, int* Lpos, int* Rpos) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (threadIdx.x == 0) {
        qn[0] = n > 16 ? 1 : 0; qn[1] = 0;
        queue[0] = make_int4(0, n, 2 * floor_log2(n > 0 ? n : 1), 0);
    }
    __syncthreads();
    int cur = 0;
    while (qn[cur] > 0) {
        const int cnt = qn[cur];
        for (int i = wid; i < cnt; i += nw) {
            const int4 r = queue[cur * kSortQueue + i];
            const int first = r.x, last = r.y, depth = r.z;
            if (depth == 0) {
                if (lane == 0) heap_sort_range(v + first, last - first);
            } else {
                const int cut = warp_partition_pivot(v, first, last, Lpos + first, Rpos + first);
                if (lane == 0) {
                    if (last - cut > 16) queue[(cur ^ 1) * kSortQueue + atomicAdd(&qn[cur ^ 1], 1)] = make_int4(cut, last, depth - 1, 0);
                    if (cut - first > 16) queue[(cur ^ 1) * kSortQueue + atomicAdd(&qn[cur ^ 1], 1)] = make_int4(first, cut, depth - 1, 0);
                }
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) qn[cur] = 0;
        cur ^= 1;
        __syncthreads();
    }
}

template <int ST_THREADS>
__global__ void __launch_bounds__(ST_THREADS) select_tracked_kernel(const __grid_constant__ DevBatch b, int parity_arg) {
    pdl_wait(); pdl_release();
    const int parity = parity_arg & 1, tphase = parity_arg >> 4;   // frame parity; trace slot (frame % 6)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int warp_cnt[ST_THREADS / 32];
    __shared__ int4 sort_queue[2 * kSortQueue];
    __shared__ int sort_qn[2];
    TraceScope trace(b.trace, TR_SELECT + 16 * tphase);
    const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, cap = b.cap;
    const SelLayout L(cap);
    float2* pts = reinterpret_cast<float2*>(smem_raw + L.off_pts);
    SortItem* items = reinterpret_cast<SortItem*>(smem_raw + L.off_items);
    int* ids = reinterpret_cast<int*>(smem_raw + L.off_ids);
    int* cnts = reinterpret_cast<int*>(smem_raw + L.off_cnt);
    int* order = reinterpret_cast<int*>(smem_raw + L.off_order);
    int* src = reinterpret_cast<int*>(smem_raw + L.off_src);
    int* centre = reinterpret_cast<int*>(smem_raw + L.off_state);
    int* lists = reinterpret_cast<int*>(smem_raw + L.off_lists);
    uint32_t* accw = reinterpret_cast<uint32_t*>(smem_raw + L.off_words);
    uint32_t* undw = accw + L.words;
    uint32_t* conf = reinterpret_cast<uint32_t*>(smem_raw + L.off_conf);   // dense path only (cap <= kConflictSmemCap)
    const bool use_grid = cap > kConflictSmemCap;
    int* cell_start = reinterpret_cast<int*>(smem_raw + L.off_grid);       // [n_cell + 1], grid path only
    int* cell_cur = cell_start + kGridCells + 1;                           // [n_cell]
    int* cell_items = lists;                                               // [n_tr] ranks grouped by cell (the partition scratch is free by then)
    const int words = L.words;
    int* counters = b.counters[parity] + s * C_COUNT;
    const int n_prev = b.counters[parity ^ 1][s * C_COUNT + C_N_TOTAL];
    if (tid < C_COUNT) counters[tid] = 0;   // first kernel of the frame that touches them: the atomics of the later stages start from 0
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
            src[j] = i;
            items[j].key = c; items[j].payload = j;
        }
        n_tr += tot;
        __syncthreads();
    }

    trace_phase(b.trace, 8);
    // ---- 2. libstdc++ std::sort order: introsort loop (sequential replay) + stable rank (parallel) ----
    block_introsort_loop(items, n_tr, sort_queue, sort_qn, lists, lists + cap);
    trace_phase(b.trace, 9);
    // What is left is std::__final_insertion_sort, a stable insertion: r(j) = #{k : key_k > key_j} + #{k < j : key_k == key_j}
    //   = j - #{k < j : key_k < key_j} + #{k > j : key_k > key_j}.
    // The introsort loop only leaves partitions of at most 16 elements unsorted and no key of an earlier partition is
    // smaller (none of a later one larger), so both sets lie within 15 positions of j: 30 compares per element instead
    // of n.  The keys go to a dense array padded with +inf / -inf sentinels on either side.
    int* dense = lists;                     // the partition's scratch is free again (2 * cap ints; cap >= 1)
    for (int j = tid; j < n_tr + 32; j += ST_THREADS)
        dense[j] = j < 16 ? 0x7fffffff : (j - 16 < n_tr ? items[j - 16].key : (int)0x80000000);
    __syncthreads();
    for (int j = tid; j < n_tr; j += ST_THREADS) {
        const int* kp = dense + 16 + j;
        const int kj = kp[0];
        int lt = 0, gt = 0;
#pragma unroll
        for (int d = 1; d <= 15; ++d) { lt += kp[-d] < kj; gt += kp[d] > kj; }
        order[j - lt + gt] = items[j].payload;
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

    trace_phase(b.trace, 10);
    const int R2 = b.min_dist * b.min_dist;
    int gsh = 0, gw = 1, gh = 1;
    if (use_grid) {
        // ---- 3a. many features: bin the rounded centres into a grid whose cells are at least min_dist wide -- everything
        //          within min_dist of a feature then lies in the 3x3 cells around it, and with features that were thinned to
        //          min_dist a frame ago that is a handful of candidates instead of a row of an n x n matrix ----
        while ((1 << gsh) < b.min_dist || (((b.W - 1) >> gsh) + 1) * (((b.H - 1) >> gsh) + 1) > kGridCells) ++gsh;
        gw = ((b.W - 1) >> gsh) + 1; gh = ((b.H - 1) >> gsh) + 1;
        const int n_cell = gw * gh;
        for (int c = tid; c <= n_cell; c += ST_THREADS) cell_start[c] = 0;
        __syncthreads();
        for (int r = tid; r < n_tr; r += ST_THREADS) {
            const int cc = centre[r];
            atomicAdd(&cell_start[(cc >> 16 >> gsh) * gw + ((int)(short)(cc & 0xffff) >> gsh) + 1], 1);   // count of cell c at c + 1
        }
        __syncthreads();
        // inclusive scan of the counts: cell_start[c] = first slot of cell c
        const int ipt = (n_cell + ST_THREADS) / ST_THREADS, i0 = min(1 + tid * ipt, n_cell + 1), i1 = min(i0 + ipt, n_cell + 1);
        int sum = 0;
        for (int i = i0; i < i1; ++i) sum += cell_start[i];
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) warp_cnt[wid] = incl;
        __syncthreads();
        int before = incl - sum;
        for (int w = 0; w < wid; ++w) before += warp_cnt[w];
        for (int i = i0; i < i1; ++i) { before += cell_start[i]; cell_start[i] = before; }
        __syncthreads();
        for (int c = tid; c < n_cell; c += ST_THREADS) cell_cur[c] = cell_start[c];
        __syncthreads();
        for (int r = tid; r < n_tr; r += ST_THREADS) {
            const int cc = centre[r];
            cell_items[atomicAdd(&cell_cur[(cc >> 16 >> gsh) * gw + ((int)(short)(cc & 0xffff) >> gsh)], 1)] = r;
        }
    } else {
    // ---- 3. conflict bits: conf[r][w] bit k set iff feature 32w+k precedes r in sorted order and lies within min_dist ----
    // Work item of a WARP = (block B of 32 rows, word w <= B) of the lower triangle: the 32 lanes take the rows of the block
    // and share the word, so centre[32w + k] is one broadcast load (lanes with different w would all hit bank k) and the
    // upper triangle is never visited.
        const int n_pairs = words * (words + 1) / 2, nw = ST_THREADS / 32;
        for (int id = wid; id < n_pairs; id += nw) {
            int B = (int)((sqrtf(8.f * (float)id + 1.f) - 1.f) * 0.5f);
            while (B * (B + 1) / 2 > id) --B;
            while ((B + 1) * (B + 2) / 2 <= id) ++B;
            const int w = id - B * (B + 1) / 2, r = 32 * B + lane;
            if (r < n_tr) {
                uint32_t bits = 0u;
                const int cx = (int)(short)(centre[r] & 0xffff), cy = centre[r] >> 16;
                const int kend = w == B ? lane : 32;
#pragma unroll 8
                for (int k = 0; k < kend; ++k) {
                    const int c2 = centre[32 * w + k];
                    const int dx = cx - (int)(short)(c2 & 0xffff), dy = cy - (c2 >> 16);
                    if (dx * dx + dy * dy <= R2) bits |= 1u << k;
                }
                conf[(size_t)r * words + w] = bits;
            }
        }
    }
    __syncthreads();

    trace_phase(b.trace, 11);
    // ---- 4. rounds: reject if an earlier conflicting feature is accepted; accept if none is still undecided ----
    while (true) {
        bool any_undecided = false;
        for (int base = 0; base < n_tr; base += ST_THREADS) {
            const int r = base + tid;
            int decision = 0;   // 0 none, 1 accept, 2 reject
            if (r < n_tr && ((undw[r >> 5] >> (r & 31)) & 1u)) {
                uint32_t hit_acc = 0u, hit_und = 0u;
                if (use_grid) {
                    const int cc = centre[r], cx = (int)(short)(cc & 0xffff), cy = cc >> 16, gx = cx >> gsh, gy = cy >> gsh;
                    for (int ny = max(gy - 1, 0); ny <= min(gy + 1, gh - 1); ++ny) {
                        // the three cells of a grid row are neighbours in memory: one run of items
                        const int i_end = cell_start[ny * gw + min(gx + 1, gw - 1) + 1];
                        for (int i = cell_start[ny * gw + max(gx - 1, 0)]; i < i_end; ++i) {
                            const int r2 = cell_items[i];
                            const int c2 = centre[r2];
                            const int dx = cx - (int)(short)(c2 & 0xffff), dy = cy - (c2 >> 16);
                            if (r2 < r && dx * dx + dy * dy <= R2) {           // an earlier feature within min_dist
                                hit_acc |= (accw[r2 >> 5] >> (r2 & 31)) & 1u;
                                hit_und |= (undw[r2 >> 5] >> (r2 & 31)) & 1u;
                            }
                        }
                    }
                } else {
#pragma unroll 4
                    for (int w = 0; w <= (r >> 5); ++w) {
                        const uint32_t c = conf[(size_t)r * words + w];
                        hit_acc |= c & accw[w];
                        hit_und |= c & undw[w];
                    }
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

    trace_phase(b.trace, 12);
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
            b.prev_idx[parity][so + k] = src[j];
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
    const void* fn = cap > ST_LARGE_CAP ? (const void*)select_tracked_kernel<ST_THREADS_LARGE> : (const void*)select_tracked_kernel<ST_THREADS_SMALL>;
    const cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, kCarveout);
    if (e != cudaSuccess) return e;
    if (smem <= 48 * 1024) return cudaSuccess;
    return cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

void launch_select_tracked(const DevBatch* d_batch, const DevBatch& h, int parity, cudaStream_t stream) {
    if (h.cap > ST_LARGE_CAP) launch_kernel(select_tracked_kernel<ST_THREADS_LARGE>, dim3(h.n_streams), dim3(ST_THREADS_LARGE), select_tracked_smem_bytes(h.cap), stream, h, parity);
    else launch_kernel(select_tracked_kernel<ST_THREADS_SMALL>, dim3(h.n_streams), dim3(ST_THREADS_SMALL), select_tracked_smem_bytes(h.cap), stream, h, parity);
}

}  // namespace vf
