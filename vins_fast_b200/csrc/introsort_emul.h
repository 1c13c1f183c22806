// Emulation of libstdc++'s std::sort tie order for the comparator of the reference's
// SetFeatureExtractorMask (vins/estimator/feature_tracker.cpp:228-231):
//     std::sort(v.begin(), v.end(), [](a, b) { return a.first > b.first; });   // first = track count
// The comparator sees the track count only and std::sort is unstable, so which of two equally old
// features is visited first (and therefore survives the min_dist mask, and the order of ids_) is decided
// by libstdc++'s introsort.  To keep the reference's ordering bit for bit without a device->host round
// trip in the middle of the frame, the device reproduces that algorithm (bits/stl_algo.h, GCC 4.7 .. 14):
//
//   __sort            = __introsort_loop(first, last, 2*floor(log2 n)) ; __final_insertion_sort
//   __introsort_loop  : while (last-first > 16) { depth==0 -> heap sort of the range, return;
//                       cut = __unguarded_partition_pivot ; recurse on [cut,last) ; last = cut }
//   __final_insertion_sort : (guarded + unguarded) insertion sort.  Insertion sort is STABLE, so its
//                       result is the stable sort of whatever order the introsort loop left behind.
//
// vf_introsort_loop() below replays the loop phase (median-of-3, Hoare partition, heap-sort fallback)
// on an array of (key, payload); the caller finishes with a stable rank by key, which a GPU does in
// parallel.  Host+device code so the CPU test-suite can check it against the real std::sort.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define VF_HD __host__ __device__ __forceinline__
#else
#define VF_HD inline
#endif

namespace vf {

struct SortItem {
    int32_t key;      // track count
    int32_t payload;  // original index
};

// comp(a, b) of the reference: a.first > b.first
VF_HD bool is_before(const SortItem& a, const SortItem& b) { return a.key > b.key; }

VF_HD void swap_items(SortItem* v, int i, int j) {
    SortItem t = v[i]; v[i] = v[j]; v[j] = t;
}

VF_HD int floor_log2(int n) {
    int l = 0;
    while (n > 1) { n >>= 1; ++l; }
    return l;
}

// std::__adjust_heap + std::__push_heap
VF_HD void adjust_heap(SortItem* first, int hole, int len, SortItem value) {
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (is_before(first[child], first[child - 1])) child--;
        first[hole] = first[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        first[hole] = first[child - 1];
        hole = child - 1;
    }
    int parent = (hole - 1) / 2;
    while (hole > top && is_before(first[parent], value)) {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}

// std::__partial_sort(first, last, last) == make_heap + sort_heap
VF_HD void heap_sort_range(SortItem* first, int len) {
    if (len >= 2) {
        int parent = (len - 2) / 2;
        while (true) {
            SortItem v = first[parent];
            adjust_heap(first, parent, len, v);
            if (parent == 0) break;
            parent--;
        }
    }
    int last = len;
    while (last > 1) {
        --last;
        SortItem v = first[last];
        first[last] = first[0];
        adjust_heap(first, 0, last, v);
    }
}

// std::__move_median_to_first(result, a, b, c)
VF_HD void move_median_to_first(SortItem* v, int result, int a, int b, int c) {
    if (is_before(v[a], v[b])) {
        if (is_before(v[b], v[c])) swap_items(v, result, b);
        else if (is_before(v[a], v[c])) swap_items(v, result, c);
        else swap_items(v, result, a);
    } else if (is_before(v[a], v[c])) swap_items(v, result, a);
    else if (is_before(v[b], v[c])) swap_items(v, result, c);
    else swap_items(v, result, b);
}

// std::__unguarded_partition(first, last, pivot)
VF_HD int unguarded_partition(SortItem* v, int first, int last, int pivot) {
    while (true) {
        while (is_before(v[first], v[pivot])) ++first;
        --last;
        while (is_before(v[pivot], v[last])) --last;
        if (!(first < last)) return first;
        swap_items(v, first, last);
        ++first;
    }
}

// std::__introsort_loop on v[0, n) with an explicit stack (the recursion goes into [cut,last) first,
// then continues on [first,cut); sub-ranges are disjoint so the visiting order does not change the result).
// stack: caller-provided scratch of at least 3*(2*floor_log2(n)+2) ints.
VF_HD void vf_introsort_loop(SortItem* v, int n, int* stack) {
    if (n <= 1) return;
    int sp = 0;
    stack[sp++] = 0; stack[sp++] = n; stack[sp++] = 2 * floor_log2(n);
    while (sp > 0) {
        int depth = stack[--sp];
        int last = stack[--sp];
        int first = stack[--sp];
        while (last - first > 16) {
            if (depth == 0) { heap_sort_range(v + first, last - first); break; }
            --depth;
            int mid = first + (last - first) / 2;
            move_median_to_first(v, first, first + 1, mid, last - 1);
            int cut = unguarded_partition(v, first + 1, last, first);
            stack[sp++] = cut; stack[sp++] = last; stack[sp++] = depth;   // "recursive" call on [cut,last)
            last = cut;
        }
    }
}

}  // namespace vf
